import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from helpers import build, prior_u, oracle_cld_sample
from gddim_b200.cld import sampling, sde_lib
from oracle import cld as oc
inv = lambda x: (x + 1.) / 2.
cfg, model, net_fn = build("cld_deep")
sde = sde_lib.from_config(cfg)
for order, nfe, seed in [(2, 8, 2), (0, 6, 9), (0, 6, 0)]:
  fn = sampling.get_deis_sampler(sde, model, (32, 32, 3), nfe, inv, order, ts_order=2, denoising=True)
  fn.core.use_graph = False
  u = prior_u(2, seed=seed)
  x, v, n, tr = fn(0, model, 2, u=u, trace=True)
  print("order", order, "seed", seed, "nan in x:", np.isnan(x).sum(), "trace nan:", np.isnan(tr).sum(), "|u_last|max", np.abs(tr[-1]).max())
  ul = tr[-1]
  xin = oc.relayout_in(ul)
  for t in (1e-3, 2e-3, 0.01):
    g = model.forward(xin, t)
    w = net_fn(xin, 999.0 * t)
    print("  t", t, "gpu nan", np.isnan(g).sum(), "inf", np.isinf(g).sum(), "|g|max", np.nanmax(np.abs(g)), "|w|max", np.abs(w).max(),
          "rel", np.linalg.norm(np.nan_to_num(g) - w) / np.linalg.norm(w))
  model.set_gemm_impl(1)
  g = model.forward(xin, 1e-3)
  print("  ref impl nan", np.isnan(g).sum())
  model.set_gemm_impl(0)
