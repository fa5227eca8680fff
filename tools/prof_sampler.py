"""A short sampler call for ncu (update / DCT kernels): usage prof_sampler.py cld|blur [batch] [nfe] [deep]
(`deep`: the full accr_dcifar10 network instead of the narrowed one -- for captures of the head convolution + update)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from gddim_b200 import configs, net
kind = sys.argv[1] if len(sys.argv) > 1 else "cld"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nfe = int(sys.argv[3]) if len(sys.argv) > 3 else 4
inv = lambda x: (x + 1.) / 2.
rng = np.random.default_rng(0)
if kind == "cld":
  from gddim_b200.cld import sampling, sde_lib
  cfg = configs.cld_accr_dcifar10()
  if not (len(sys.argv) > 4 and sys.argv[4] == "deep"):
    cfg.model.nf, cfg.model.num_res_blocks = 64, 1      # small net: the update kernels are what is profiled
  model = net.ScoreNet(cfg, cld=True); model.init_params(seed=1, nondegenerate=True)
  fn = sampling.get_deis_sampler(sde_lib.from_config(cfg), model, (32, 32, 3), nfe, inv, 2, ts_order=2, denoising=True)
  fn.core.use_graph = False
  u = torch.from_numpy(np.stack([rng.standard_normal((B, 32, 32, 3)), rng.standard_normal((B, 32, 32, 3)) / 2], -1).astype(np.float32)).cuda()
  x = fn(0, model, B, u=u)[0]
else:
  from gddim_b200.blur import sampling, sde_lib
  cfg = configs.blur_ddpm_deep_cifar10(1.0); cfg.model.nf, cfg.model.num_res_blocks = 64, 1
  model = net.ScoreNet(cfg, cld=False); model.init_params(seed=1, nondegenerate=True)
  fn = sampling.get_order0_sampler(sde_lib.from_config(cfg), model, (32, 32, 3), 2, nfe, inv)
  fn.core.use_graph = False
  y = torch.from_numpy(rng.standard_normal((B, 32, 32, 3)).astype(np.float32)).cuda()
  x = fn(0, model, B, u=y)[0]
  from gddim_b200.blur import blur as gblur
  gblur.batch_img_dct(y)
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(x).all()))
