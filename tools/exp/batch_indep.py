import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from helpers import prior_u
from gddim_b200 import configs, net
from gddim_b200.cld import sampling, sde_lib
cfg = configs.cld_accr_dcifar10()
model = net.ScoreNet(cfg, cld=True); model.init_params(seed=1234, nondegenerate=True)
sde = sde_lib.from_config(cfg)
inv = lambda x: (x + 1.) / 2.
fn = sampling.get_deis_sampler(sde, model, (32, 32, 3), 6, inv, 2, ts_order=2, denoising=True)
u = prior_u(256, seed=11); ud = torch.as_tensor(u).cuda()
x, v, _ = fn(0, model, 256, u=ud)
for nb, sl in ((2, slice(100, 102)), (64, slice(0, 64)), (128, slice(64, 192)), (192, slice(0, 192))):
  x2, v2, _ = fn(0, model, nb, u=ud[sl].contiguous())
  d = (x[sl] - x2).abs().max().item(); dv = (v[sl] - v2).abs().max().item()
  print(f"batch 256 rows {sl} vs batch {nb}: max|dx| {d:.3e} max|dv| {dv:.3e} bit-identical {bool(d == 0 and dv == 0)}")
