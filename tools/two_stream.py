"""Experiment: two half-batch samplers on two CUDA streams (GroupNorm of one overlapping the GEMMs of the other)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from gddim_b200 import configs, net
from gddim_b200.cld import sampling, sde_lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256  # total images over all streams
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = configs.cld_accr_dcifar10(); cfg.sampling.nfe, cfg.sampling.deis_order = 50, 2
sde = sde_lib.from_config(cfg)
inv = lambda x: (x + 1.) / 2.
models, cores, us, streams = [], [], [], []
p = None
for i in range(NS):
  m = net.ScoreNet(cfg, cld=True)
  p = m.init_params(seed=1234, nondegenerate=True) if p is None else (m.set_params(p) or p)
  fn = sampling.get_deis_sampler(sde, m, (32, 32, 3), 50, inv, 2, ts_order=2, denoising=True)
  models.append(m); cores.append(fn.core)
  us.append(torch.randn(B // NS, 32, 32, 3, 2, device="cuda"))
  streams.append(torch.cuda.Stream())
def step():
  outs = []
  for i in range(NS):
    with torch.cuda.stream(streams[i]):
      outs.append(cores[i].run(models[i], B // NS, us[i]))
  return outs
for _ in range(3):
  step()
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 3
for _ in range(K):
  step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / K
print(f"B={B} streams={NS}: {dt*1e3:.1f} ms/step  {B/dt:.1f} img/s")
