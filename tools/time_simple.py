"""Evaluation time of the simple_cifar10 network (nf = 32: CUDA-core GEMMs for the 32- / 96-channel layers) at batch 256,
with its per-op table: usage time_simple.py [batch] [csv]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from gddim_b200 import configs, net
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = configs.cld_simple_cifar10()
model = net.ScoreNet(cfg, cld=True); model.init_params(seed=1, nondegenerate=True)
x = torch.randn(B, 32, 32, 6, device="cuda")
for _ in range(3):
  y = model.forward(x, 0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
  y = model.forward(x, 0.5)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"simple_cifar10 nf=32 batch {B}: {ms:.3f} ms per evaluation (eager launches) = {B / ms * 1e3 / 50:.0f} img/s at 50 NFE; finite={bool(torch.isfinite(y).all())}")
if len(sys.argv) > 2:
  model.set_profile(True)
  for _ in range(3):
    model.forward(x, 0.5)
  torch.cuda.synchronize()
  model.dump_profile(sys.argv[2])
  model.set_profile(False)
  import csv, collections
  rows = list(csv.DictReader(open(sys.argv[2])))
  agg = collections.defaultdict(float)
  for r in rows:
    key = ("cuda-core gemm" if r["kind"] == "2" and r["block_n"] == "0" else {"1": "groupnorm", "2": "tcgen05 gemm"}.get(r["kind"], "other"))
    agg[key] += float(r["ms_per_forward"])
  print({k: round(v, 3) for k, v in agg.items()})
