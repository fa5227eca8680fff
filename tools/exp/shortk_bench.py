"""Experiment: the K = 256 GEMMs of the attention blocks (qkv N=768 fp16 out; proj N=256 + residual, fp32 out) at
batch 256, 16x16 -- time per launch under the GDDIM_GEMM_DBG ablations (1 = TMEM drain only, 2 = no stores)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from gddim_b200 import ops
B, H, C = 256, 16, 256
def t(fn, iters=20):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters * 1e3
a = torch.randn(B, H, H, C, device="cuda").half()
wq = (torch.randn(3 * C, C, device="cuda") * 0.05).half()
wp = (torch.randn(C, C, device="cuda") * 0.05).half()
bq, bp = torch.randn(3 * C, device="cuda"), torch.randn(C, device="cuda")
r = torch.randn(B, H, H, C, device="cuda")
print("dbg", os.environ.get("GDDIM_GEMM_DBG", "0"))
for bn in (256, 128):
  us = t(lambda: ops.conv_gemm(a, wq, 3 * C, taps0=1, bias=bq, out_fp32=False, out_fp16=True, force_block_n=bn))
  print(f"qkv  N=768 bn={bn}: {us:6.1f} us   ({(a.numel()*2 + B*H*H*3*C*2)/us*1e-6:.2f} TB/s algorithmic)")
  us = t(lambda: ops.conv_gemm(a, wp, C, taps0=1, bias=bp, residual=r, scale=0.7, force_block_n=bn))
  print(f"proj N=256 bn={bn}: {us:6.1f} us   ({(a.numel()*2 + 2*r.numel()*4)/us*1e-6:.2f} TB/s algorithmic)")
  us = t(lambda: ops.conv_gemm(a, wp, C, taps0=1, bias=bp, out_fp32=False, out_fp16=True, force_block_n=bn))
  print(f"o16  N=256 bn={bn}: {us:6.1f} us   ({(a.numel()*2 + r.numel()*2)/us*1e-6:.2f} TB/s algorithmic)")
