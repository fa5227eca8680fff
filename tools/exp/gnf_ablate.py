"""Epilogue ablation timing (make ABLATE=1 builds only; results are INVALID by construction): one shape, GroupNorm epilogue,
GDDIM_GEMM_DBG from the environment (4: no TMEM loads in pass 2, 5: no swish, 6: no pass 1, 7: no stores in pass 2, 8: no pass 2).
usage: GDDIM_GEMM_DBG=k python tools/exp/gnf_ablate.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from gddim_b200 import ops
B = 256
def timeit(fn, iters=20):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters * 1e3
out = []
for (H, cin, cout) in [(32, 128, 128), (32, 256, 128), (16, 256, 256), (8, 512, 256)]:
  a = torch.randn(B, H, H, cin, device="cuda").half()
  w = ops.pack_conv_weight(np.random.default_rng(0).standard_normal((3, 3, cin, cout)).astype(np.float32) * 0.02)
  bias = torch.randn(cout, device="cuda")
  gamma, beta = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
  t0 = timeit(lambda: ops.conv_gemm(a, w, cout, bias=bias, out_fp32=False, out_fp16=True))
  t1 = timeit(lambda: ops.conv_gemm(a, w, cout, bias=bias, gn=(gamma, beta, 32, True)))
  out.append(f"H{H} K{9*cin}: lin16 {t0:6.1f} gnf {t1:6.1f}")
print(f"DBG={os.environ.get('GDDIM_GEMM_DBG', '0')}  " + " | ".join(out), flush=True)
