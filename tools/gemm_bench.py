"""Times individual conv/GEMM shapes of the deep net at batch 256 through the operator-level ABI (CUDA events)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from gddim_b200 import ops
B = 256
SHAPES = [  # H, Cin(A0), Cout, two_seg_Cin1, residual, stats-less
    (32, 128, 128, 0, False), (32, 128, 128, 0, True), (32, 128, 128, 256, False), (16, 256, 256, 0, True),
    (16, 256, 256, 512, False), (8, 256, 256, 0, True), (4, 256, 256, 0, True)]
def run(H, cin, cout, c1, res, iters=20):
  a = torch.randn(B, H, H, cin, device="cuda").half()
  k = np.random.default_rng(0).standard_normal((3, 3, cin, cout)).astype(np.float32) * 0.02
  k1 = np.random.default_rng(1).standard_normal((1, 1, c1, cout)).astype(np.float32) * 0.02 if c1 else None
  w = ops.pack_conv_weight(k, k1)
  a1 = torch.randn(B, H, H, c1, device="cuda").half() if c1 else None
  r = torch.randn(B, H, H, cout, device="cuda") if res else None
  bias = torch.randn(cout, device="cuda")
  for _ in range(3):
    ops.conv_gemm(a, w, cout, a1=a1, bias=bias, residual=r, scale=0.7)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    ops.conv_gemm(a, w, cout, a1=a1, bias=bias, residual=r, scale=0.7)
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / iters
  fl = 2.0 * B * H * H * cout * (9 * cin + c1)
  return ms, fl / ms * 1e-9
print("dbg", os.environ.get("GDDIM_GEMM_DBG", "0"))
for s in SHAPES:
  ms, tf = run(*s)
  print(f"H={s[0]:2d} K={9*s[1]+s[3]:5d} N={s[2]} res={int(s[4])}: {ms*1e3:7.1f} us  {tf:7.1f} TFLOP/s")
