"""Times conv shapes of the deep net (batch 256) with the plain linear epilogue (fp32 out + column statistics) and with the
GroupNorm-fused epilogue (epi 2), plus the separate GroupNorm pass the fusion removes.  usage: python tools/exp/gnf_bench.py"""
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
for (H, cin, cout) in [(32, 128, 128), (32, 256, 128), (16, 256, 256), (16, 512, 256), (8, 256, 256), (8, 512, 256), (4, 256, 256)]:
  a = torch.randn(B, H, H, cin, device="cuda").half()
  w = ops.pack_conv_weight(np.random.default_rng(0).standard_normal((3, 3, cin, cout)).astype(np.float32) * 0.02)
  bias = torch.randn(cout, device="cuda")
  gamma, beta = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
  x32 = torch.randn(B, H, H, cout, device="cuda")
  fl = 2.0 * B * H * H * cout * 9 * cin
  res = {}
  res["linear"] = timeit(lambda: ops.conv_gemm(a, w, cout, bias=bias))
  res["linear_cg1"] = timeit(lambda: ops.conv_gemm(a, w, cout, bias=bias, force_cta_pairs=1))
  res["linear_f16out"] = timeit(lambda: ops.conv_gemm(a, w, cout, bias=bias, out_fp32=False, out_fp16=True))
  res["gnf"] = timeit(lambda: ops.conv_gemm(a, w, cout, bias=bias, gn=(gamma, beta, 32, True)))
  res["gnf_nosilu"] = timeit(lambda: ops.conv_gemm(a, w, cout, bias=bias, gn=(gamma, beta, 32, False)))
  res["linear_res"] = timeit(lambda: ops.conv_gemm(a, w, cout, bias=bias, residual=x32, scale=0.7071))
  res["gnf_dual"] = timeit(lambda: ops.conv_gemm(a, w, cout, bias=bias, residual=x32, scale=0.7071, gn=(gamma, beta, 32, True, 1e-6, True)))
  res["gn_pass"] = timeit(lambda: ops.group_norm(x32, gamma, beta, silu=True))
  print(f"H={H:2d} {cin}->{cout} K={9*cin}: " + "  ".join(f"{k}={v:6.1f}us" for k, v in res.items()) +
        f"   linear {fl/res['linear']*1e-6:6.0f} TF/s  gnf {fl/res['gnf']*1e-6:6.0f} TF/s", flush=True)
