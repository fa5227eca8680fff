"""4x4 / 8x8-level conv GEMMs at batch 256: block_n / tile-shape variants, timed inside a CUDA graph (20 launches)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from gddim_b200 import ops
B = 256
def run(H, cin, cout, bn, ms=0, cg=0, iters=20):
  a = torch.randn(B, H, H, cin, device="cuda").half()
  k = np.random.default_rng(0).standard_normal((3, 3, cin, cout)).astype(np.float32) * 0.02
  w = ops.pack_conv_weight(k)
  r = torch.randn(B, H, H, cout, device="cuda")
  bias = torch.randn(cout, device="cuda")
  f = lambda: ops.conv_gemm(a, w, cout, bias=bias, residual=r, scale=0.7, force_block_n=bn, force_m_sub=ms, force_cta_pairs=cg)
  for _ in range(3): f()
  torch.cuda.synchronize()
  g = torch.cuda.CUDAGraph()
  with torch.cuda.graph(g):
    for _ in range(iters): f()
  g.replay(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(5): g.replay()
  e1.record(); torch.cuda.synchronize()
  ms_ = e0.elapsed_time(e1) / (5 * iters)
  fl = 2.0 * B * H * H * cout * 9 * cin
  return ms_ * 1e3, fl / ms_ * 1e-9
for H, cin in [(4, 256), (4, 512), (8, 256), (8, 512)]:
  for bn, ms, cg in [(64, 0, 0), (128, 0, 0), (256, 0, 0), (64, 2, 0), (128, 2, 0), (256, 0, 2)]:
    try:
      us, tf = run(H, cin, 256, bn, ms, cg)
      print(f"SN H={H} K={9*cin} bn={bn} ms={ms} cg={cg}: {us:7.1f} us {tf:7.1f} TF/s", flush=True)
    except Exception as e:
      print(f"SN H={H} K={9*cin} bn={bn} ms={ms} cg={cg}: failed {str(e)[:80]}", flush=True)
