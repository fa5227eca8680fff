"""K = 256 GEMMs of the attention block (qkv: N = 768, fp16 out; proj: N = 256, residual + fp32 out) at batch 256, 16x16:
tile-shape variants, timed inside a CUDA graph."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from gddim_b200 import ops
B, H, C = 256, 16, 256
def run(N, res, bn, ms=0, cg=0, iters=20):
  a = torch.randn(B, H, H, C, device="cuda").half()
  k = np.random.default_rng(0).standard_normal((1, 1, C, N)).astype(np.float32) * 0.05
  w = ops.pack_conv_weight(k)
  r = torch.randn(B, H, H, N, device="cuda") if res else None
  bias = torch.randn(N, device="cuda")
  f = lambda: ops.conv_gemm(a, w, N, taps0=1, bias=bias, residual=r, out_fp32=res, out_fp16=not res, force_block_n=bn,
                            force_m_sub=ms, force_cta_pairs=cg)
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
  return e0.elapsed_time(e1) / (5 * iters) * 1e3
for N, res in [(768, False), (256, True)]:
  for bn, ms, cg in [(256, 0, 1), (256, 0, 2), (128, 0, 1), (128, 2, 1), (64, 2, 1)]:
    try:
      print(f"K256 N={N} res={int(res)} bn={bn} ms={ms} cg={cg}: {run(N, res, bn, ms, cg):7.1f} us", flush=True)
    except Exception as e:
      print(f"K256 N={N} res={int(res)} bn={bn} ms={ms} cg={cg}: failed {str(e)[:100]}", flush=True)
