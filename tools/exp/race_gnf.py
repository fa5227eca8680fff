"""One small GroupNorm-epilogue convolution for compute-sanitizer racecheck.  usage: race_gnf.py B H C dual"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from gddim_b200 import ops
B, H, C, dual = (int(v) for v in sys.argv[1:5])
a = torch.randn(B, H, H, C, device="cuda").half()
w = ops.pack_conv_weight(np.random.default_rng(0).standard_normal((3, 3, C, C)).astype(np.float32) * 0.02)
bias = torch.randn(C, device="cuda")
gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
res = torch.randn(B, H, H, C, device="cuda") if dual else None
o32, o16 = ops.conv_gemm(a, w, C, bias=bias, residual=res, gn=(gamma, beta, 32, True, 1e-6, bool(dual)))
torch.cuda.synchronize()
print("ok", float(o16.float().abs().max()))
