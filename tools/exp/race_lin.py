"""One small plain convolution for compute-sanitizer racecheck.  usage: race_lin.py B H C"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from gddim_b200 import ops
B, H, C = (int(v) for v in sys.argv[1:4])
a = torch.randn(B, H, H, C, device="cuda").half()
w = ops.pack_conv_weight(np.random.default_rng(0).standard_normal((3, 3, C, C)).astype(np.float32) * 0.02)
bias = torch.randn(C, device="cuda")
o32, o16 = ops.conv_gemm(a, w, C, bias=bias)
torch.cuda.synchronize()
print("ok", float(o32.abs().max()))
