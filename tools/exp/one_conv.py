"""One conv launch (batch 256) for ncu: usage one_conv.py H cin cout mode   (mode: 0 linear, 1 GroupNorm epilogue, 2 dual)"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from gddim_b200 import ops
H, cin, cout, mode = (int(v) for v in sys.argv[1:5])
B = 256
a = torch.randn(B, H, H, cin, device="cuda").half()
w = ops.pack_conv_weight(np.random.default_rng(0).standard_normal((3, 3, cin, cout)).astype(np.float32) * 0.02)
bias = torch.randn(cout, device="cuda")
res = torch.randn(B, H, H, cout, device="cuda") if mode == 2 else None
gamma, beta = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
for _ in range(3):
  if mode == 0: ops.conv_gemm(a, w, cout, bias=bias)
  else: ops.conv_gemm(a, w, cout, bias=bias, residual=res, gn=(gamma, beta, 32, True, 1e-6, mode == 2))
torch.cuda.synchronize()
