"""clock64 timeline of CTA 0 for one conv launch (needs a `make ABLATE=1` build; GDDIM_CLK=1).  usage: clk_timeline.py H cin cout gnf"""
import os, sys
os.environ["GDDIM_CLK"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from gddim_b200 import ops
H, cin, cout, gnf = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
B = 256
a = torch.randn(B, H, H, cin, device="cuda").half()
w = ops.pack_conv_weight(np.random.default_rng(0).standard_normal((3, 3, cin, cout)).astype(np.float32) * 0.02)
bias = torch.randn(cout, device="cuda")
gamma, beta = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
print(f"== H={H} {cin}->{cout} gnf={gnf}", file=sys.stderr, flush=True)
for _ in range(2):
  if gnf: ops.conv_gemm(a, w, cout, bias=bias, gn=(gamma, beta, 32, gnf == 1))
  else: ops.conv_gemm(a, w, cout, bias=bias)
torch.cuda.synchronize()
