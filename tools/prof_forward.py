"""Two eager forwards of the deep NCSN++ at batch 256 (the workload of BASELINE config 2), for ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from gddim_b200 import configs, net
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = net.ScoreNet(configs.cld_accr_dcifar10(), cld=True)
model.init_params(seed=1234, nondegenerate=True)
x = torch.randn(B, 32, 32, 6, device="cuda")
for _ in range(2):
  y = model.forward(x, 0.5)
torch.cuda.synchronize()
print("launches", model.launch_count(), float(y.abs().max()))
