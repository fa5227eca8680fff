"""EXPERIMENTAL normalise-on-load convolution (GDDIM_XF=1, csrc/conv_xf.cu): first thing to run on a GPU box.
Runs one forward of the deep NCSN++ (batch 8 and 256) with and without the switch in separate processes (the switch is
read once per process), compares the outputs (only the fp32 accumulation order of the affected convolutions differs:
expect rel. L2 ~1e-6) and reports the per-evaluation time.  usage: python tools/exp/xf_check.py"""
import os, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
CHILD = r'''
import os, sys, time
import numpy as np, torch
sys.path.insert(0, sys.argv[1])
from gddim_b200 import configs, net
B = int(sys.argv[2])
m = net.ScoreNet(configs.cld_accr_dcifar10(), cld=True)
m.init_params(seed=1234, nondegenerate=True)
x = torch.from_numpy(np.random.default_rng(0).standard_normal((B, 32, 32, 6)).astype(np.float32)).cuda()
for _ in range(2): y = m.forward(x, 0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): y = m.forward(x, 0.5)
e1.record(); torch.cuda.synchronize()
print("ms_per_forward", e0.elapsed_time(e1) / 5, flush=True)
np.save(sys.argv[3], y.cpu().numpy())
'''

def run(xf, B):
  out = tempfile.mktemp(suffix=".npy")
  env = dict(os.environ, GDDIM_XF="1" if xf else "0")
  r = subprocess.run([sys.executable, "-c", CHILD, ROOT, str(B), out], env=env, capture_output=True, text=True, timeout=150)
  print(f"--- GDDIM_XF={int(xf)} batch {B}: rc={r.returncode}\n{r.stdout[-400:]}{r.stderr[-800:]}")
  return np.load(out) if r.returncode == 0 and os.path.exists(out) else None

for B in (8, 256):
  a, b = run(False, B), run(True, B)
  if a is None or b is None:
    print(f"batch {B}: FAILED to run"); continue
  print(f"batch {B}: rel_l2(xf, plain) = {np.linalg.norm(a - b) / np.linalg.norm(a):.3e}  finite={np.isfinite(b).all()}")
