"""Per-op CUDA-event times of eager forwards of the deep NCSN++ (batch 256), grouped by op family and shape.
Use with the experiment toggles (GDDIM_GEMM_DBG, GDDIM_ZIGZAG, ...).  usage: python tools/op_times.py [out.csv] [forwards]"""
import collections, csv, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from gddim_b200 import configs, net
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/op_times.csv"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
model = net.ScoreNet(configs.cld_accr_dcifar10(), cld=True)
model.init_params(seed=1234, nondegenerate=True)
x = torch.randn(256, 32, 32, 6, device="cuda")
for _ in range(2):
  model.forward(x, 0.5)
torch.cuda.synchronize()
model.set_profile(True)
for _ in range(n):
  model.forward(x, 0.5)
torch.cuda.synchronize()
model.dump_profile(out)
model.set_profile(False)
rows = list(csv.DictReader(open(out)))
grp = collections.OrderedDict()
for r in rows:
  tag = r["op"].split("/")[-1] if r["kind"] != "2" or "Attn" in r["op"] else "conv"
  key = (r["kind"], tag if "Attn" in r["op"] else ("gn" if r["kind"] == "1" else tag), r["H"], r["N"], r["K"])
  g = grp.setdefault(key, [0, 0.0, 0.0])
  g[0] += 1; g[1] += float(r["ms_per_forward"]); g[2] += float(r["gflop"])
tog = {k: v for k, v in os.environ.items() if k.startswith("GDDIM_")}
tot = sum(g[1] for g in grp.values())
print(f"OPT {tog}: total {tot:.3f} ms per forward")
for k, g in sorted(grp.items(), key=lambda kv: -kv[1][1])[:24]:
  print(f"OPT   kind={k[0]} {k[1]:12s} H={k[2]:>3s} N={k[3]:>4s} K={k[4]:>5s} x{g[0]:3d}: {g[1]*1e3:8.1f} us  "
        f"{(g[2] / g[1] * 1e-3) if g[1] > 0 and g[2] > 0 else 0:7.1f} TF/s")
