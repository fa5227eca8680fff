"""A/B timing of one sampler configuration under environment toggles (GDDIM_PDL, GDDIM_ZIGZAG, ...).
Prints ms per sampler call, images/s and a checksum of the samples (toggles that only change scheduling must
leave the checksum bit-identical).  usage: GDDIM_PDL=0 python tools/ab.py [batch] [steps]"""
import hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from gddim_b200 import configs, net
from gddim_b200.cld import sampling, sde_lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = configs.cld_accr_dcifar10(); cfg.sampling.method, cfg.sampling.nfe, cfg.sampling.deis_order = "deis", 50, 2
model = net.ScoreNet(cfg, cld=True)
model.init_params(seed=1234, nondegenerate=True)
sde = sde_lib.from_config(cfg)
core = sampling.get_sampling_fn(cfg, sde, model, None, lambda x: (x + 1.) / 2.).core
rng = np.random.default_rng(0)
shape = (B, 32, 32, 3)
u = torch.from_numpy(np.stack([rng.standard_normal(shape), rng.standard_normal(shape) / 2.], -1).astype(np.float32)).cuda()
for _ in range(3):
  x = core.run(model, B, u)[0]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = torch.cuda.current_stream()
e0.record(st)
for _ in range(K):
  x = core.run(model, B, u)[0]
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
tog = {k: v for k, v in os.environ.items() if k.startswith("GDDIM_")}
print(f"AB {tog}: {ms:.1f} ms/step  {B / ms * 1e3:.1f} img/s  finite={bool(torch.isfinite(x).all())} "
      f"sha={hashlib.sha1(x.cpu().numpy().tobytes()).hexdigest()[:12]}", flush=True)
