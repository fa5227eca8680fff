"""Shared builders for the parity tests: small-but-complete network configs and oracle closures."""
import numpy as np

from gddim_b200 import configs, net, params
from oracle import blur as ob
from oracle import cld as oc
from oracle import ncsnpp as on


def small_cfg(kind):
  """Same control flow as the shipped configs, narrowed to nf=64 / 1 res-block so the CPU oracle runs in
  seconds.  'deep': FIR + input pyramid + Fourier embedding (accr_dcifar10 / ddpm_deep); 'ddpmpp': naive
  resampling, positional embedding, no pyramid."""
  if kind == "cld_deep":
    c = configs.cld_accr_dcifar10()
  elif kind == "cld_ddpmpp":
    c = configs.cld_ddpmpp_cifar10()
  elif kind == "cld_mixed":
    c = configs.cld_deep_cifar10()
    c.model.mixed_score = True
    c.model.R_dt = 1e-4          # cheaper table; same code path
  elif kind == "blur_deep":
    c = configs.blur_ddpm_deep_cifar10(1.0)
  else:
    raise KeyError(kind)
  c.model.nf = 64
  c.model.num_res_blocks = 1
  return c


_cache = {}


def build(kind, nondegenerate=True, seed=1234, precise=False):
  """-> (cfg, ScoreNet with parameters, oracle net_fn(x, labels) in fp32).  precise: convolution weights as fp16 (hi, lo)
  pairs (the library's parity mode)."""
  key = (kind, nondegenerate, seed, precise)
  if key not in _cache:
    cfg = small_cfg(kind)
    cld = not kind.startswith("blur")
    model = net.ScoreNet(cfg, cld=cld, precise=precise)
    p = model.init_params(seed=seed, nondegenerate=nondegenerate)
    _cache[key] = (cfg, model, on.make_net_fn(p, cfg))
  return _cache[key]


def build_oracle_only(kind, nondegenerate=True, seed=1234):
  """-> (cfg, oracle net_fn) without touching the CUDA library's context (specs still come from its walk)."""
  cfg, _, net_fn = build(kind, nondegenerate, seed)
  return cfg, net_fn


def oracle_cld_sample(cfg, net_fn, u, nfe, order, denoising=True, method="deis", trace=None):
  sde = oc.from_config(cfg)
  eps_fn = oc.make_eps_fn(sde, net_fn)
  if method == "deis":
    return oc.deis_sampler(sde, eps_fn, u, nfe, order, ts_order=cfg.sampling.ts_order, denoising=denoising,
                           centered=cfg.data.centered, dtype=np.float32, trace=trace)
  return oc.order0_sampler(sde, eps_fn, u, nfe, denoising=denoising, centered=cfg.data.centered, dtype=np.float32)


def oracle_blur_sample(cfg, net_fn, y, nfe, trace=None):
  sde = ob.from_config(cfg)
  return ob.order0_sampler(sde, net_fn, y, nfe, ts_order=cfg.sampling.ts_order, centered=cfg.data.centered,
                           dtype=np.float32, trace=trace)


def prior_u(batch, seed=0, cld=True):
  rng = np.random.default_rng(seed)
  if cld:
    return oc.prior_sampling(rng, (batch, 32, 32, 3)).astype(np.float32)
  return rng.standard_normal((batch, 32, 32, 3)).astype(np.float32)
