"""Stand-in for cld_jax/sampling.py: sampler factory and the gDDIM / DEIS samplers on CLD.

`get_sampling_fn(config, sde, model, shape, inverse_scaler)` (sampling.py:41-154) returns
`psampler(prng, pstate, batch_size, u=None) -> (xs, vs, nfe)` with a leading device axis (n_dev = 1: one
process drives one GPU; multi-GPU runs launch one process per GPU and shard the batch axis).
The whole loop (network evaluations + multistep updates + denoising step) runs inside libgddim_b200.so.

Implemented: 'deis' (204-253), 'order0' (156-202, is_em=False), 'hybdeis' (255-269) and the stochastic 'sdeis'
(380-427, on LambdaSDE).  'ldeis', 'mldeis', 'ode', 'sscs', 'em' are SURVEY.md 8(f) "next" rows and raise
NotImplementedError; unknown names raise a bare RuntimeError exactly like sampling.py:152-153.
"""
import ctypes as C

import numpy as np

from .. import _lib
from .. import net as _net

_NEXT = ("ldeis", "mldeis", "ode", "sscs", "em")


def get_data_shape(config):
  """cld_jax/utils.py:172-177."""
  if "ps" in config.data.dataset:
    return (config.data.dim,)
  return (config.data.image_size, config.data.image_size, config.data.num_channels)


def get_rev_ts(sde, ts_order, num_step):
  """sampling.py:241-249 (fp64 table from the library, returned as fp32 like the reference)."""
  if float(ts_order) != int(ts_order):
    out = np.power(np.linspace(np.power(sde.T, 1.0 / ts_order), np.power(sde.sampling_eps, 1.0 / ts_order),
                               num_step + 1), ts_order)
  else:
    out = np.empty(num_step + 1)
    _lib.check(_lib.lib().gddim_rev_ts(float(sde.T), float(sde.sampling_eps), int(ts_order), int(num_step),
                                        out.ctypes.data))
  return out.astype(np.float64 if getattr(sde, "x64", False) else np.float32)


def _affine_of(inverse_scaler):
  """(mul, add, exact): detects an affine inverse scaler by probing it."""
  if inverse_scaler is None:
    return 1.0, 0.0, True
  p = np.asarray([0.0, 1.0, -1.0, 0.37], np.float64)
  q = np.asarray(inverse_scaler(p), np.float64)
  add, mul = q[0], q[1] - q[0]
  return float(mul), float(add), bool(np.allclose(q, p * mul + add, atol=1e-12))


def get_sampling_fn(config, sde, model, shape, inverse_scaler):
  del shape
  name = config.sampling.method.lower()
  data_shape = get_data_shape(config)
  if name == "order0":
    return get_order0_sampler(sde=sde, model=model, data_shape=data_shape, nfe=config.sampling.nfe,
                              inverse_scaler=inverse_scaler, is_em=config.sampling.is_em,
                              denoising=config.sampling.noise_removal, is_p=True)
  if name == "deis":
    return get_deis_sampler(sde=sde, model=model, data_shape=data_shape, nfe=config.sampling.nfe,
                            inverse_scaler=inverse_scaler, deis_order=config.sampling.deis_order,
                            ts_order=config.sampling.ts_order, denoising=config.sampling.noise_removal, is_p=True)
  if name == "sdeis":
    return get_sdeis_sampler(sde=sde, model=model, data_shape=data_shape, nfe=config.sampling.nfe,
                             inverse_scaler=inverse_scaler, deis_order=config.sampling.deis_order,
                             lambda_coef=config.sampling.lambda_coef, use_order0=config.sampling.sdeis_use_order0,
                             ts_order=config.sampling.ts_order, denoising=config.sampling.noise_removal, is_p=True)
  if name == "hybdeis":
    return get_hyd_deis_sampler(sde=sde, model=model, data_shape=data_shape, nfe=config.sampling.nfe,
                                inverse_scaler=inverse_scaler, deis_order=config.sampling.deis_order,
                                noise_nfe_ratio=config.sampling.noise_nfe_ratio,
                                img_t_ratio=config.sampling.img_t_ratio, ts_order=config.sampling.ts_order,
                                denoising=config.sampling.noise_removal, is_p=True)
  if name in _NEXT:
    raise NotImplementedError(f"sampler '{name}' is not part of the round-1 hot path (SURVEY.md 8f)")
  raise RuntimeError


class _Sampler:
  """Owns the C sampler object for one (network context, batch) pair."""

  def __init__(self, kind, sde, model, data_shape, nfe, inverse_scaler, deis_order, ts_order, denoising, is_p,
               use_graph=True, rev_ts=None, lambda_coef=0.0, use_order0=True):
    self.kind, self.sde, self.model, self.data_shape = kind, sde, model, tuple(data_shape)
    self.nfe, self.order, self.ts_order, self.denoising, self.is_p = int(nfe), int(deis_order), int(ts_order), \
        bool(denoising), is_p
    self.inverse_scaler = inverse_scaler
    self.mul, self.add, self.affine = _affine_of(inverse_scaler)
    self.use_graph = use_graph
    self.rev_ts = None if rev_ts is None else np.ascontiguousarray(np.asarray(rev_ts, np.float64))
    self.lambda_coef, self.use_order0, self.seed = float(lambda_coef), bool(use_order0), 0
    self._h, self._ctx_id, self._net = None, None, None

  def _destroy(self):
    if self._h is not None:
      _lib.lib().gddim_sampler_destroy(self._h)
      self._h = None

  def __del__(self):
    try:
      self._destroy()
    except Exception:
      pass

  def handle(self, net, batch):
    ctx = net.ensure(batch)
    if self._h is not None and self._ctx_id == ctx.value and self._net is net:
      return self._h
    self._destroy()
    cfg = _lib.SamplerCfg(kind=self.kind, nfe=self.nfe, deis_order=self.order, ts_order=self.ts_order,
                          denoising=int(self.denoising), mixed_score=int(bool(self.sde.mixed_score)),
                          use_graph=int(self.use_graph), x_mul=self.mul if self.affine else 1.0,
                          x_add=self.add if self.affine else 0.0, lambda_coef=self.lambda_coef,
                          sdeis_use_order0=int(self.use_order0), seed=int(self.seed))
    h = C.c_void_p()
    if self.rev_ts is None:
      rc = _lib.lib().gddim_sampler_create(ctx, C.byref(cfg), self.sde._h, None, C.byref(h))
    else:
      rc = _lib.lib().gddim_sampler_create_ts(ctx, C.byref(cfg), self.sde._h, None, self.rev_ts.ctypes.data,
                                               self.rev_ts.size, C.byref(h))
    _lib.check(rc, "gddim_sampler_create")
    self._h, self._ctx_id, self._net = h, ctx.value, net
    return h

  # -- introspection used by parity tests --------------------------------------------------------------
  def coef_table(self, net, batch):
    h = self.handle(net, batch)
    n = _lib.lib().gddim_sampler_coef(h, None, 0)
    out = np.empty(n, np.float32)
    _lib.lib().gddim_sampler_coef(h, out.ctypes.data, n)
    per = {_lib.CLD_DEIS: self.order + 3, _lib.CLD_ORDER0: 3, _lib.CLD_SDEIS: self.order + 4}[self.kind]
    return out.reshape(-1, per, 2, 2)

  def launch_count(self):
    return int(_lib.lib().gddim_sampler_launch_count(self._h)) if self._h is not None else 0

  def run(self, pstate, batch_size, u, trace=False, noise=None):
    import torch
    _lib.require_cuda("sampler")
    net = _net.resolve_net(self.model, pstate, cld=True)
    shape = (batch_size,) + self.data_shape + (2,)
    is_np = not torch.is_tensor(u)
    if tuple(u.shape) != shape:
      raise ValueError(f"u has shape {tuple(u.shape)}, expected {shape}")
    h = self.handle(net, batch_size)
    st = torch.cuda.current_stream().cuda_stream
    tr = None
    n_steps = _lib.lib().gddim_sampler_num_steps(h)
    if trace:
      tr = torch.empty((n_steps,) + shape, dtype=torch.float32, device="cuda")
    nz, nz_keep = None, None
    if noise is not None:                 # explicit standard normals [n_steps, B, H, W, C, 2] (parity hook for sdeis)
      if self.kind != _lib.CLD_SDEIS:
        raise ValueError("noise= is only meaningful for the sdeis sampler")
      nz_keep = (noise.detach().to(device="cuda", dtype=torch.float32) if torch.is_tensor(noise)
                 else torch.as_tensor(np.ascontiguousarray(noise, dtype=np.float32)).cuda()).contiguous()
      if tuple(nz_keep.shape) != (n_steps,) + shape:
        raise ValueError(f"noise has shape {tuple(nz_keep.shape)}, expected {(n_steps,) + shape}")
      nz = nz_keep.data_ptr()
    if is_np:
      uh = np.ascontiguousarray(u, dtype=np.float32)
      x = np.empty((batch_size,) + self.data_shape, np.float32)
      v = np.empty_like(x)
      _lib.check(_lib.lib().gddim_sample_noise(h, uh.ctypes.data, x.ctypes.data, v.ctypes.data, batch_size, 1,
                                                tr.data_ptr() if tr is not None else None, nz, st), "gddim_sample")
    else:
      ud = u.detach().to(device="cuda", dtype=torch.float32).contiguous()
      x = torch.empty((batch_size,) + self.data_shape, dtype=torch.float32, device="cuda")
      v = torch.empty_like(x)
      _lib.check(_lib.lib().gddim_sample_noise(h, ud.data_ptr(), x.data_ptr(), v.data_ptr(), batch_size, 0,
                                                tr.data_ptr() if tr is not None else None, nz, st), "gddim_sample")
    if not self.affine:
      x = self.inverse_scaler(x)
    if trace:
      return x, v, self.nfe, (tr.cpu().numpy() if is_np else tr)
    return x, v, self.nfe


def _seed_of(rng):
  if rng is None:
    return 0
  return int(np.asarray(rng).astype(np.uint64).ravel().sum() % (2 ** 63))


def _wrap(core, sde, data_shape, is_p):
  def sampler(rng, state, batch_size, u=None, trace=False, noise=None):
    """sampling.py:212-230 (non-pmapped): u (B,H,W,C,2) -> (x, v, nfe)."""
    if u is None:
      u = sde.prior_sampling(rng, (batch_size,) + tuple(data_shape))
    if core.kind == _lib.CLD_SDEIS and noise is None and _seed_of(rng) != core.seed:
      core.seed = _seed_of(rng)          # a new key re-seeds the Philox stream (sampler object is rebuilt)
      core._destroy()
    return core.run(state, batch_size, u, trace=trace, noise=noise)

  def psampler(prng, pstate, batch_size, u=None):
    """sampling.py:232-237: leading axis = local devices driven by this process (1)."""
    import torch
    if u is None:
      u = sde.prior_sampling(prng, (1, batch_size) + tuple(data_shape))
    if u.shape[0] != 1:
      raise ValueError("this process drives one GPU: the leading device axis of u must be 1 "
                       "(launch one process per GPU and shard the batch, see bench.py)")
    x, v, nfe = core.run(pstate, batch_size, u[0])
    if torch.is_tensor(x):
      return x[None], v[None], nfe
    return x[None], v[None], nfe

  fn = psampler if is_p else sampler
  fn.core = core
  return fn


def get_order0_sampler(sde, model, data_shape, nfe, inverse_scaler, is_em=False, denoising=False, is_p=False):
  """sampling.py:156-202.  is_em=True (prepare_naive_coef) is the 'em'-flavoured variant: not in round 1."""
  if is_em:
    raise NotImplementedError("order0 with is_em=True (Euler coefficients) is a SURVEY.md 8(f) 'next' row")
  core = _Sampler(_lib.CLD_ORDER0, sde, model, data_shape, nfe, inverse_scaler, 0, 2, denoising, is_p)
  return _wrap(core, sde, data_shape, is_p)


def get_deis_sampler(sde, model, data_shape, nfe, inverse_scaler, deis_order, ts_order=2, denoising=False,
                     is_p=False):
  """sampling.py:251-253 -> _impl_deis_sampler (204-239)."""
  if float(ts_order) != int(ts_order):          # non-integer exponent (jnp.power accepts floats): explicit grid
    grid = get_rev_ts(sde, ts_order, nfe - 1 if denoising else nfe)
    return _impl_deis_sampler(sde, model, data_shape, nfe, inverse_scaler, deis_order, grid, denoising, is_p)
  core = _Sampler(_lib.CLD_DEIS, sde, model, data_shape, nfe, inverse_scaler, deis_order, int(ts_order), denoising,
                  is_p)
  return _wrap(core, sde, data_shape, is_p)


def _next(name):
  def f(*a, **k):
    raise NotImplementedError(f"{name} is not part of the round-1 hot path (SURVEY.md 8f)")
  f.__name__ = name
  return f


def _impl_deis_sampler(sde, model, data_shape, nfe, inverse_scaler, deis_order, rev_ts, denoising=False, is_p=False):
  """sampling.py:204-239: the DEIS sampler on an explicit time grid rev_ts[num_step + 1]."""
  num_step = nfe - 1 if denoising else nfe
  rev_ts = np.asarray(rev_ts, np.float64)
  assert rev_ts.shape[0] == num_step + 1
  core = _Sampler(_lib.CLD_DEIS, sde, model, data_shape, nfe, inverse_scaler, deis_order, 2, denoising, is_p,
                  rev_ts=rev_ts)
  return _wrap(core, sde, data_shape, is_p)


def get_hyd_deis_sampler(sde, model, data_shape, nfe, inverse_scaler, deis_order, noise_nfe_ratio=0.3, img_t_ratio=0.3,
                         ts_order=2.0, denoising=False, is_p=False):
  """sampling.py:255-269: uniform steps from T to img_t_ratio*T, then the polynomial grid (restated literally,
  including that the second grid restarts at sde.T)."""
  num_step = nfe - 1 if denoising else nfe
  mid_t = sde.T * img_t_ratio
  noise_nfe = int(num_step * noise_nfe_ratio)
  img_nfe = num_step - noise_nfe
  noise_ts = np.linspace(sde.T, mid_t, noise_nfe, endpoint=False)
  img_ts = np.asarray(get_rev_ts(sde, ts_order, img_nfe), np.float64)
  rev_ts = np.concatenate([noise_ts, img_ts])
  assert rev_ts.shape[0] == num_step + 1
  return _impl_deis_sampler(sde, model, data_shape, nfe, inverse_scaler, deis_order, rev_ts, denoising, is_p)


def get_sdeis_sampler(sde, model, data_shape, nfe, inverse_scaler, deis_order, lambda_coef=0, use_order0=True,
                      ts_order=2, denoising=False, is_p=False):
  """sampling.py:423-427 -> _impl_sdeis_sampler (416-421) on LambdaSDE(sde, lambda_coef, use_order0): the DEIS mean
  update plus N(0, P_i) noise per (x, v) pair with P_i the conditional reverse covariance (sde_lib.py:381-399).
  The injected normals come from a Philox4x32-10 stream keyed by `rng` (JAX's threefry is not reproduced) or from
  the explicit `noise=` array of the non-pmapped sampler."""
  core = _Sampler(_lib.CLD_SDEIS, sde, model, data_shape, nfe, inverse_scaler, deis_order, int(ts_order), denoising,
                  is_p, lambda_coef=lambda_coef, use_order0=use_order0)
  return _wrap(core, sde, data_shape, is_p)
get_L_deis_sampler = _next("get_L_deis_sampler")
get_mldeis_sampler = _next("get_mldeis_sampler")
get_ode_sampler = _next("get_ode_sampler")
get_sscs_sampler = _next("get_sscs_sampler")
get_em_sampler = _next("get_em_sampler")
