"""Stand-in for cld_jax/sampling.py: sampler factory and the gDDIM / DEIS samplers on CLD.

`get_sampling_fn(config, sde, model, shape, inverse_scaler)` (sampling.py:41-154) returns
`psampler(prng, pstate, batch_size, u=None) -> (xs, vs, nfe)` with a leading device axis (n_dev = 1: one
process drives one GPU; multi-GPU runs launch one process per GPU and shard the batch axis).
The whole loop (network evaluations + multistep updates + denoising step) runs inside libgddim_b200.so.

All nine method names of sampling.py:57-151 dispatch: 'deis' (204-253), 'order0' (156-202, incl. is_em), 'hybdeis'
(255-269), 'sdeis' (380-427, on LambdaSDE), 'ldeis' (497-540), 'mldeis' (272-378), 'em' (624-669), 'sscs' (542-622) as
step programs executed by the library; 'ode' (432-495) drives gddim_unet_forward from scipy's RK45 on the host like the
reference (only the network evaluation is native there: the 2x2 algebra runs as torch ops with one host round trip per
RK45 stage).  Unknown names raise a bare RuntimeError exactly like sampling.py:152-153.
"""
import ctypes as C

import numpy as np

from .. import _lib
from .. import net as _net


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
  if name == "ldeis":
    return get_L_deis_sampler(sde=sde, model=model, data_shape=data_shape, nfe=config.sampling.nfe,
                              inverse_scaler=inverse_scaler, deis_order=config.sampling.deis_order,
                              ts_order=config.sampling.ts_order, denoising=config.sampling.noise_removal, is_p=True)
  if name == "sscs":
    return get_sscs_sampler(sde=sde, model=model, data_shape=data_shape, nfe=config.sampling.nfe,
                            inverse_scaler=inverse_scaler, ts_order=config.sampling.ts_order,
                            denoising=config.sampling.noise_removal, is_p=True)
  if name == "em":
    return get_em_sampler(sde=sde, model=model, data_shape=data_shape, nfe=config.sampling.nfe,
                          inverse_scaler=inverse_scaler, lambda_coef=config.sampling.lambda_coef,
                          ts_order=config.sampling.ts_order, denoising=config.sampling.noise_removal, is_p=True)
  if name == "mldeis":
    return get_mldeis_sampler(sde=sde, model=model, data_shape=data_shape, nfe=config.sampling.nfe,
                              inverse_scaler=inverse_scaler, deis_order=config.sampling.deis_order,
                              ts_order=config.sampling.ts_order, denoising=config.sampling.noise_removal, is_p=True)
  if name == "ode":
    return get_ode_sampler(sde=sde, model=model, data_shape=data_shape, inverse_scaler=inverse_scaler,
                           denoising=config.sampling.noise_removal, atol=config.sampling.atol,
                           rtol=config.sampling.rtol, method=config.sampling.ode_method, is_p=True)
  if name == "hybdeis":
    return get_hyd_deis_sampler(sde=sde, model=model, data_shape=data_shape, nfe=config.sampling.nfe,
                                inverse_scaler=inverse_scaler, deis_order=config.sampling.deis_order,
                                noise_nfe_ratio=config.sampling.noise_nfe_ratio,
                                img_t_ratio=config.sampling.img_t_ratio, ts_order=config.sampling.ts_order,
                                denoising=config.sampling.noise_removal, is_p=True)
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
    self._h, self._gen, self._net = None, None, None

  def _destroy(self):
    if self._h is not None:
      _lib.lib().gddim_sampler_destroy(self._h)
      self._h = None
    self._gen = None

  def _current(self, net, batch):
    """The C sampler is tied to ONE network context (weights, workspace, captured graphs): it is reused only while
    that context lives (ScoreNet.generation counts re-creations -- a freed-and-reallocated ctx can have the same
    address) and the library still reports it attached."""
    ctx = net.ensure(batch)
    if self._h is not None and self._net is net and self._gen == net.generation and \
        _lib.lib().gddim_sampler_alive(self._h):
      return ctx, True
    self._destroy()
    return ctx, False

  def _adopt(self, h, net):
    self._h, self._gen, self._net = h, net.generation, net
    net.register_sampler(self)
    _lib.check(_lib.lib().gddim_sampler_set_seed(h, int(self.seed)), "gddim_sampler_set_seed")
    return h

  def set_seed(self, seed):
    """Key of the internally drawn noise for the following calls (no rebuild; cf. gddim_sampler_set_seed)."""
    self.seed = int(seed) % (2 ** 64)
    if self._h is not None and _lib.lib().gddim_sampler_alive(self._h):
      _lib.check(_lib.lib().gddim_sampler_set_seed(self._h, self.seed), "gddim_sampler_set_seed")

  def __del__(self):
    try:
      self._destroy()
    except Exception:
      pass

  def handle(self, net, batch):
    ctx, ok = self._current(net, batch)
    if ok:
      return self._h
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
    return self._adopt(h, net)

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
      n_noise = getattr(self, "n_noise", n_steps)
      if tuple(nz_keep.shape) != (n_noise,) + shape:
        raise ValueError(f"noise has shape {tuple(nz_keep.shape)}, expected {(n_noise,) + shape}")
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


def _seed_of(rng, rank=None):
  """64-bit Philox key from whatever the caller passes as `rng` (jax PRNGKey uint32[2], int, None), with the process
  rank folded in: under one-process-per-GPU every rank must draw different noise, like the per-device keys the
  reference hands to pmap(sampler) (run_lib.py:718-722)."""
  if rng is None:
    base = 0
  else:
    a = np.asarray(rng).astype(np.uint64).ravel()
    base = 0
    for w in a:                                    # order-sensitive mix (the old sum mapped (0,7) and (7,0) together)
      base = (base * 0x9E3779B97F4A7C15 + int(w) + 0x7F4A7C15) % (2 ** 64)
  if rank is None:
    import os
    rank = int(os.environ.get("RANK", "0"))
  return (base ^ ((rank * 0xD1B54A32D192ED03) % (2 ** 64))) % (2 ** 64)


def _wrap(core, sde, data_shape, is_p):
  stochastic = core.kind == _lib.CLD_SDEIS        # sdeis, and the em / sscs programs with a noise term

  def sampler(rng, state, batch_size, u=None, trace=False, noise=None):
    """sampling.py:212-230 (non-pmapped): u (B,H,W,C,2) -> (x, v, nfe)."""
    if u is None:
      from . import sde_lib as _sl
      prior_rng = rng
      if _sl._is_jax_key(rng):                         # rng, step_rng = random.split(rng)   (sampling.py:213)
        from .. import jax_random
        prior_rng = jax_random.split(rng)[1]
      u = sde.prior_sampling(prior_rng, (batch_size,) + tuple(data_shape))
    if stochastic and noise is None:
      core.set_seed(_seed_of(rng))                     # a new key = a new noise field; no rebuild
    return core.run(state, batch_size, u, trace=trace, noise=noise)

  def psampler(prng, pstate, batch_size, u=None):
    """sampling.py:232-237: leading axis = local devices driven by this process (1).  `prng` carries one key per local
    device ([1, 2] uint32); key 0 draws the prior when u is None and -- as in pmap(sampler)(prng, ...) -- seeds the
    noise this device injects (stochastic samplers)."""
    import torch
    rng = prng
    if isinstance(prng, np.ndarray) and prng.dtype == np.uint32 and prng.ndim == 2 and prng.shape[1] == 2:
      rng = prng[0]                                     # flax.jax_utils.unreplicate(prng)   (sampling.py:233)
    if u is None:
      u = sde.prior_sampling(rng, (1, batch_size) + tuple(data_shape))
    if u.shape[0] != 1:
      raise ValueError("this process drives one GPU: the leading device axis of u must be 1 "
                       "(launch one process per GPU and shard the batch, see bench.py)")
    if stochastic:
      core.set_seed(_seed_of(rng))
    x, v, nfe = core.run(pstate, batch_size, u[0])
    if torch.is_tensor(x):
      return x[None], v[None], nfe
    return x[None], v[None], nfe

  fn = psampler if is_p else sampler
  fn.core = core
  return fn


def get_order0_sampler(sde, model, data_shape, nfe, inverse_scaler, is_em=False, denoising=False, is_p=False):
  """sampling.py:156-202 (is_em=True uses prepare_naive_coef, sde_lib.py:276-287)."""
  if is_em:                                   # sampling.py:171-172: Euler coefficients (sde_lib.py:276-287)
    num_step = nfe - 1 if denoising else nfe
    rev_ts = np.asarray(get_rev_ts(sde, 2, num_step), np.float64)
    steps = []
    for i in range(num_step):
      cur_t, dt = rev_ts[i], rev_ts[i + 1] - rev_ts[i]
      G, R = sde._G64(cur_t), sde._R64([cur_t])[0]
      steps.append(_step(t=cur_t, A=np.eye(2) + sde._F64(cur_t) * dt, Cs=[0.5 * (G @ G) @ np.linalg.inv(R).T * dt],
                         M=_mix_matrix(sde, cur_t), trace=1))
    if denoising:
      steps.append(_denoise_step(sde))
    core = _ProgramSampler(sde, model, data_shape, nfe, inverse_scaler, denoising, is_p, steps, 1)
    return _wrap(core, sde, data_shape, is_p)
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


# ---- samplers expressed as explicit step programs (gddim_sampler_create_program) ---------------------------------
class _ProgramSampler(_Sampler):
  """Every remaining CLD sampler of the reference is a sequence of affine steps u <- A u + sum_j C_j eps_j + F z;
  the tables are assembled here (fp64 numpy on top of the library's host functions) and executed by the library."""

  def __init__(self, sde, model, data_shape, nfe, inverse_scaler, denoising, is_p, steps, history):
    super().__init__(_lib.CLD_PROGRAM, sde, model, data_shape, nfe, inverse_scaler, 0, 2, denoising, is_p)
    self.steps, self.history = steps, int(history)
    self.kind = _lib.CLD_SDEIS if any(any(st.F) for st in steps) else _lib.CLD_PROGRAM   # noise= allowed iff stochastic
    self.n_noise = sum(1 for st in steps if any(st.F))

  def handle(self, net, batch):
    ctx, ok = self._current(net, batch)
    if ok:
      return self._h
    cfg = _lib.SamplerCfg(kind=_lib.CLD_PROGRAM, nfe=self.nfe, deis_order=0, ts_order=2, denoising=int(self.denoising),
                          mixed_score=int(bool(self.sde.mixed_score)), use_graph=int(self.use_graph),
                          x_mul=self.mul if self.affine else 1.0, x_add=self.add if self.affine else 0.0,
                          lambda_coef=0.0, sdeis_use_order0=0, seed=int(self.seed))
    arr = (_lib.Step * len(self.steps))(*self.steps)
    h = C.c_void_p()
    _lib.check(_lib.lib().gddim_sampler_create_program(ctx, C.byref(cfg), arr, len(self.steps), self.history, C.byref(h)),
               "gddim_sampler_create_program")
    return self._adopt(h, net)


def _step(t=-1.0, A=None, Cs=(), F=None, M=None, first_eps=0, trace=0, P=None):
  st = _lib.Step()
  st.has_P = int(P is not None)
  for k in range(4):
    st.P[k] = float(np.eye(2).ravel()[k]) if P is None else float(np.asarray(P, np.float64).ravel()[k])
  st.t, st.n_eps, st.first_eps, st.trace = float(t), len(Cs), int(first_eps), int(trace)
  A = np.eye(2) if A is None else np.asarray(A, np.float64)
  for k in range(4):
    st.A[k] = float(A.ravel()[k])
    st.F[k] = 0.0 if F is None else float(np.asarray(F, np.float64).ravel()[k])
    st.M[k] = 0.0 if M is None else float(np.asarray(M, np.float64).ravel()[k])
  for j, c in enumerate(Cs):
    for k in range(4):
      st.C[j][k] = float(np.asarray(c, np.float64).ravel()[k])
  return st


def _base(sde):
  return getattr(sde, "sde", sde)


def _mix_matrix(sde, t):
  """models/utils.py:174-176: eps += R(t)^-1 [0, v]."""
  if not sde.mixed_score:
    return None
  ri = np.linalg.inv(_base(sde)._R64([t])[0])
  return np.array([[0.0, ri[0, 1]], [0.0, ri[1, 1]]])


def _denoise_step(sde):
  """get_denoising_step (sampling.py:30-39) at t = dt = sampling_eps as one affine step."""
  b = _base(sde)
  t = b.sampling_eps
  F, G, R = b._F64(t), b._G64(t), b._R64([t])[0]
  return _step(t=t, A=np.eye(2) - t * F, Cs=[-t * (G @ G) @ np.linalg.inv(R).T], M=_mix_matrix(sde, t))


def get_L_deis_sampler(sde, model, data_shape, nfe, inverse_scaler, deis_order, ts_order=2, denoising=False,
                       is_p=False):
  """sampling.py:536-540 -> _impl_Ldeis_sampler (497-534): DEIS in the L_t parameterisation; the network output is
  mapped by L^T R^-T (sde_lib.py:493-499) before it enters the history, which is folded into the coefficients here.
  (The reference's denoising branch calls LSDE.s_F, which does not exist; here it uses the base SDE's step.)"""
  from . import sde_lib as _sl
  num_step = nfe - 1 if denoising else nfe
  rev_ts = np.asarray(get_rev_ts(sde, ts_order, num_step), np.float64)
  lsde = _sl.LSDE(sde)
  coef = np.asarray(lsde.get_deis_coef(deis_order, rev_ts), np.float64)
  conv = [lsde._epsR2epsL_matrix64(t) for t in rev_ts[:-1]]
  steps = []
  for i in range(num_step):
    r = min(i, deis_order)
    steps.append(_step(t=rev_ts[i], A=coef[i, 0], Cs=[coef[i, 1 + j] @ conv[i - j] for j in range(r + 1)],
                       M=_mix_matrix(sde, rev_ts[i]), trace=1))
  if denoising:
    steps.append(_denoise_step(sde))
  core = _ProgramSampler(sde, model, data_shape, nfe, inverse_scaler, denoising, is_p, steps, deis_order + 1)
  return _wrap(core, sde, data_shape, is_p)


def get_em_sampler(sde, model, data_shape, nfe, inverse_scaler, lambda_coef=0, ts_order=2, denoising=False, is_p=False):
  """sampling.py:624-669: Euler-Maruyama on the lambda-family reverse SDE (score = -R^-T eps, sde_lib.py:246-253)."""
  num_step = nfe - 1 if denoising else nfe
  rev_ts = np.asarray(get_rev_ts(sde, ts_order, num_step), np.float64)
  steps = []
  for i in range(num_step):
    cur_t, dt = rev_ts[i], rev_ts[i + 1] - rev_ts[i]
    F, G, R = sde._F64(cur_t), sde._G64(cur_t), sde._R64([cur_t])[0]
    steps.append(_step(t=cur_t, A=np.eye(2) + F * dt, Cs=[(1.0 + lambda_coef) / 2.0 * dt * (G @ G.T) @ np.linalg.inv(R).T],
                       F=(lambda_coef * np.sqrt(abs(dt)) * G) if lambda_coef != 0 else None,
                       M=_mix_matrix(sde, cur_t), trace=1))
  if denoising:
    steps.append(_denoise_step(sde))
  core = _ProgramSampler(sde, model, data_shape, nfe, inverse_scaler, denoising, is_p, steps, 1)
  return _wrap(core, sde, data_shape, is_p)


def _sscs_ou(sde, s_t, s_t_next):
  """get_sscs_ou_fn (sampling.py:542-566): mean matrix and covariance of the analytic OU half step."""
  bi = -1 * (sde.beta_int(1 - s_t_next) - sde.beta_int(1 - s_t))
  Gm = sde.Gamma
  mean = np.array([[1 + 2 * bi / Gm, -4 * bi / Gm / Gm], [bi, 1 - 2 * bi / Gm]]) * np.exp(-2.0 * bi / Gm)
  cov_xx = np.exp(4 * bi / Gm) - 1 - 4 * bi / Gm - 8 * bi ** 2 / Gm / Gm
  cov_xv = -4 * bi ** 2 / Gm
  cov_vv = (Gm / 2) ** 2 * (np.exp(4 * bi / Gm) - 1) + bi * Gm - 2 * bi ** 2
  return mean, np.array([[cov_xx, cov_xv], [cov_xv, cov_vv]]) * np.exp(-4 * bi / Gm)


def get_sscs_sampler(sde, model, data_shape, nfe, inverse_scaler, ts_order=2, denoising=False, is_p=False):
  """sampling.py:568-622: symmetric splitting (OU half step, score kick on v, OU half step) per step."""
  from . import sde_lib as _sl
  num_step = nfe - 1 if denoising else nfe
  rev_ts = np.asarray(get_rev_ts(sde, ts_order, num_step), np.float64)
  ts = 1 - rev_ts
  steps = []
  for i in range(num_step):
    cur_t, next_t = ts[i], ts[i + 1]
    mid = (cur_t + next_t) / 2.0
    m, c = _sscs_ou(sde, cur_t, mid)
    steps.append(_step(A=m, F=_sl.mvn_factor_svd(c)))
    te = sde.T - cur_t                                            # evaluation time of the score (sampling.py:572)
    rit = np.linalg.inv(sde._R64([te])[0]).T
    k = 2 * sde.beta(cur_t) * sde.Gamma * (next_t - cur_t)
    steps.append(_step(t=te, A=np.array([[1.0, 0.0], [0.0, 1.0 + k * sde.m_inv]]),
                       Cs=[k * np.array([[0.0, 0.0], [-rit[1, 0], -rit[1, 1]]])], M=_mix_matrix(sde, te)))
    m, c = _sscs_ou(sde, mid, next_t)
    steps.append(_step(A=m, F=_sl.mvn_factor_svd(c), trace=1))
  if denoising:
    steps.append(_denoise_step(sde))
  core = _ProgramSampler(sde, model, data_shape, nfe, inverse_scaler, denoising, is_p, steps, 1)
  return _wrap(core, sde, data_shape, is_p)


def get_ode_sampler(sde, model, data_shape, inverse_scaler, denoising=False, rtol=1e-5, atol=1e-5, method="RK45",
                    is_p=False):
  """sampling.py:432-495: probability-flow ODE integrated by scipy's black-box solver on the host; every drift
  evaluation is one network evaluation on the GPU (gddim_unet_forward) followed by 2x2 algebra in torch.
  Like the reference, the pmapped variant skips the denoising step (sampling.py:489) and takes the *global* batch."""
  import torch
  from scipy import integrate

  def _run(pstate, batch_size, u, apply_denoise):
    _lib.require_cuda("ode sampler")
    net = _net.resolve_net(model, pstate, cld=True)
    is_np = not torch.is_tensor(u)
    ud = torch.as_tensor(np.ascontiguousarray(u, dtype=np.float32)).cuda() if is_np else u.float().cuda()
    d_shape = tuple(ud.shape)
    Cc = d_shape[-2]
    mats = {}

    def eps_of(x, t):
      net_in = torch.cat([x[..., 0], x[..., 1]], dim=-1).contiguous()
      out = net.forward(net_in, float(t))
      eps = torch.stack([out[..., :Cc], out[..., Cc:]], dim=-1)
      if sde.mixed_score:
        inv_r = torch.as_tensor(np.linalg.inv(sde._R64([t])[0]), dtype=torch.float32, device=x.device)
        x0 = x.clone()
        x0[..., 0] = 0.0
        eps = eps + torch.einsum("ij,...j->...i", inv_r, x0)
      return eps

    def drift(x, t):
      F, G, R = sde._F64(t), sde._G64(t), sde._R64([t])[0]
      f = torch.as_tensor(F, dtype=torch.float32, device=x.device)
      # grad = F x - 0.5 G G score, score = -R^-T eps  =>  F x + 0.5 G G R^-T eps
      c = torch.as_tensor(0.5 * (G @ G) @ np.linalg.inv(R).T, dtype=torch.float32, device=x.device)
      return torch.einsum("ij,...j->...i", f, x) + torch.einsum("ij,...j->...i", c, eps_of(x, t))

    def ode_func(t, xf):
      x = torch.as_tensor(xf.reshape(d_shape), dtype=torch.float32).cuda()
      return drift(x, t).cpu().numpy().reshape(-1).astype(np.float64)

    sol = integrate.solve_ivp(ode_func, (sde.T, sde.sampling_eps), ud.cpu().numpy().reshape(-1).astype(np.float64),
                              rtol=rtol, atol=atol, method=method)
    nfe = sol.nfev
    uo = torch.as_tensor(sol.y[:, -1].reshape(d_shape), dtype=torch.float32).cuda()
    if apply_denoise and denoising:
      t = sde.sampling_eps
      F, G, R = sde._F64(t), sde._G64(t), sde._R64([t])[0]
      a = torch.as_tensor(np.eye(2) - t * F, dtype=torch.float32, device=uo.device)
      c = torch.as_tensor(-t * (G @ G) @ np.linalg.inv(R).T, dtype=torch.float32, device=uo.device)
      uo = torch.einsum("ij,...j->...i", a, uo) + torch.einsum("ij,...j->...i", c, eps_of(uo, t))
    x, v = uo[..., 0], uo[..., 1]
    x = inverse_scaler(x) if inverse_scaler is not None else x
    if is_np:
      return np.asarray(x.cpu()), np.asarray(v.cpu()), nfe
    return x, v, nfe

  def sampler(rng, state, batch_size, u=None):
    if u is None:
      u = sde.prior_sampling(rng, (batch_size,) + tuple(data_shape))
    return _run(state, batch_size, u, True)

  def psampler(prng, pstate, batch_size, u=None):
    if u is None:
      u = sde.prior_sampling(prng, (1, batch_size) + tuple(data_shape))
    if u.shape[0] != 1:
      raise ValueError("this process drives one GPU: the leading device axis of u must be 1")
    x, v, nfe = _run(pstate, batch_size, u[0], False)
    return x[None], v[None], nfe

  return psampler if is_p else sampler


def _psi1(sde, t, inverse=False):
  out = np.empty((2, 2))
  _lib.check(_lib.lib().gddim_cld_psi1(sde._h, float(t), int(inverse), out.ctypes.data))
  return out


def get_mldeis_sampler(sde, model, data_shape, nfe, inverse_scaler, deis_order, ts_order=2, denoising=False, is_p=False):
  """sampling.py:328-378 on MLCLD (272-325): DEIS in the frame rotated by psi1 = expm(int F_1).  The state is
  y = psi1(T)^-1 u; the network sees psi1(t_i) y; the result is mapped back with psi1(sampling_eps / 2).  As in the
  reference, the denoising step is applied to the rotated variable as if it were u (sampling.py:366)."""
  if sde.beta_1 != 0:
    raise AssertionError("MLCLD requires beta_1 == 0 (sampling.py:289)")
  num_step = nfe - 1 if denoising else nfe
  rev_ts = np.asarray(get_rev_ts(sde, ts_order, num_step), np.float64)
  coef = np.empty((num_step, deis_order + 3, 2, 2))
  _lib.check(_lib.lib().gddim_cld_mldeis_coef(sde._h, int(deis_order), rev_ts.ctypes.data, rev_ts.size, coef.ctypes.data),
             "gddim_cld_mldeis_coef")
  steps = [_step(A=_psi1(sde, sde.T, inverse=True))]                              # x2y at T
  for i in range(num_step):
    r = min(i, deis_order)
    P = _psi1(sde, rev_ts[i])
    M = _mix_matrix(sde, rev_ts[i])
    steps.append(_step(t=rev_ts[i], A=coef[i, 0], Cs=[coef[i, 1 + j] for j in range(r + 1)], P=P,
                       M=None if M is None else M @ P, trace=1))
  if denoising:
    steps.append(_denoise_step(sde))
  steps.append(_step(A=_psi1(sde, sde.sampling_eps / 2)))                          # y2x at sampling_eps / 2
  core = _ProgramSampler(sde, model, data_shape, nfe, inverse_scaler, denoising, is_p, steps, deis_order + 1)
  return _wrap(core, sde, data_shape, is_p)
