"""ORACLE (test infrastructure, NOT product code) -- CLD SDE tables + DEIS sampler, numpy fp64.

Parity status: PINNED to outputs of the reference itself.  The reference ships no golden vectors and its tests
assert nothing, but its own source files run unmodified under tests/refshim (numpy stand-in for
jax/flax); tests/golden/make_ref_golden.py wrote their outputs to tests/golden/ref_cld_*.npz and
tests/test_ref_golden.py holds every function below to them in fp64: R(t)/Psi/F/G/integrand tables (RK4 and
Euler scans, beta_1 != 0), get_deis_coef orders 0-3, prepare_order0/naive_coef, LambdaSDE / LSDE / MLCLD
tables (<= 1e-7), multistep_ab_step (1e-13), and all ten sampler factories end to end on a small NCSN++
(<= 1e-7 relative L2).  The analytic known-answer identities of SURVEY.md 8(c) stay in tests/test_oracle_cld.py.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.

Follows (file:line relative to /root/reference):
  cld_jax/sde_lib.py:17-43     inv_2x2, get_interp_fn
  cld_jax/sde_lib.py:45-118    CLD.__init__, _get_s_R_fn (scan in oracle/cld_ode.c or the python loop here)
  cld_jax/sde_lib.py:182-253   s_psi, s_eps_integrand, s_F, s_G, eps2score
  cld_jax/sde_lib.py:289-319   prepare_order0_coef, get_deis_coef
  cld_jax/deis.py:19-95        DEIS Adams-Bashforth coefficient tables
  cld_jax/deis.py:141-151      multistep_ab_step
  cld_jax/sampling.py:23-39    get_denoising_step
  cld_jax/sampling.py:156-253  get_order0_sampler, _impl_deis_sampler, get_rev_ts, get_deis_sampler
  cld_jax/models/utils.py:153-176  (d g)->(g d) relayout, labels = 999 t, mixed_score
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcld_ode.so")


def build_c(force=False):
  """gcc-compile the C scan (oracle/cld_ode.c).  Called by __graft_entry__.build() and lazily here."""
  src = os.path.join(_HERE, "cld_ode.c")
  if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
  return _SO


_lib = None


def _clib():
  global _lib
  if _lib is None:
    _lib = ctypes.CDLL(build_c())
    _lib.oracle_cld_scan_R.restype = None
    _lib.oracle_cld_scan_R.argtypes = [ctypes.c_double] * 4 + [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_double, ctypes.c_int, ctypes.c_void_p]
  return _lib


def inv_2x2(m):
  """sde_lib.py:17-26 (batched over leading axes)."""
  a, b, c, d = m[..., 0, 0], m[..., 0, 1], m[..., 1, 0], m[..., 1, 1]
  coef = 1.0 / (a * d - b * c)
  out = np.empty_like(m)
  out[..., 0, 0] = d * coef
  out[..., 0, 1] = -b * coef
  out[..., 1, 0] = -c * coef
  out[..., 1, 1] = a * coef
  return out


def get_rev_ts(T, sampling_eps, ts_order, num_step):
  """sampling.py:241-249 (fp64 here; the reference evaluates it in fp32)."""
  return np.power(np.linspace(np.power(T, 1.0 / ts_order), np.power(sampling_eps, 1.0 / ts_order), num_step + 1),
                  ts_order)


class CLD:
  """cld_jax/sde_lib.py:45-319, fp64, no pickle cache."""

  def __init__(self, m_inv=4.0, beta_0=4.0, beta_1=0.0, vv_gamma=0.04, numerical_eps=1e-6,
               mixed_score=False, is_R_rk=False, R_dt=1e-5, use_c=True):
    self.mixed_score = mixed_score
    self.m_inv = float(m_inv)
    self.Gamma = 2.0 / np.sqrt(m_inv)
    self.beta_0 = float(beta_0)
    self.beta_1 = float(beta_1)
    self.R_0 = np.array([[np.sqrt(numerical_eps), 0.0], [0.0, np.sqrt(vv_gamma / m_inv + numerical_eps)]])
    self.sampling_eps = 1e-3
    self.T = 1.0
    self.is_R_rk = bool(is_R_rk)
    self.R_dt = float(R_dt)
    self._build_R_table(use_c)

  # ---- schedule -------------------------------------------------------------------------------
  def beta(self, t):
    return self.beta_0 + self.beta_1 * t

  def beta_int(self, t):
    return self.beta_0 * t + 0.5 * self.beta_1 * t ** 2

  # ---- R(t) -----------------------------------------------------------------------------------
  def _ode_rhs(self, R, t):
    """sde_lib.py:94-97."""
    F, G = self.s_F(t), self.s_G(t)
    return F @ R + 0.5 * G @ G.T @ inv_2x2(R).T

  def _scan_py(self, ts):
    """Pure-python scan (slow; used to cross-check the C scan on short grids).  sde_lib.py:98-107."""
    R = self.R_0.copy()
    out = np.empty((len(ts), 2, 2))
    dt = self.R_dt
    for k, t in enumerate(ts):
      out[k] = R
      if self.is_R_rk:                                        # deis.py:5-17
        g1 = self._ode_rhs(R, t)
        g2 = self._ode_rhs(R + g1 * dt / 2, t + dt / 2)
        g3 = self._ode_rhs(R + g2 * dt / 2, t + dt / 2)
        g4 = self._ode_rhs(R + g3 * dt, t + dt)
        R = R + dt / 6 * (g1 + 2 * g2 + 2 * g3 + g4)
      else:                                                   # "mid point integral"
        F = (self.s_F(t) + self.s_F(t + dt)) / 2.0
        G = (self.s_G(t) + self.s_G(t + dt)) / 2.0
        R = R + dt * (F @ R + 0.5 * G @ G @ np.linalg.inv(R).T)
    return out

  def _grid(self):
    n = int(1.0 / self.R_dt)                                  # sde_lib.py:109 (99999 for 1e-5!)
    return np.linspace(0, 1.0 + self.R_dt, n + 1, endpoint=False)

  def _build_R_table(self, use_c):
    ts = self._grid()
    if use_c:
      Rs = np.empty((len(ts), 2, 2))
      R0 = np.ascontiguousarray(self.R_0)
      tsc = np.ascontiguousarray(ts)
      _clib().oracle_cld_scan_R(self.m_inv, self.beta_0, self.beta_1, self.Gamma,
                                R0.ctypes.data, tsc.ctypes.data, len(tsc), self.R_dt, int(self.is_R_rk),
                                Rs.ctypes.data)
    else:
      Rs = self._scan_py(ts)
    idx = np.linspace(0, Rs.shape[0] - 1, 100_000).astype(np.int64)   # sde_lib.py:117 (dtype=int: floor)
    self._xp, self._fp = ts[idx], Rs[idx]

  def R(self, t):
    """get_interp_fn, sde_lib.py:32-43: piecewise-linear table lookup; t scalar or array -> [...,2,2]."""
    x = np.asarray(t, dtype=np.float64)
    xp, fp = self._xp, self._fp
    i = np.clip(np.searchsorted(xp, x, side="right"), 1, len(xp) - 1)
    df = fp[i] - fp[i - 1]
    dx = xp[i] - xp[i - 1]
    delta = x - xp[i - 1]
    return fp[i - 1] + (delta / dx)[..., None, None] * df

  s_R = R

  def invR(self, t):
    return inv_2x2(self.R(t))

  def cov(self, t):
    R = self.R(t)
    return R @ np.swapaxes(R, -1, -2)

  # ---- closed forms -----------------------------------------------------------------------------
  def psi(self, s, t):
    """s_psi, sde_lib.py:182-205.  Broadcasts over s, t -> [...,2,2]."""
    B = self.beta_int(np.asarray(t, dtype=np.float64)) - self.beta_int(np.asarray(s, dtype=np.float64))
    a = 2.0 * np.sqrt(self.m_inv)
    coef = np.exp(-a * B / 2)
    out = np.empty(np.shape(B) + (2, 2))
    out[..., 0, 0] = (1 + a * B / 2) * coef
    out[..., 0, 1] = 0.25 * a * a * B * coef
    out[..., 1, 0] = -B * coef
    out[..., 1, 1] = (1 - a * B / 2) * coef
    return out

  def s_F(self, t):
    b = self.beta(t)
    return np.array([[0.0, b * self.m_inv], [-b, -self.Gamma * b * self.m_inv]])

  def s_G(self, t):
    b = self.beta(t)
    return np.array([[0.0, 0.0], [0.0, np.sqrt(2 * self.Gamma * b)]])

  def eps_integrand(self, t):
    """s_eps_integrand, sde_lib.py:207-212, vectorised: 0.5 G G R^{-T}."""
    t = np.asarray(t, dtype=np.float64)
    g2 = 2 * self.Gamma * self.beta(t)                         # (G@G)[1,1]
    iRT = np.swapaxes(self.invR(t), -1, -2)
    out = np.zeros(np.shape(t) + (2, 2))
    out[..., 1, :] = 0.5 * g2[..., None] * iRT[..., 1, :]
    return out

  def eps2score(self, eps, t):
    """sde_lib.py:246-253 with a scalar t shared by the batch: score = -R^{-T} eps on the last axis."""
    m = -self.invR(t).T
    return np.einsum("ij,...j->...i", m, eps)

  # ---- tables ---------------------------------------------------------------------------------
  def prepare_order0_coef(self, rev_ts, num_item=1000):
    """sde_lib.py:289-306."""
    rev_ts = np.asarray(rev_ts, dtype=np.float64)
    mean = self.psi(rev_ts[:-1], rev_ts[1:])
    eps = np.stack([_riemann(self, s, e, None, 0, num_item) for s, e in zip(rev_ts[:-1], rev_ts[1:])])
    return mean, eps

  def prepare_naive_coef(self, rev_ts):
    """sde_lib.py:276-287: Euler step matrices I + F(t) dt and 0.5 G G R^{-T} dt."""
    rev_ts = np.asarray(rev_ts, dtype=np.float64)
    dts = rev_ts[1:] - rev_ts[:-1]
    mean = np.stack([np.eye(2) + self.s_F(t) * dt for t, dt in zip(rev_ts[:-1], dts)])
    eps = self.eps_integrand(rev_ts[:-1]) * dts[:, None, None]
    return mean, eps

  def get_deis_coef(self, order, rev_ts):
    """sde_lib.py:308-319 -> [N, order+3, 2, 2]."""
    rev_ts = np.asarray(rev_ts, dtype=np.float64)
    x_coef = self.psi(rev_ts[:-1], rev_ts[1:])
    eps_coef = get_ab_eps_coef(self, order + 1, rev_ts, order)
    return np.concatenate([x_coef[:, None], eps_coef], axis=1)


def from_config(config):
  """sde_lib.py:321-331."""
  m = config.model
  return CLD(m_inv=m.m_inv, beta_0=m.beta_0, beta_1=m.beta_1, vv_gamma=m.vv_gamma,
             mixed_score=m.mixed_score, is_R_rk=m.is_R_rk, R_dt=m.R_dt)


# ---- DEIS coefficient tables (cld_jax/deis.py:19-95) ---------------------------------------------
def _lagrange(t_val, ts_poly, coef_idx):
  """single_poly_coef / vec_poly_coef, deis.py:30-38, vectorised over t_val."""
  out = np.ones_like(t_val)
  for k in range(len(ts_poly)):
    if k != coef_idx:
      out = out * (t_val - ts_poly[k]) / (ts_poly[coef_idx] - ts_poly[k])
  return out


def _riemann(sde, t_start, t_end, ts_poly, coef_idx, num_item=10000):
  """get_eps_coef_worker_fn + get_eps_single_coef_fn, deis.py:19-47: left Riemann sum, num_item nodes."""
  dt = (t_end - t_start) / num_item
  t_inter = np.linspace(t_start, t_end, num_item, endpoint=False)
  integrand = sde.psi(t_inter, t_end) @ sde.eps_integrand(t_inter)
  w = np.ones(num_item) if ts_poly is None else _lagrange(t_inter, ts_poly, coef_idx)
  return np.sum(integrand * w[:, None, None], axis=0) * dt


def _coef_row(sde, highest_order, order, t_start, t_end, ts_poly):
  """get_eps_coef_fn._worker, deis.py:49-59: rtn[j] uses node ts_poly[order-j] (the jnp.flip)."""
  rtn = np.zeros((highest_order + 1, 2, 2))
  ts_poly = ts_poly[:order + 1]
  for j, coef_idx in enumerate(range(order, -1, -1)):
    rtn[j] = _riemann(sde, t_start, t_end, ts_poly, coef_idx)
  return rtn


def get_ab_eps_coef(sde, highest_order, timesteps, order):
  """deis.py:61-95 (recursion included)."""
  timesteps = np.asarray(timesteps, dtype=np.float64)
  if order == 0:
    return np.stack([_coef_row(sde, highest_order, 0, timesteps[i], timesteps[i + 1], timesteps[i:i + 1])
                     for i in range(len(timesteps) - 1)])
  prev = get_ab_eps_coef(sde, highest_order, timesteps[:order + 1], order - 1)
  rows = []
  for k in range(len(timesteps) - order - 1):
    ts_poly = timesteps[k:k + order + 1]
    rows.append(_coef_row(sde, highest_order, order, timesteps[order + k], timesteps[order + k + 1], ts_poly))
  cur = np.stack(rows) if rows else np.zeros((0, highest_order + 1, 2, 2))
  return np.concatenate([prev, cur], axis=0)


# ---- update ops -----------------------------------------------------------------------------------
def multistep_ab_step(x, deis_coef, new_eps, eps_pred):
  """deis.py:141-151."""
  x_coef, eps_coef = deis_coef[0], deis_coef[1:]
  full_eps = np.concatenate([new_eps[None], eps_pred])
  linear_term = np.einsum("ij,b...j->b...i", x_coef, x)
  eps_term = np.einsum("oij,o...j->...i", eps_coef, full_eps)
  return linear_term + eps_term, full_eps[:-1]


def relayout_in(u):
  """'b ... d g -> b ... (g d)', models/utils.py:153."""
  return np.concatenate([u[..., 0], u[..., 1]], axis=-1)


def relayout_out(o):
  """'b ... (g d) -> b ... d g', g=2, models/utils.py:158."""
  c = o.shape[-1] // 2
  return np.stack([o[..., :c], o[..., c:]], axis=-1)


def make_eps_fn(sde, net_fn):
  """get_eps_fn, models/utils.py:168-182.  net_fn(x[B,H,W,2C], labels scalar) -> [B,H,W,2C]."""
  def eps_fn(u, t):
    out = relayout_out(np.asarray(net_fn(relayout_in(u), 999.0 * t), dtype=u.dtype))
    if sde.mixed_score:
      u0 = u.copy()
      u0[..., 0] = 0.0
      out = out + np.einsum("ij,...j->...i", sde.invR(t), u0).astype(u.dtype)
    return out
  return eps_fn


def denoise_step(sde, eps_fn, u):
  """get_denoising_step, sampling.py:30-39, with t = denoising_eps = sde.sampling_eps."""
  t = sde.sampling_eps
  F, G = sde.s_F(t), sde.s_G(t)
  dt = -t
  eps = eps_fn(u, t)
  score = sde.eps2score(eps, t)
  return u + np.einsum("ij,...j->...i", F, u) * dt - np.einsum("ij,...j->...i", G @ G, score) * dt


def deis_sampler(sde, eps_fn, u, nfe, deis_order, ts_order=2, denoising=True, centered=True,
                 dtype=np.float64, trace=None, rev_ts=None):
  """_impl_deis_sampler.sampler + get_deis_sampler, sampling.py:204-253 (single device, explicit u)."""
  num_step = nfe - 1 if denoising else nfe
  if rev_ts is None:
    rev_ts = get_rev_ts(sde.T, sde.sampling_eps, ts_order, num_step)
  assert len(rev_ts) == num_step + 1                         # sampling.py:268
  coef = sde.get_deis_coef(deis_order, rev_ts).astype(dtype)
  u = np.asarray(u, dtype=dtype)
  eps_pred = np.stack([u] * (deis_order + 1))
  for i in range(num_step):
    eps = np.asarray(eps_fn(u, rev_ts[i]), dtype=dtype)
    u, eps_pred = multistep_ab_step(u, coef[i], eps, eps_pred)
    u = u.astype(dtype)
    if trace is not None:
      trace.append(u.copy())
  if denoising:
    u = denoise_step(sde, eps_fn, u).astype(dtype)
  x, v = u[..., 0], u[..., 1]
  if centered:
    x = (x + 1.0) / 2.0
  return x, v, nfe


def hyd_rev_ts(sde, nfe, noise_nfe_ratio=0.3, img_t_ratio=0.3, ts_order=2.0, denoising=True):
  """get_hyd_deis_sampler's grid, sampling.py:255-268 (the polynomial part restarts at sde.T, as written there)."""
  num_step = nfe - 1 if denoising else nfe
  mid_t = sde.T * img_t_ratio
  noise_nfe = int(num_step * noise_nfe_ratio)
  img_nfe = num_step - noise_nfe
  noise_ts = np.linspace(sde.T, mid_t, noise_nfe, endpoint=False)
  return np.concatenate([noise_ts, get_rev_ts(sde.T, sde.sampling_eps, ts_order, img_nfe)])


def order0_sampler(sde, eps_fn, u, nfe, denoising=True, centered=True, dtype=np.float64):
  """get_order0_sampler (is_em=False), sampling.py:156-202; ts_order hard-coded 2 (162)."""
  num_step = nfe - 1 if denoising else nfe
  rev_ts = get_rev_ts(sde.T, sde.sampling_eps, 2, num_step)
  mean, epsm = sde.prepare_order0_coef(rev_ts)
  u = np.asarray(u, dtype=dtype)
  for i in range(num_step):
    eps = np.asarray(eps_fn(u, rev_ts[i]), dtype=dtype)
    u = (np.einsum("ij,...j->...i", mean[i], u) + np.einsum("ij,...j->...i", epsm[i], eps)).astype(dtype)
  if denoising:
    u = denoise_step(sde, eps_fn, u).astype(dtype)
  x, v = u[..., 0], u[..., 1]
  if centered:
    x = (x + 1.0) / 2.0
  return x, v, nfe


def prior_sampling(rng, shape, m_inv=4.0):
  """CLD.prior_sampling, sde_lib.py:270-274, with a numpy Generator instead of jax threefry."""
  xs = rng.standard_normal(shape)
  vs = rng.standard_normal(shape) / np.sqrt(m_inv)
  return np.stack([xs, vs], axis=-1)


# ---- stochastic gDDIM: LambdaSDE + sdeis sampler (cld_jax/sde_lib.py:334-466, sampling.py:380-427) --------------
class LambdaSDE:
  """sde_lib.py:334-466, fp64.  The reference integrates hat-Psi(0,t) with 1e5 RK4 steps (dt = 1e-5) and every
  conditional covariance with 1e4 RK4 steps; `n_hat` / `n_cov` allow shorter grids in tests."""

  def __init__(self, sde, lambda_coef=0.1, use_order0=True, hat_dt=1e-5, n_cov=10_000):
    self.sde, self.lambda_coef, self.use_order0 = sde, float(lambda_coef), use_order0
    self.mixed_score, self.T, self.sampling_eps = sde.mixed_score, sde.T, sde.sampling_eps
    self.n_cov = n_cov
    dt = hat_dt                                                 # sde_lib.py:358 (dt = 1e-5)
    ts = np.linspace(0, 1.0 + dt, int(1.0 / dt) + 1, endpoint=False)   # :370-373 (int(1/1e-5) == 99999)
    out = np.empty((len(ts), 2, 2))
    x = np.eye(2)
    fn = lambda _x, _t: self.s_hat_F(_t) @ _x
    for k, t in enumerate(ts):                                  # scan emits the carry before the update (:366-367)
      out[k] = x
      x = _rk4(x, t, dt, fn)
    self._hxp, self._hfp = ts, out

  def s_hat_F(self, t):
    """sde_lib.py:352-356."""
    G = self.sde.s_G(t)
    return self.sde.s_F(t) + 0.5 * (1 + self.lambda_coef ** 2) * G @ G.T @ inv_2x2(self.sde.cov(t))

  def s_hat_psi_02t(self, t):
    xp, fp = self._hxp, self._hfp
    i = int(np.clip(np.searchsorted(xp, t, side="right"), 1, len(xp) - 1))
    return fp[i - 1] + (t - xp[i - 1]) / (xp[i] - xp[i - 1]) * (fp[i] - fp[i - 1])

  def s_hat_psi(self, s, t):
    return self.s_hat_psi_02t(t) @ inv_2x2(self.s_hat_psi_02t(s))

  def cond_rev_cov(self, s, t):
    """sde_lib.py:381-399, literally (note `_x @ cur_hat_F`, not its transpose, and ts built with endpoint=False
    over n_step + 1 points while the step is (t - s)/n_step)."""
    sign = 1.0 if t > s else -1.0
    n = self.n_cov
    dt = (t - s) / n
    ts = np.linspace(s, t, n + 1, endpoint=False)

    def fn(_x, _t):
      hF, G = self.s_hat_F(_t), self.sde.s_G(_t)
      return hF @ _x + _x @ hF + sign * self.lambda_coef ** 2 * G @ G.T
    cov = np.zeros((2, 2))
    for i in range(n):
      cov = _rk4(cov, ts[i], dt, fn)
    return cov

  def update_coef(self, s, t):
    x_coef = self.sde.psi(s, t)
    eps_coef = (self.s_hat_psi(s, t) - x_coef) @ self.sde.R(s)
    return np.stack([x_coef, eps_coef, self.cond_rev_cov(s, t)])

  def get_poly_eps_coef(self, order, rev_ts):
    """sde_lib.py:410-433."""
    outer = self

    class _sde:
      @staticmethod
      def psi(ss, t):
        return np.stack([outer.s_hat_psi(s, t) for s in np.atleast_1d(ss)])

      @staticmethod
      def eps_integrand(ts):
        o = []
        for _t in np.atleast_1d(ts):
          G = outer.sde.s_G(_t)
          o.append(0.5 * (1 + outer.lambda_coef ** 2) * G @ G.T @ inv_2x2(outer.sde.cov(_t)) @ outer.sde.psi(0.0, _t))
        return np.stack(o)
    ab = get_ab_eps_coef(_sde, order + 1, rev_ts, order)                    # [N, order+2, 2, 2]
    last = np.stack([self.sde.psi(s, 0.0) @ self.sde.R(s) for s in rev_ts[:-1]])
    return np.einsum("b...ij,bjk->b...ik", ab, last)

  def get_deis_coef(self, order, rev_ts):
    """sde_lib.py:435-454 -> [N, order+4, 2, 2]: x_coef, order+2 eps slots, covariance."""
    rev_ts = np.asarray(rev_ts, np.float64)
    if self.use_order0 and order == 0:
      c = np.stack([self.update_coef(s, t) for s, t in zip(rev_ts[:-1], rev_ts[1:])])
      return np.stack([c[:, 0], c[:, 1], np.zeros_like(c[:, 0]), c[:, 2]], axis=1)
    x_coef = self.sde.psi(rev_ts[:-1], rev_ts[1:])
    eps_coef = self.get_poly_eps_coef(order, rev_ts)
    covs = np.stack([self.cond_rev_cov(s, t) for s, t in zip(rev_ts[:-1], rev_ts[1:])])
    return np.concatenate([x_coef[:, None], eps_coef, covs[:, None]], axis=1)


def _rk4(x, t, dt, fn):
  """deis.py:5-17."""
  g1 = fn(x, t)
  g2 = fn(x + g1 * dt / 2, t + dt / 2)
  g3 = fn(x + g2 * dt / 2, t + dt / 2)
  g4 = fn(x + g3 * dt, t + dt)
  return x + dt / 6 * (g1 + 2 * g2 + 2 * g3 + g4)


def mvn_factor_svd(cov):
  """The factor jax.random.multivariate_normal(method='svd') applies to standard normals: u * sqrt(s)."""
  u, s, _ = np.linalg.svd(cov)
  # the sign of a singular vector is implementation-defined (LAPACK vs cuSolver under jax): fix it so that the
  # largest-magnitude entry of every column is positive (first entry on ties), the convention the library uses
  for j in range(u.shape[1]):
    k = 0 if abs(u[0, j]) >= abs(u[1, j]) else 1
    if u[k, j] < 0:
      u[:, j] = -u[:, j]
  return u * np.sqrt(s)[None, :]


def sdeis_sampler(lsde, eps_fn, u, nfe, deis_order, z, ts_order=2, denoising=True, centered=True, dtype=np.float64,
                  trace=None):
  """_impl_sdeis_sampler + _impl_sdeis_update_fn, sampling.py:380-427.  `z` [num_step, *u.shape] are the standard
  normals (the reference draws them with jax.random; here they are an input so that both sides share them)."""
  sde = lsde.sde
  num_step = nfe - 1 if denoising else nfe
  rev_ts = get_rev_ts(sde.T, sde.sampling_eps, ts_order, num_step)
  coef = lsde.get_deis_coef(deis_order, rev_ts)
  coef[-1, -1] = 0.0                                            # sampling.py:420
  coef = coef.astype(dtype)
  u = np.asarray(u, dtype=dtype)
  eps_pred = np.stack([u] * (deis_order + 1))
  for i in range(num_step):
    eps = np.asarray(eps_fn(u, rev_ts[i]), dtype=dtype)
    mean, eps_pred = multistep_ab_step(u, coef[i][:-1], eps, eps_pred)
    noise = np.einsum("ij,...j->...i", mvn_factor_svd(coef[i][-1].astype(np.float64)), z[i])
    u = (mean + noise).astype(dtype)
    if trace is not None:
      trace.append(u.copy())
  if denoising:
    u = denoise_step(sde, eps_fn, u).astype(dtype)
  x, v = u[..., 0], u[..., 1]
  if centered:
    x = (x + 1.0) / 2.0
  return x, v, nfe


# ---- remaining CLD samplers (cld_jax/sampling.py:497-669; sde_lib.py:469-519) ----------------------------------------
class LSDE:
  """sde_lib.py:469-519: the L_t (Cholesky) parameterisation used by the 'ldeis' baseline."""

  def __init__(self, sde):
    self.sde, self.mixed_score, self.T, self.sampling_eps = sde, sde.mixed_score, sde.T, sde.sampling_eps

  def s_L(self, t):
    return np.linalg.cholesky(self.sde.cov(t))

  def epsR2epsL_matrix(self, t):
    """coef of epsR2epsL (sde_lib.py:494-499): L^T R^-T."""
    return self.s_L(t).T @ inv_2x2(self.sde.R(t).T)

  def psi(self, s, t):
    return self.sde.psi(s, t)

  def eps_integrand(self, ts):
    out = []
    for t in np.atleast_1d(ts):
      G = self.sde.s_G(t)
      out.append(0.5 * G @ G @ inv_2x2(self.s_L(t)).T)
    return np.stack(out)

  def get_deis_coef(self, order, rev_ts):
    rev_ts = np.asarray(rev_ts, np.float64)
    x_coef = self.sde.psi(rev_ts[:-1], rev_ts[1:])
    return np.concatenate([x_coef[:, None], get_ab_eps_coef(self, order + 1, rev_ts, order)], axis=1)


def ldeis_sampler(sde, eps_fn, u, nfe, deis_order, ts_order=2, denoising=False, centered=True, dtype=np.float64):
  """_impl_Ldeis_sampler + get_L_deis_sampler, sampling.py:497-540 (the reference's denoising branch needs
  LSDE.s_F, which does not exist; denoising here falls back to the base SDE's step)."""
  lsde = LSDE(sde)
  num_step = nfe - 1 if denoising else nfe
  rev_ts = get_rev_ts(sde.T, sde.sampling_eps, ts_order, num_step)
  coef = lsde.get_deis_coef(deis_order, rev_ts).astype(dtype)
  u = np.asarray(u, dtype=dtype)
  eps_pred = np.stack([u] * (deis_order + 1))
  for i in range(num_step):
    eps = np.asarray(eps_fn(u, rev_ts[i]), dtype=np.float64)
    eps = np.einsum("ij,...j->...i", lsde.epsR2epsL_matrix(rev_ts[i]), eps).astype(dtype)
    u, eps_pred = multistep_ab_step(u, coef[i], eps, eps_pred)
    u = u.astype(dtype)
  if denoising:
    u = denoise_step(sde, eps_fn, u).astype(dtype)
  x, v = u[..., 0], u[..., 1]
  return ((x + 1.0) / 2.0 if centered else x), v, nfe


def em_sampler(sde, eps_fn, u, nfe, z, lambda_coef=0.0, ts_order=2, denoising=False, centered=True, dtype=np.float64):
  """get_em_sampler, sampling.py:624-669; z [num_step, *u.shape] standard normals."""
  num_step = nfe - 1 if denoising else nfe
  rev_ts = get_rev_ts(sde.T, sde.sampling_eps, ts_order, num_step)
  u = np.asarray(u, dtype=dtype)
  for i in range(num_step):
    cur_t, next_t = rev_ts[i], rev_ts[i + 1]
    dt = next_t - cur_t
    G = sde.s_G(cur_t)
    score = sde.eps2score(np.asarray(eps_fn(u, cur_t), np.float64), cur_t)
    grad = np.einsum("ij,...j->...i", sde.s_F(cur_t), u) - (1.0 + lambda_coef) / 2.0 * np.einsum("ij,...j->...i", G @ G.T, score)
    noise = z[i] * np.sqrt(np.abs(dt))
    u = (u + grad * dt + np.einsum("ij,...j->...i", G, noise) * lambda_coef).astype(dtype)
  if denoising:
    u = denoise_step(sde, eps_fn, u).astype(dtype)
  x, v = u[..., 0], u[..., 1]
  return ((x + 1.0) / 2.0 if centered else x), v, nfe


def _sscs_ou(sde, u, s_t, s_t_next, z):
  """get_sscs_ou_fn, sampling.py:542-566."""
  bi = -1 * (sde.beta_int(1 - s_t_next) - sde.beta_int(1 - s_t))
  Gm = sde.Gamma
  mean_matrix = np.array([[1 + 2 * bi / Gm, -4 * bi / Gm / Gm], [bi, 1 - 2 * bi / Gm]]) * np.exp(-2.0 * bi / Gm)
  cov_xx = np.exp(4 * bi / Gm) - 1 - 4 * bi / Gm - 8 * bi ** 2 / Gm / Gm
  cov_xv = -4 * bi ** 2 / Gm
  cov_vv = (Gm / 2) ** 2 * (np.exp(4 * bi / Gm) - 1) + bi * Gm - 2 * bi ** 2
  cov = np.array([[cov_xx, cov_xv], [cov_xv, cov_vv]]) * np.exp(-4 * bi / Gm)
  return np.einsum("ij,...j->...i", mean_matrix, u) + np.einsum("ij,...j->...i", mvn_factor_svd(cov), z)


def sscs_sampler(sde, eps_fn, u, nfe, z, ts_order=2, denoising=False, centered=True, dtype=np.float64):
  """get_sscs_sampler, sampling.py:568-622; z [num_step, 2, *u.shape]: the two OU half-step draws of every step."""
  num_step = nfe - 1 if denoising else nfe
  ts = 1 - get_rev_ts(sde.T, sde.sampling_eps, ts_order, num_step)
  u = np.asarray(u, dtype=dtype)
  for i in range(num_step):
    cur_t, next_t = ts[i], ts[i + 1]
    mid = (cur_t + next_t) / 2.0
    u = _sscs_ou(sde, u, cur_t, mid, z[i, 0]).astype(dtype)
    score = sde.eps2score(np.asarray(eps_fn(u, sde.T - cur_t), np.float64), sde.T - cur_t)
    v = u[..., 1] + 2 * sde.beta(cur_t) * sde.Gamma * (score[..., 1] + sde.m_inv * u[..., 1]) * (next_t - cur_t)
    u = np.stack([u[..., 0], v], axis=-1).astype(dtype)
    u = _sscs_ou(sde, u, mid, next_t, z[i, 1]).astype(dtype)
  if denoising:
    u = denoise_step(sde, eps_fn, u).astype(dtype)
  x, v = u[..., 0], u[..., 1]
  return ((x + 1.0) / 2.0 if centered else x), v, nfe


def ode_sampler(sde, eps_fn, u, denoising=False, rtol=1e-5, atol=1e-5, method="RK45", centered=True):
  """get_ode_sampler.sampler, sampling.py:432-464: scipy RK45 on the probability-flow ODE."""
  from scipy import integrate
  shape = u.shape

  def ode_func(t, xf):
    x = xf.reshape(shape)
    score = sde.eps2score(np.asarray(eps_fn(x.astype(np.float32), t), np.float64), t)
    F, G = sde.s_F(t), sde.s_G(t)
    return (np.einsum("ij,...j->...i", F, x) - 0.5 * np.einsum("ij,...j->...i", G @ G, score)).reshape(-1)
  sol = integrate.solve_ivp(ode_func, (sde.T, sde.sampling_eps), np.asarray(u, np.float64).reshape(-1), rtol=rtol,
                            atol=atol, method=method)
  uo = sol.y[:, -1].reshape(shape)
  if denoising:
    uo = denoise_step(sde, eps_fn, uo)
  x, v = uo[..., 0], uo[..., 1]
  return ((x + 1.0) / 2.0 if centered else x), v, sol.nfev


class MLCLD:
  """cld_jax/sampling.py:272-325 on the CLD helpers sde_lib.py:120-181 (rotating frame psi1 = expm(int F_1))."""

  def __init__(self, sde, n=100_000):
    assert sde.beta_1 == 0
    self.sde, self.T, self.sampling_eps, self.mixed_score = sde, sde.T, sde.sampling_eps, sde.mixed_score
    dt = 1.0 / n
    fn = lambda p2, t: self.inv_psi1(t) @ self.F2(t) @ self.psi1(t) @ p2
    xs, ts = np.empty((n + 1, 2, 2)), np.empty(n + 1)
    x, t = np.eye(2), 0.0
    for k in range(n + 1):                                     # scan emits (prev_psi2, cur_t), sampling.py:276-283
      xs[k], ts[k] = x, t
      x = _rk4(x, t, dt, fn)
      t = t + dt
    self._xp, self._fp = ts, xs

  def f1_psi(self, s, t):
    bi = self.sde.beta_int(t) - self.sde.beta_int(s)
    sm, ism = np.sqrt(1.0 / self.sde.m_inv), np.sqrt(self.sde.m_inv)
    return np.array([[np.cos(bi * ism), ism * np.sin(bi * ism)], [-sm * np.sin(bi * ism), np.cos(bi * ism)]])

  def psi1(self, t):
    return self.f1_psi(0.0, t)

  def inv_psi1(self, t):
    return self.f1_psi(t, 0.0)

  def F2(self, t):
    return np.array([[0.0, 0.0], [0.0, -self.sde.Gamma * self.sde.beta(t) * self.sde.m_inv]])

  def psi2(self, t):
    xp, fp = self._xp, self._fp
    i = int(np.clip(np.searchsorted(xp, t, side="right"), 1, len(xp) - 1))
    return fp[i - 1] + (t - xp[i - 1]) / (xp[i] - xp[i - 1]) * (fp[i] - fp[i - 1])

  def psi(self, ss, t):
    p2t = self.psi2(t)
    if np.ndim(ss) == 0:
      return p2t @ np.linalg.inv(self.psi2(ss))
    return np.stack([p2t @ np.linalg.inv(self.psi2(s)) for s in ss])

  def eps_integrand(self, ts):
    out = []
    for t in np.atleast_1d(ts):
      G = self.sde.s_G(t)
      out.append(0.5 * self.inv_psi1(t) @ G @ G.T @ self.sde.invR(t).T)
    return np.stack(out)

  def get_deis_coef(self, order, rev_ts):
    rev_ts = np.asarray(rev_ts, np.float64)
    x_coef = np.stack([self.psi(s, t) for s, t in zip(rev_ts[:-1], rev_ts[1:])])
    return np.concatenate([x_coef[:, None], get_ab_eps_coef(self, order + 1, rev_ts, order)], axis=1)


def mldeis_sampler(sde, eps_fn, u, nfe, deis_order, ts_order=2, denoising=False, centered=True, dtype=np.float64,
                   ml=None):
  """get_mldeis_sampler, sampling.py:328-378 (the denoising step acts on the rotated variable, as written there)."""
  ml = ml or MLCLD(sde)
  num_step = nfe - 1 if denoising else nfe
  rev_ts = get_rev_ts(sde.T, sde.sampling_eps, ts_order, num_step)
  coef = ml.get_deis_coef(deis_order, rev_ts).astype(dtype)
  y = np.einsum("ij,...j->...i", ml.inv_psi1(sde.T), np.asarray(u, np.float64)).astype(dtype)
  eps_pred = np.stack([y] * (deis_order + 1))
  for i in range(num_step):
    x_u = np.einsum("ij,...j->...i", ml.psi1(rev_ts[i]), y.astype(np.float64)).astype(dtype)
    eps = np.asarray(eps_fn(x_u, rev_ts[i]), dtype=dtype)
    y, eps_pred = multistep_ab_step(y, coef[i], eps, eps_pred)
    y = y.astype(dtype)
  if denoising:
    y = denoise_step(sde, eps_fn, y).astype(dtype)
  uo = np.einsum("ij,...j->...i", ml.psi1(sde.sampling_eps / 2), y.astype(np.float64)).astype(dtype)
  x, v = uo[..., 0], uo[..., 1]
  return ((x + 1.0) / 2.0 if centered else x), v, nfe
