"""Stand-in for cld_jax/sde_lib.py: the CLD SDE (tables computed in fp64 by libgddim_b200.so on the host).

Mirrors `CLD` (sde_lib.py:45-319) and `from_config` (321-331).  Returned arrays are numpy float32 unless
x64=True (the reference runs with jax_enable_x64 = config.model.x64 = False).  The pickle cache of the
reference (`used_cache`) is not reproduced: tables are recomputed (sub-second).
`LambdaSDE` (334-404) and `LSDE` (407-519) are host-table classes over the same library (stochastic gDDIM and
the L-parameterised score); both are pinned to the reference's own classes in tests/test_ref_golden.py.
"""
import ctypes as C

import numpy as np

from .. import _lib


def inv_2x2(matrix):
  """sde_lib.py:17-26."""
  m = np.asarray(matrix)
  a, b, c, d = m[..., 0, 0], m[..., 0, 1], m[..., 1, 0], m[..., 1, 1]
  coef = 1.0 / (a * d - b * c)
  out = np.empty_like(m)
  out[..., 0, 0], out[..., 0, 1], out[..., 1, 0], out[..., 1, 1] = d * coef, -b * coef, -c * coef, a * coef
  return out


inv_2x2s = inv_2x2


def _d(a):
  return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class CLD:
  def __init__(self, m_inv=4.0, beta_0=4.0, beta_1=0.0, vv_gamma=0.04, numerical_eps=1e-6, mixed_score=False,
               used_cache=True, x64=False, is_R_rk=False, R_dt=1e-5):
    self.mixed_score, self.used_cache, self.x64 = mixed_score, used_cache, x64
    self.m_inv = m_inv
    self.Gamma = 2. / np.sqrt(m_inv)
    self.beta_0, self.beta_1 = beta_0, beta_1
    self.vv_gamma, self.numerical_eps = vv_gamma, numerical_eps
    self.is_R_rk, self.R_dt = bool(is_R_rk), float(R_dt)
    self._dt = np.float64 if x64 else np.float32
    self.R_0 = np.asarray([[np.sqrt(numerical_eps), 0], [0, np.sqrt(vv_gamma / m_inv + numerical_eps)]], self._dt)
    h = C.c_void_p()
    _lib.check(_lib.lib().gddim_cld_create(float(m_inv), float(beta_0), float(beta_1), float(vv_gamma),
                                            float(numerical_eps), float(R_dt), int(bool(is_R_rk)), C.byref(h)),
               "gddim_cld_create")
    self._h = h
    self.sampling_eps = 1e-3
    self.T = 1.0

  def __del__(self):
    try:
      if getattr(self, "_h", None):
        _lib.lib().gddim_cld_destroy(self._h)
        self._h = None
    except Exception:
      pass

  # -- schedule ------------------------------------------------------------------------------------------
  def beta(self, t):
    return self.beta_0 + self.beta_1 * t

  def beta_int(self, t):
    return self.beta_0 * t + 0.5 * self.beta_1 * t ** 2

  _beta, _beta_int = beta, beta_int

  # -- matrices (fp64 inside; "64" variants return fp64) --------------------------------------------------------
  def _R64(self, ts):
    ts = _d(ts).ravel()
    out = np.empty((ts.size, 2, 2))
    _lib.check(_lib.lib().gddim_cld_R(self._h, ts.ctypes.data, ts.size, out.ctypes.data))
    return out

  def _psi64(self, s, t):
    s, t = np.broadcast_arrays(_d(s), _d(t))
    s, t = _d(s).ravel(), _d(t).ravel()
    out = np.empty((s.size, 2, 2))
    _lib.check(_lib.lib().gddim_cld_psi(self._h, s.ctypes.data, t.ctypes.data, s.size, out.ctypes.data))
    return out

  def s_R(self, t):
    return self._R64([t])[0].astype(self._dt)

  def v_R(self, ts):
    return self._R64(ts).astype(self._dt)

  def s_invR(self, t):
    return inv_2x2(self._R64([t])[0]).astype(self._dt)

  def v_invR(self, ts):
    return inv_2x2(self._R64(ts)).astype(self._dt)

  def s_cov(self, t):
    r = self._R64([t])[0]
    return (r @ r.T).astype(self._dt)

  def v_cov(self, ts):
    r = self._R64(ts)
    return (r @ np.swapaxes(r, -1, -2)).astype(self._dt)

  def s_psi(self, s, t):
    return self._psi64(s, t)[0].astype(self._dt)

  def vv_psi(self, s, t):
    return self._psi64(s, t).astype(self._dt)

  def vs_psi(self, s, t):
    return self._psi64(s, t).astype(self._dt)

  def _F64(self, t):
    out = np.empty((2, 2))
    _lib.check(_lib.lib().gddim_cld_F(self._h, float(t), out.ctypes.data))
    return out

  def _G64(self, t):
    out = np.empty((2, 2))
    _lib.check(_lib.lib().gddim_cld_G(self._h, float(t), out.ctypes.data))
    return out

  def s_F(self, t):
    out = np.empty((2, 2))
    _lib.check(_lib.lib().gddim_cld_F(self._h, float(t), out.ctypes.data))
    return out.astype(self._dt)

  def s_G(self, t):
    out = np.empty((2, 2))
    _lib.check(_lib.lib().gddim_cld_G(self._h, float(t), out.ctypes.data))
    return out.astype(self._dt)

  def s_eps_integrand(self, s_t):
    return self.v_eps_integrand([s_t])[0]

  def v_eps_integrand(self, ts):
    ts = _d(ts).ravel()
    out = np.empty((ts.size, 2, 2))
    _lib.check(_lib.lib().gddim_cld_eps_integrand(self._h, ts.ctypes.data, ts.size, out.ctypes.data))
    return out.astype(self._dt)

  # -- batch helpers (host numpy; not on the per-step path) -------------------------------------------------------
  def eps2score(self, eps, ts):
    """sde_lib.py:246-253: score = -R^{-T} eps, eps (B, ..., c, 2), ts (B,)."""
    inv_rs = inv_2x2(self._R64(ts))
    return np.einsum("bji,b...dj->b...di", -inv_rs, np.asarray(eps, np.float64)).astype(self._dt)

  def mean(self, batch, ts):
    psis = self._psi64(np.zeros_like(_d(ts)), ts)
    return np.einsum("bij,b...dj->b...di", psis, np.asarray(batch, np.float64)).astype(self._dt)

  def perturb_data(self, batch, ts, rng):
    mean = self.mean(batch, ts)
    raw_noise = _np_rng(rng).standard_normal(mean.shape).astype(self._dt)
    perb = np.einsum("bij,b...dj->b...di", self._R64(ts), raw_noise.astype(np.float64)).astype(self._dt)
    return mean + perb, mean, raw_noise

  def prior_sampling(self, rng, shape):
    """sde_lib.py:270-274.  A jax PRNGKey (uint32[2]) reproduces jax.random's threefry stream
    (gddim_b200/jax_random.py: split into x / v keys, normal = sqrt(2) erfinv(uniform)); anything else (int seed,
    numpy Generator, None) seeds numpy's default_rng with two child streams."""
    if _is_jax_key(rng):
      from .. import jax_random
      return jax_random.cld_prior(np.asarray(rng, np.uint32), shape, self.m_inv).astype(self._dt)
    g = _np_rng(rng)
    seeds = g.integers(0, 2 ** 63 - 1, size=2)
    xs = np.random.default_rng(int(seeds[0])).standard_normal(tuple(shape)).astype(self._dt)
    vs = (np.random.default_rng(int(seeds[1])).standard_normal(tuple(shape)) / np.sqrt(self.m_inv)).astype(self._dt)
    return np.stack([xs, vs], axis=-1)

  # -- coefficient tables -----------------------------------------------------------------------------------------
  def prepare_naive_coef(self, rev_ts):
    """sde_lib.py:276-287 (Euler-Maruyama style coefficients)."""
    rev_ts = _d(rev_ts)
    mean, eps = [], []
    for cur, nxt in zip(rev_ts[:-1], rev_ts[1:]):
      mean.append(np.eye(2) + self.s_F(cur).astype(np.float64) * (nxt - cur))
      eps.append(self.s_eps_integrand(cur).astype(np.float64) * (nxt - cur))
    return np.stack(mean).astype(self._dt), np.stack(eps).astype(self._dt)

  def prepare_order0_coef(self, rev_ts):
    rev_ts = _d(rev_ts)
    n = rev_ts.size
    mean, eps = np.empty((n - 1, 2, 2)), np.empty((n - 1, 2, 2))
    _lib.check(_lib.lib().gddim_cld_order0_coef(self._h, rev_ts.ctypes.data, n, mean.ctypes.data, eps.ctypes.data))
    return mean.astype(self._dt), eps.astype(self._dt)

  def get_deis_coef(self, order, rev_timesteps, used_cache=True):
    """sde_lib.py:308-319 -> [N, order+3, 2, 2]."""
    rev = _d(rev_timesteps)
    out = np.empty((rev.size - 1, order + 3, 2, 2))
    _lib.check(_lib.lib().gddim_cld_deis_coef(self._h, int(order), rev.ctypes.data, rev.size, out.ctypes.data),
               "gddim_cld_deis_coef")
    return out.astype(self._dt)


def _is_jax_key(rng):
  return isinstance(rng, np.ndarray) and rng.dtype == np.uint32 and rng.shape == (2,)


def _np_rng(rng):
  if isinstance(rng, np.random.Generator):
    return rng
  if rng is None:
    return np.random.default_rng()
  return np.random.default_rng(np.asarray(rng).astype(np.uint32).ravel().tolist())


class LambdaSDE:
  """Stand-in for sde_lib.py:334-466 (stochastic gDDIM): hat-Psi table, conditional reverse covariance and the
  coefficient tables, computed in fp64 by the library.  Only what the sdeis sampler reads is exposed."""

  def __init__(self, sde, lambda_coef=0.1, use_order0=True, used_cache=True):
    self.sde = sde
    self.mixed_score = sde.mixed_score
    self.prior_sampling = sde.prior_sampling
    self.v_invR = sde.v_invR
    self.x64 = sde.x64
    self.use_order0, self.lambda_coef = use_order0, lambda_coef
    self.T, self.sampling_eps = sde.T, sde.sampling_eps
    self._h = sde._h
    self._dt = np.float64 if sde.x64 else np.float32

  def get_deis_coef(self, order, rev_timesteps, used_cache=True):
    """sde_lib.py:435-454 -> [N, order+4, 2, 2]: x_coef, order+2 eps slots, covariance."""
    rev = _d(rev_timesteps)
    out = np.empty((rev.size - 1, order + 4, 2, 2))
    _lib.check(_lib.lib().gddim_cld_sdeis_coef(self._h, float(self.lambda_coef), int(bool(self.use_order0)), int(order),
                                                rev.ctypes.data, rev.size, out.ctypes.data), "gddim_cld_sdeis_coef")
    return out.astype(self._dt)

  def get_order0_coef(self, rev_timesteps, used_cache=True):
    """sde_lib.py:457-466 -> [N, 3, 2, 2] (x_coef, eps_coef, cov)."""
    rev = _d(rev_timesteps)
    out = np.empty((rev.size - 1, 4, 2, 2))
    _lib.check(_lib.lib().gddim_cld_sdeis_coef(self._h, float(self.lambda_coef), 1, 0, rev.ctypes.data, rev.size,
                                                out.ctypes.data), "gddim_cld_sdeis_coef")
    return out[:, [0, 1, 3]].astype(self._dt)


def mvn_factor_svd(cov):
  """U sqrt(S) of a 2x2 matrix as jax.random.multivariate_normal(method='svd') applies it (sign-normalised)."""
  c = _d(cov)
  out = np.empty((2, 2))
  _lib.check(_lib.lib().gddim_mvn_factor_svd(c.ctypes.data, out.ctypes.data))
  return out


class LSDE:
  """Stand-in for sde_lib.py:469-519: the L_t (Cholesky of Sigma_t) parameterisation of the 'ldeis' baseline."""

  def __init__(self, sde, used_cache=True):
    self.sde = sde
    self.mixed_score = sde.mixed_score
    self.prior_sampling = sde.prior_sampling
    self.v_invR, self.s_G = sde.v_invR, sde.s_G
    self.vs_psi, self.vv_psi = sde.vs_psi, sde.vv_psi
    self.x64 = sde.x64
    self.T, self.sampling_eps = sde.T, sde.sampling_eps
    self._h = sde._h
    self._dt = np.float64 if sde.x64 else np.float32

  def _L64(self, t):
    r = self.sde._R64([t])[0]
    return np.linalg.cholesky(r @ r.T)

  def s_L(self, s_t):
    return self._L64(s_t).astype(self._dt)

  def _epsR2epsL_matrix64(self, t):
    return self._L64(t).T @ inv_2x2(self.sde._R64([t])[0].T)

  def epsR2epsL(self, s_t, eps):
    """sde_lib.py:493-499: (L^T R^-T) eps on the last axis."""
    return np.einsum("ij,b...dj->b...di", self._epsR2epsL_matrix64(s_t), np.asarray(eps, np.float64)).astype(self._dt)

  def s_eps_integrand(self, s_t):
    g = self.sde.s_G(s_t).astype(np.float64)
    return (0.5 * g @ g @ inv_2x2(self._L64(s_t)).T).astype(self._dt)

  def get_deis_coef(self, order, rev_timesteps, used_cache=True):
    rev = _d(rev_timesteps)
    out = np.empty((rev.size - 1, order + 3, 2, 2))
    _lib.check(_lib.lib().gddim_cld_ldeis_coef(self._h, int(order), rev.ctypes.data, rev.size, out.ctypes.data),
               "gddim_cld_ldeis_coef")
    return out.astype(self._dt)


def from_config(config):
  """sde_lib.py:321-331."""
  m = config.model
  return CLD(m_inv=m.m_inv, beta_0=m.beta_0, beta_1=m.beta_1, vv_gamma=m.vv_gamma, mixed_score=m.mixed_score,
             is_R_rk=m.is_R_rk, used_cache=m.used_cache, R_dt=m.R_dt, x64=m.x64)
