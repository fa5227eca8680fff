"""Stand-in for blur_jax/sde_lib.py: blurring-diffusion SDE in DCT space (sde_lib.py:18-163).

Scalars / per-frequency tables come from the fp64 host code of libgddim_b200.so; DCTs run on the GPU.
"""
import ctypes as C

import numpy as np

from .. import _lib
from . import blur


def _np_rng(rng):
  if isinstance(rng, np.random.Generator):
    return rng
  if rng is None:
    return np.random.default_rng()
  return np.random.default_rng(np.asarray(rng).astype(np.uint32).ravel().tolist())


class SDE:
  def __init__(self, min_scale=0.001, sigma_blur_max=10.0, sampling_eps=1e-5):
    self.min_scale, self.sigma_blur_max, self.sampling_eps = min_scale, sigma_blur_max, sampling_eps
    if min_scale != 0.001:
      raise ValueError("min_scale is fixed to 0.001 in the library tables")
    h = C.c_void_p()
    _lib.check(_lib.lib().gddim_blur_create(float(sigma_blur_max), float(sampling_eps), C.byref(h)))
    self._h = h
    self.T = 1.0
    self.alpha_start = float(_lib.lib().gddim_blur_t2alpha(self._h, 0.0))      # fp64: rho2t / sampling_T derive from it

  def __del__(self):
    try:
      if getattr(self, "_h", None):
        _lib.lib().gddim_blur_destroy(self._h)
        self._h = None
    except Exception:
      pass

  @property
  def sampling_T(self):
    """sde_lib.py:33-35: rho2t(80)."""
    return float(_lib.lib().gddim_blur_sampling_T(self._h))

  def t2alpha_fn(self, t):
    t = np.asarray(t, np.float64)
    return np.vectorize(lambda s: _lib.lib().gddim_blur_t2alpha(self._h, float(s)))(t).astype(np.float32) \
        if t.ndim else np.float32(_lib.lib().gddim_blur_t2alpha(self._h, float(t)))

  def alpha2t_fn(self, alpha):
    return np.arccos(np.sqrt(alpha)) * 2 / np.pi * 1.008 - 0.004

  def rho2t(self, rho):
    a0 = float(self.alpha_start)
    return self.alpha2t_fn(a0 / ((rho + np.sqrt(1 - a0)) ** 2 + a0))

  def _mean1(self, t):
    out = np.empty((32, 32))
    _lib.check(_lib.lib().gddim_blur_y_mean_coef(self._h, float(t), out.ctypes.data))
    return out

  def get_frequency_scaling(self, t):
    """sde_lib.py:79-88: t (B,) -> (B,32,32,1)."""
    t = np.atleast_1d(np.asarray(t, np.float64))
    sa = np.sqrt(np.asarray([_lib.lib().gddim_blur_t2alpha(self._h, float(s)) for s in t]))
    return (np.stack([self._mean1(s) for s in t]) / sa[:, None, None])[..., None].astype(np.float32)

  def y_mean_coef(self, t):
    """sde_lib.py:90-93: t (B,) -> (B,32,32,1)."""
    t = np.atleast_1d(np.asarray(t, np.float64))
    return np.stack([self._mean1(s) for s in t])[..., None].astype(np.float32)

  def y_std_coef(self, t):
    """sde_lib.py:95-97: t (B,) -> (B,)."""
    t = np.atleast_1d(np.asarray(t, np.float64))
    return np.asarray([_lib.lib().gddim_blur_y_std_coef(self._h, float(s)) for s in t], np.float32)

  def prior_sampling(self, rng, shape):
    return _np_rng(rng).standard_normal(tuple(shape)).astype(np.float32)

  def x2y(self, xs):
    return blur.batch_img_dct(xs)

  def y2x(self, ys):
    return blur.batch_img_idct(ys)

  def encode_t(self, t):
    return 999 * t

  def encode_x(self, xs):
    return xs

  def model2eps(self, xs, ts, model_output):
    del xs, ts
    return model_output


def from_config(config):
  return SDE(sigma_blur_max=config.model.sigma_blur_max, sampling_eps=config.sampling.t0)
