"""ORACLE (test infrastructure, NOT product code) -- blurring-diffusion SDE, DCT and order-0 sampler, numpy fp64.

Parity status: PINNED to outputs of the reference itself (blur_jax/{sde_lib,blur,fft,sampling,multistep}.py run
unmodified under tests/refshim; fixtures tests/golden/ref_blur_*.npz; tests/test_ref_golden.py): schedule tables
for sigma_blur_max 1 and 10 (1e-12), the 51-point time grid, rho2t, batch_img_dct / idct (1e-11), ab_step (1e-13),
the network forward (1e-10) and the order-0 sampler end to end, with and without pmap (1e-8).  scipy.fft.dctn,
IDCT(DCT) = id and the grid endpoints stay in tests/test_oracle_blur.py.

Restates (file:line relative to /root/reference/blur_jax):
  blur.py:11-107          Makhoul-FFT DCT-II / DCT-III, batch_img_dct / batch_img_idct
  sde_lib.py:18-163       SDE: t2alpha_fn, alpha2t_fn, rho2t, get_frequency_scaling, y_mean_coef, y_std_coef
  sampling.py:42-90       get_rev_ts, get_order0_sampler
  models/utils.py:141-160 get_eps_fn (labels = encode_t(t) = 999 t), get_yeps_fn = DCT o net o IDCT
  multistep.py:94-98      ab_step (scalar-coefficient update; dead code in the reference)
"""
import numpy as np


# ---- DCT via FFT, literal restatement of blur.py ------------------------------------------------
def _impl_dct(x):
  """blur.py:11-37 on the last axis."""
  N = x.shape[-1]
  v = np.concatenate([x[..., ::2], x[..., 1::2][..., ::-1]], axis=-1)
  f = np.fft.fft(v, axis=-1)
  k = -np.arange(N, dtype=np.float64) * np.pi / (2 * N)
  V = f.real * np.cos(k) - f.imag * np.sin(k)
  factor = np.concatenate([[np.sqrt(N) * 2], np.full(N - 1, np.sqrt(N * 2))])
  return 2 * V / factor


def _impl_idct(X):
  """blur.py:50-85 on the last axis."""
  N = X.shape[-1]
  factor = np.concatenate([[np.sqrt(N) * 2], np.full(N - 1, np.sqrt(N * 2))])
  Xv = X / 2 * factor
  k = np.arange(N, dtype=np.float64) * np.pi / (2 * N)
  Wr, Wi = np.cos(k), np.sin(k)
  Vtr = Xv
  Vti = np.concatenate([Xv[..., :1] * 0, -Xv[..., ::-1][..., :-1]], axis=-1)
  Vr = Vtr * Wr - Vti * Wi
  Vi = Vtr * Wi + Vti * Wr
  v = np.fft.irfft((Vr + 1j * Vi)[..., :N // 2 + 1], n=N, axis=-1)
  out = np.empty_like(v)
  out[..., ::2] = v[..., :N - N // 2]
  out[..., 1::2] = v[..., ::-1][..., :N // 2]
  return out


def batch_img_dct(xs):
  """blur.py:41-48,99-102: separable DCT-II over W then H of NHWC images."""
  x = np.transpose(xs, (0, 3, 1, 2))
  x = _impl_dct(x)                                          # over W
  x = np.swapaxes(_impl_dct(np.swapaxes(x, -1, -2)), -1, -2)  # over H
  return np.transpose(x, (0, 2, 3, 1))


def batch_img_idct(ys):
  y = np.transpose(ys, (0, 3, 1, 2))
  y = _impl_idct(y)
  y = np.swapaxes(_impl_idct(np.swapaxes(y, -1, -2)), -1, -2)
  return np.transpose(y, (0, 2, 3, 1))


# ---- SDE ------------------------------------------------------------------------------------------
class SDE:
  """blur_jax/sde_lib.py:18-163 (the pieces the sampler touches)."""

  def __init__(self, min_scale=0.001, sigma_blur_max=10.0, sampling_eps=1e-5):
    self.min_scale, self.sigma_blur_max, self.sampling_eps = min_scale, sigma_blur_max, sampling_eps
    img_dim = 32                                             # hard-coded, sde_lib.py:24
    freqs = np.pi * np.linspace(0, img_dim - 1, img_dim) / img_dim
    self.labda = freqs[:, None, None] ** 2 + freqs[None, :, None] ** 2      # [H, W, 1]
    self.alpha_start = self.t2alpha_fn(0.0)
    self.T = 1.0

  def t2alpha_fn(self, t):
    return np.cos((t + 0.004) / 1.008 * np.pi / 2) ** 2

  def alpha2t_fn(self, alpha):
    return np.arccos(np.sqrt(alpha)) * 2 / np.pi * 1.008 - 0.004

  def rho2t(self, rho):
    num = self.alpha_start
    denum = (rho + np.sqrt(1 - self.alpha_start)) ** 2 + self.alpha_start
    return self.alpha2t_fn(num / denum)

  @property
  def sampling_T(self):
    return self.rho2t(80.0)

  def get_frequency_scaling(self, t):
    """scalar t -> [H, W, 1]  (sde_lib.py:79-88)."""
    sigma_blur = self.sigma_blur_max * np.sin(t * np.pi / 2) ** 2
    dissipation_time = sigma_blur ** 2 / 2
    return np.exp(-dissipation_time * self.labda) * (1 - self.min_scale) + self.min_scale

  def y_mean_coef(self, t):
    return np.sqrt(self.t2alpha_fn(t)) * self.get_frequency_scaling(t)

  def y_std_coef(self, t):
    return np.sqrt(1 - self.t2alpha_fn(t))


def from_config(config):
  return SDE(sigma_blur_max=config.model.sigma_blur_max, sampling_eps=config.sampling.t0)


def get_rev_ts(sde, ts_order, num_step):
  """blur_jax/sampling.py:42-51."""
  return np.power(np.linspace(np.power(sde.sampling_T, 1.0 / ts_order),
                              np.power(sde.sampling_eps, 1.0 / ts_order), num_step + 1), ts_order)


def order0_sampler(sde, net_fn, y, nfe, ts_order=2, centered=True, dtype=np.float64, trace=None):
  """get_order0_sampler.sampler, blur_jax/sampling.py:53-80.  net_fn(x[B,H,W,C], labels) -> eps_x."""
  rev_ts = get_rev_ts(sde, ts_order, nfe)
  y = np.asarray(y, dtype=dtype)
  for i in range(nfe):
    t, tn = rev_ts[i], rev_ts[i + 1]
    y_eps = batch_img_dct(np.asarray(net_fn(batch_img_idct(y).astype(dtype), 999.0 * t), dtype=dtype))
    y_0 = 1.0 / sde.y_mean_coef(t) * (y - sde.y_std_coef(t) * y_eps)
    y = (sde.y_mean_coef(tn) * y_0 + sde.y_std_coef(tn) * y_eps).astype(dtype)
    if trace is not None:
      trace.append(y.copy())
  x = batch_img_idct(y)
  if centered:
    x = (x + 1.0) / 2.0
  return x.astype(dtype), nfe


def ab_step(x, ei_coef, new_eps, eps_pred):
  """blur_jax/multistep.py:94-98."""
  x_coef, eps_coef = ei_coef[0], ei_coef[1:]
  full_eps = np.concatenate([new_eps[None], eps_pred])
  eps_term = np.einsum("i,i...->...", eps_coef, full_eps)
  return x_coef * x + eps_term, full_eps[:-1]
