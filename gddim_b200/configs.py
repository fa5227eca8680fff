"""Reference configurations restated as plain data (SimpleNamespace trees).

The reference builds `ml_collections.ConfigDict`s in Python modules; `ml_collections` is not
installed here, and the hot path only *reads* attribute paths, so any object with the same
attributes works.  Values follow (relative to /root/reference):
  cld_jax/configs/default_cifar10_config.py:29-89, accr_dcifar10_config.py:15-47,
  ddpmpp_cifar10_config.py:17-43, simple_cifar10_config.py:36-63,
  blur_jax/configs/default_cifar10_config.py:31-34,68, ddpm_deep_cifar10_config.py:16-46.
"""
from types import SimpleNamespace as NS


def _common_model(**kw):
  m = NS(name="ncsnpp", scale_by_sigma=False, ema_rate=0.9999, normalization="GroupNorm", nonlinearity="swish",
         nf=128, ch_mult=(1, 2, 2, 2), num_res_blocks=8, attn_resolutions=(16,), resamp_with_conv=True,
         conditional=True, fir=True, fir_kernel=[1, 3, 3, 1], skip_rescale=True, resblock_type="biggan",
         progressive="none", progressive_input="residual", progressive_combine="sum", attention_type="ddpm",
         init_scale=0.0, embedding_type="fourier", fourier_scale=16, conv_size=3, dropout=0.1,
         sigma_min=0.01, sigma_max=50, num_scales=1000, beta_min=0.1, beta_max=20.0)
  for k, v in kw.items():
    setattr(m, k, v)
  return m


def _cld_base(model):
  for k, v in dict(m_inv=4.0, beta_0=4.0, beta_1=0.0, vv_gamma=0.04, mixed_score=False, is_R_rk=False,
                   R_dt=1e-5, used_cache=True, x64=False).items():
    if not hasattr(model, k):
      setattr(model, k, v)
  return NS(
      training=NS(continuous=True, batch_size=128),
      sampling=NS(method="deis", nfe=20, is_em=False, deis_order=1, ts_order=2, noise_removal=True,
                  noise_nfe_ratio=0.3, img_t_ratio=0.3, atol=1e-5, rtol=1e-5, ode_method="RK45",
                  lambda_coef=1.0, sdeis_use_order0=True),
      eval=NS(batch_size=1024, num_samples=50000),
      data=NS(dataset="CIFAR10", image_size=32, num_channels=3, centered=True, random_flip=True,
              uniform_dequantization=False),
      model=model, seed=42)


def cld_accr_dcifar10():
  """The README evaluation config (deep NCSN++, 107.6 M parameters, RK4 R-table with R_dt=1e-6)."""
  return _cld_base(_common_model(mixed_score=False, is_R_rk=True, R_dt=1e-6))


def cld_deep_cifar10():
  return _cld_base(_common_model())


def cld_ddpmpp_cifar10():
  return _cld_base(_common_model(num_res_blocks=4, fir=False, progressive_input="none",
                                 embedding_type="positional"))


def cld_simple_cifar10():
  return _cld_base(_common_model(nf=32, num_res_blocks=4, fir=False, progressive_input="none",
                                 embedding_type="positional", ema_rate=0.999))


def _blur_base(model):
  model.sigma_blur_max = getattr(model, "sigma_blur_max", 10.0)
  return NS(
      training=NS(continuous=True, batch_size=128),
      sampling=NS(method="order0", nfe=50, ts_order=2, t0=1e-5),
      eval=NS(batch_size=1024, num_samples=50000),
      data=NS(dataset="CIFAR10", image_size=32, num_channels=3, centered=True, random_flip=True,
              uniform_dequantization=False, is_partial=False),
      model=model, seed=42)


def blur_ddpm_deep_cifar10(sigma_blur_max=10.0):
  return _blur_base(_common_model(sigma_blur_max=sigma_blur_max))


def blur_simple_cifar10(sigma_blur_max=10.0):
  """blur_jax/configs/simple_cifar10_config.py:36-63 (nf = 32, same model section as the CLD one)."""
  return _blur_base(_common_model(nf=32, num_res_blocks=4, fir=False, progressive_input="none",
                                  embedding_type="positional", ema_rate=0.999, sigma_blur_max=sigma_blur_max))


def blur_ddpmpp_cifar10(sigma_blur_max=10.0):
  return _blur_base(_common_model(num_res_blocks=4, fir=False, progressive_input="none",
                                  embedding_type="positional", sigma_blur_max=sigma_blur_max))


def tiny(cfg, nf=16, num_res_blocks=1, image_size=None):
  """Shrink a config for CPU-sized parity cases (same control flow, fewer/narrower blocks)."""
  cfg.model.nf = nf
  cfg.model.num_res_blocks = num_res_blocks
  if image_size is not None:
    cfg.data.image_size = image_size
  return cfg
