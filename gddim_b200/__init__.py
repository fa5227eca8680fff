"""gddim_b200: B200-native (sm_100a) implementation of the gDDIM sampling hot path.

  gddim_b200.cld.{sampling,deis,sde_lib}            stand-ins for cld_jax/{sampling,deis,sde_lib}.py
  gddim_b200.blur.{sampling,deis,multistep,sde_lib,blur}  stand-ins for the blur_jax modules
  gddim_b200.net                                     model adapter (ScoreNet handle, eps functions)
  gddim_b200.configs / params                        reference configs as data, synthetic parameters

All compute goes through libgddim_b200.so (include/gddim_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
