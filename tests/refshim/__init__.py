"""TEST INFRASTRUCTURE: a numpy/torch-CPU stand-in for the third-party packages the reference imports, so that the
reference's own source files under /root/reference/{cld_jax,blur_jax} can be imported and executed UNMODIFIED in this
container (jax 0.2.8 / flax 0.3.1 / ml_collections / jammy / tensorflow are not installable offline).

What is restated here is only the *published behaviour of the third-party API surface* the hot path touches:
  jax.numpy -> numpy (fp64, i.e. `jax_enable_x64=True`; arrays carry `.at[idx].set()`), jax.jit -> identity,
  jax.vmap / pmap -> python loop + stack, jax.lax.scan / fori_loop / cond -> python control flow,
  jax.lax.conv_general_dilated -> torch.nn.functional.conv2d (fp64), jax.lax.fft -> numpy.fft,
  jax.random -> a recorded numpy stream (NOT threefry; every normal draw is logged so fixtures can carry the noise),
  flax.linen.Module -> compact-module auto-naming (`<Class>_<k>` per parent scope, numbered at construction),
  nn.Conv / Dense / GroupNorm(epsilon=1e-6) / Dropout / avg_pool / swish, flax.struct.dataclass, jax_utils.replicate.
Everything between those calls -- the sampler loops, coefficient tables, DCT, FIR resamplers, network control flow --
is the reference's code.  tests/golden/make_ref_golden.py uses this to emit the reference-generated fixtures.

Only tests/ may import this package; /root/reference does not exist on the GPU box, so nothing here runs there.
"""
from .shim import install, reset_modules, draws, clear_draws, linen_collect  # noqa: F401
