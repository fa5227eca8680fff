"""Drop-in stand-ins for blur_jax/{sampling,deis,multistep,sde_lib,blur}.py."""
from . import blur, deis, multistep, sampling, sde_lib  # noqa: F401
