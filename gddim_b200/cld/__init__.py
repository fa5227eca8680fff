"""Drop-in stand-ins for cld_jax/{sampling,deis,sde_lib}.py (same names, arity and error behaviour)."""
from . import deis, sampling, sde_lib  # noqa: F401
