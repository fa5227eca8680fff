"""blur_jax/deis.py is byte-identical to cld_jax/deis.py: one implementation serves both."""
from ..cld.deis import *  # noqa: F401,F403
from ..cld.deis import get_ab_eps_coef, multistep_ab_step, runge_kutta  # noqa: F401
