"""threefry2x32 PRNG restatement (gddim_b200/jax_random.py) against the known answers printed in the JAX
documentation ("Pseudo random numbers in JAX": PRNGKey(0), its split and the first normals; "JAX - the sharp bits":
normal(PRNGKey(42), (3,))), plus distributional sanity of the CLD prior."""
import numpy as np

from gddim_b200 import configs, jax_random as jr
from gddim_b200.cld import sde_lib


def test_known_answers_from_the_jax_documentation():
  key = jr.PRNGKey(0)
  np.testing.assert_array_equal(key, [0, 0])
  new_key, subkey = jr.split(key)
  np.testing.assert_array_equal(new_key, [4146024105, 967050713])
  np.testing.assert_array_equal(subkey, [2718843009, 1272950319])
  np.testing.assert_allclose(jr.normal(key, (1,)), [-0.20584226], rtol=2e-7)
  np.testing.assert_allclose(jr.normal(subkey, (1,)), [-1.2515389], rtol=2e-7)
  np.testing.assert_allclose(jr.normal(jr.PRNGKey(42), (3,)), [0.18693547, -1.2806505, -1.5593132], rtol=2e-7)


def test_cld_prior_from_a_jax_key():
  sde = sde_lib.from_config(configs.cld_ddpmpp_cifar10())
  key = jr.PRNGKey(7)
  u = sde.prior_sampling(key, (1, 64, 32, 32, 3))
  assert u.shape == (1, 64, 32, 32, 3, 2) and u.dtype == np.float32
  np.testing.assert_array_equal(u, sde.prior_sampling(key, (1, 64, 32, 32, 3)))       # same key, same draw
  x, v = u[..., 0].ravel(), u[..., 1].ravel()
  assert abs(x.mean()) < 0.01 and abs(x.std() - 1.0) < 0.01
  assert abs(v.mean()) < 0.01 and abs(v.std() - 0.5) < 0.01                          # m_inv = 4
  assert abs(np.corrcoef(x, v)[0, 1]) < 0.01
  # odd sizes exercise the zero-padded counter half
  assert jr.random_bits(key, (3,)).shape == (3,) and jr.normal(key, (5, 3)).shape == (5, 3)
