"""Pins the blur oracle (oracle/blur.py): DCT vs scipy, grid endpoints, and the algebraic form of the update."""
import numpy as np
import scipy.fft

from oracle import blur as ob


def test_dct_is_scipy_ortho_dctn():
  x = np.random.default_rng(0).standard_normal((2, 32, 32, 3))
  want = np.stack([scipy.fft.dctn(x[..., c], type=2, norm="ortho", axes=(1, 2)) for c in range(3)], -1)
  np.testing.assert_allclose(ob.batch_img_dct(x), want, atol=1e-12)
  np.testing.assert_allclose(ob.batch_img_idct(ob.batch_img_dct(x)), x, atol=1e-12)


def test_grid_and_sampling_T():
  sde = ob.SDE(sigma_blur_max=1.0)
  assert abs(sde.sampling_T - 0.9959798) < 1e-6
  ts = ob.get_rev_ts(sde, 2, 50)
  assert len(ts) == 51 and abs(ts[0] - sde.sampling_T) < 1e-15 and abs(ts[-1] - 1e-5) < 1e-18


def test_update_is_affine_in_y_and_eps():
  sde = ob.SDE(sigma_blur_max=1.0)
  rng = np.random.default_rng(1)
  y = rng.standard_normal((1, 32, 32, 3))
  e = rng.standard_normal((1, 32, 32, 3))
  net_fn = lambda x, lab: ob.batch_img_idct(e)          # so that DCT(net(.)) = e
  tr = []
  ob.order0_sampler(sde, net_fn, y, 5, trace=tr)
  ts = ob.get_rev_ts(sde, 2, 5)
  a = sde.y_mean_coef(ts[1]) / sde.y_mean_coef(ts[0])
  b = sde.y_std_coef(ts[1]) - a * sde.y_std_coef(ts[0])
  np.testing.assert_allclose(tr[0], a * y + b * e, atol=1e-10)


def test_scalar_ab_step():
  rng = np.random.default_rng(2)
  x = rng.standard_normal((2, 5)); e = rng.standard_normal((2, 5)); h = rng.standard_normal((2, 2, 5))
  c = rng.standard_normal(4)
  xn, hn = ob.ab_step(x, c, e, h)
  np.testing.assert_allclose(xn, c[0] * x + c[1] * e + c[2] * h[0] + c[3] * h[1], atol=1e-12)
  np.testing.assert_array_equal(hn, np.stack([e, h[0]]))
