"""Pins the CLD oracle (oracle/cld.py, oracle/cld_ode.c) with the analytic known-answer identities of
SURVEY.md 8(c) -- the reference ships no numeric fixtures ("parity unpinned")."""
import numpy as np
import pytest
import scipy.linalg

from oracle import cld as oc


@pytest.fixture(scope="module")
def sde():
  return oc.CLD()          # m_inv=4, beta=4, gamma=0.04, midpoint Euler R_dt=1e-5 (default_cifar10_config.py:81-88)


@pytest.fixture(scope="module")
def sde_rk():
  return oc.CLD(is_R_rk=True, R_dt=1e-6)     # accr_dcifar10_config.py:16-17


def test_rev_ts_grid(sde):
  ts = oc.get_rev_ts(sde.T, sde.sampling_eps, 2, 49)
  assert len(ts) == 50 and ts[0] == 1.0 and abs(ts[-1] - 1e-3) < 1e-15
  np.testing.assert_allclose(ts[1:3], [0.96086, 0.92251], atol=5e-6)
  assert np.all(np.diff(ts) < 0)


def test_psi_is_matrix_exponential(sde):
  for s, t in [(1.0, 0.9), (0.5, 0.1), (0.2, 0.19), (0.0, 1.0)]:
    B = sde.beta_int(t) - sde.beta_int(s)
    F = np.array([[0.0, sde.m_inv * B], [-B, -sde.Gamma * sde.m_inv * B]])
    np.testing.assert_allclose(sde.psi(s, t), scipy.linalg.expm(F), atol=1e-12)


def test_R_spot_values(sde_rk):
  want = {1e-3: [[8.4875e-4, 1.86379e-3], [-2.673631e-2, 1.2978354e-1]],
          0.1: [[0.11275192, -0.45727261], [0.4289476, -0.17821993]],
          0.5: [[-0.58415205, 0.80326032], [-0.40580997, -0.28853392]],
          1.0: [[0.9966196, -0.08205698], [0.04103561, 0.4983101]]}
  for t, m in want.items():
    np.testing.assert_allclose(sde_rk.R(t), np.array(m), rtol=2e-5, atol=2e-7)


def test_R_Rt_is_covariance(sde_rk):
  """Sigma' = F Sigma + Sigma F^T + G G^T integrated independently (RK4, fp64)."""
  def rhs(S, t):
    F, G = sde_rk.s_F(t), sde_rk.s_G(t)
    return F @ S + S @ F.T + G @ G.T
  S = sde_rk.R_0 @ sde_rk.R_0.T
  n, t = 20000, 0.0
  dt = 0.5 / n
  for _ in range(n):
    k1 = rhs(S, t); k2 = rhs(S + k1 * dt / 2, t + dt / 2); k3 = rhs(S + k2 * dt / 2, t + dt / 2); k4 = rhs(S + k3 * dt, t + dt)
    S = S + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    t += dt
  np.testing.assert_allclose(sde_rk.cov(0.5), S, atol=5e-9)


def test_c_scan_matches_python_scan():
  a = oc.CLD(R_dt=1e-3, use_c=True)
  b = oc.CLD(R_dt=1e-3, use_c=False)
  np.testing.assert_allclose(a._fp, b._fp, rtol=1e-12, atol=1e-15)
  a = oc.CLD(R_dt=1e-3, is_R_rk=True, use_c=True)
  b = oc.CLD(R_dt=1e-3, is_R_rk=True, use_c=False)
  np.testing.assert_allclose(a._fp, b._fp, rtol=1e-12, atol=1e-15)


def test_gddim_exactness_and_partition_of_unity(sde_rk):
  rev = oc.get_rev_ts(1.0, 1e-3, 2, 49)
  c0 = sde_rk.get_deis_coef(0, rev)
  c2 = sde_rk.get_deis_coef(2, rev)
  assert c0.shape == (49, 3, 2, 2) and c2.shape == (49, 5, 2, 2)
  np.testing.assert_allclose(c0[0, 1], [[0.02175982, 0.11881718], [-0.09453399, -0.41587199]], atol=2e-7)
  for i in (0, 10, 48):            # C_i0 = R(t_{i+1}) - Psi(t_i,t_{i+1}) R(t_i) up to the left-Riemann error
    exact = sde_rk.R(rev[i + 1]) - sde_rk.psi(rev[i], rev[i + 1]) @ sde_rk.R(rev[i])
    assert np.abs(c0[i, 1] - exact).max() < 3e-5
  np.testing.assert_allclose(c2[:, 1:].sum(axis=1), c0[:, 1], atol=1e-12)     # sum_j l_j = 1
  assert np.all(c2[:, 4] == 0)                                                 # padding row (deis.py:53)
  assert np.all(c2[0, 2:] == 0) and np.all(c2[1, 3:] == 0)                     # warm-up rows (deis.py:75)
  np.testing.assert_allclose(c2[:, 0], sde_rk.psi(rev[:-1], rev[1:]), atol=0)


@pytest.mark.parametrize("order,nfe,denoise", [(0, 10, True), (1, 12, True), (2, 20, True), (3, 50, False)])
def test_dirac_data_end_to_end(sde_rk, order, nfe, denoise):
  """With eps_theta(u,t) = R(t)^-1 (u - Psi(0,t) u0) the sampler returns ~ Psi(0,eps) u0 + R(eps) z.
  (High orders need enough steps: extrapolation feeds errors back through R^-1 on this stiff toy model.)"""
  rng = np.random.default_rng(0)
  u0 = rng.standard_normal((3, 4, 2)) * np.array([1.0, 0.0])      # data: x0 arbitrary, v0 = 0
  z = rng.standard_normal((3, 4, 2))
  uT = np.einsum("ij,...j->...i", sde_rk.psi(0.0, 1.0), u0) + np.einsum("ij,...j->...i", sde_rk.R(1.0), z)

  def eps_fn(u, t):
    return np.einsum("ij,...j->...i", sde_rk.invR(t), u - np.einsum("ij,...j->...i", sde_rk.psi(0.0, t), u0))

  x, v, n = oc.deis_sampler(sde_rk, eps_fn, uT, nfe, order, denoising=denoise, centered=False)
  te = 1e-3
  want = np.einsum("ij,...j->...i", sde_rk.psi(0.0, te), u0) + np.einsum("ij,...j->...i", sde_rk.R(te), z)
  got = np.stack([x, v], -1)
  if denoise:       # the Euler denoising step is a deterministic map of the exact state at t = eps
    want = oc.denoise_step(sde_rk, eps_fn, want)
  assert np.abs(got - want).max() < 3e-3
  assert n == nfe


def test_multistep_ab_step_matches_einsum():
  rng = np.random.default_rng(1)
  x = rng.standard_normal((2, 4, 4, 3, 2)); e = rng.standard_normal(x.shape)
  hist = rng.standard_normal((3,) + x.shape); coef = rng.standard_normal((5, 2, 2))
  xn, hn = oc.multistep_ab_step(x, coef, e, hist)
  want = np.einsum("ij,...j->...i", coef[0], x) + np.einsum("ij,...j->...i", coef[1], e)
  for j in range(3):
    want += np.einsum("ij,...j->...i", coef[2 + j], hist[j])
  np.testing.assert_allclose(xn, want, atol=1e-12)
  np.testing.assert_array_equal(hn[0], e)
  np.testing.assert_array_equal(hn[1:], hist[:2])


def test_relayout_roundtrip():
  u = np.random.default_rng(2).standard_normal((2, 4, 4, 3, 2))
  n = oc.relayout_in(u)
  assert n.shape == (2, 4, 4, 6)
  np.testing.assert_array_equal(n[..., :3], u[..., 0]); np.testing.assert_array_equal(n[..., 3:], u[..., 1])
  np.testing.assert_array_equal(oc.relayout_out(n), u)
