"""C-ABI checks that need no GPU: the library loads, exports every symbol the header declares, its fp64 host
tables agree with the oracle, and compute entry points fail loudly without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from gddim_b200 import _lib, configs
from gddim_b200.blur import sampling as bsampling
from gddim_b200.blur import sde_lib as bsde
from gddim_b200.cld import deis, sampling, sde_lib
from oracle import blur as ob
from oracle import cld as oc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported_and_bound():
  hdr = open(os.path.join(ROOT, "include", "gddim_b200.h")).read()
  hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
  declared = set(re.findall(r"\b(gddim_[A-Za-z0-9_]+)\s*\(", hdr))
  assert len(declared) >= 35
  L = _lib.lib()
  for name in declared:
    assert hasattr(L, name), f"{name} declared in include/gddim_b200.h but not exported"
  assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
  assert L.gddim_abi_version() == 1


def test_struct_sizes_match_header():
  # catches drift between the ctypes mirrors and the C structs: ints/floats/pointers only, natural alignment
  assert C.sizeof(_lib.ModelCfg) == 4 * (5 + 8 + 2 + 8 + 6)
  assert C.sizeof(_lib.SamplerCfg) == 4 * 11 + 4 + 8        # 11 ints/floats, padding to 8, one u64
  assert C.sizeof(_lib.NormDesc) % 8 == 0 and C.sizeof(_lib.GemmDesc) % 8 == 0


@pytest.fixture(scope="module")
def pair():
  cfg = configs.cld_accr_dcifar10()
  return sde_lib.from_config(cfg), oc.from_config(cfg)


def test_cld_tables_match_oracle(pair):
  lib_sde, o = pair
  ts = np.array([1e-3, 0.0123, 0.1, 0.37, 0.5, 0.99, 1.0])
  np.testing.assert_allclose(lib_sde._R64(ts), o.R(ts), rtol=1e-9, atol=1e-12)
  np.testing.assert_allclose(lib_sde._psi64(ts[:-1], ts[1:]), o.psi(ts[:-1], ts[1:]), rtol=1e-13, atol=1e-15)
  np.testing.assert_allclose(lib_sde.v_eps_integrand(ts), o.eps_integrand(ts), rtol=2e-6)
  np.testing.assert_allclose(lib_sde.s_F(0.3), o.s_F(0.3), rtol=1e-7)
  np.testing.assert_allclose(lib_sde.s_G(0.3), o.s_G(0.3), rtol=1e-7)
  assert lib_sde.T == 1.0 and lib_sde.sampling_eps == 1e-3 and lib_sde.mixed_score is False


@pytest.mark.parametrize("order,nfe", [(0, 10), (1, 20), (2, 50), (3, 50)])
def test_deis_coef_matches_oracle(pair, order, nfe):
  lib_sde, o = pair
  rev = oc.get_rev_ts(1.0, 1e-3, 2, nfe - 1)
  np.testing.assert_allclose(sampling.get_rev_ts(lib_sde, 2, nfe - 1), rev, rtol=1e-6)
  got = lib_sde.get_deis_coef(order, rev)
  want = o.get_deis_coef(order, rev)
  assert got.shape == want.shape == (nfe - 1, order + 3, 2, 2) and got.dtype == np.float32
  np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-9)
  ab = deis.get_ab_eps_coef(lib_sde, order + 1, rev, order)
  np.testing.assert_allclose(ab, want[:, 1:], rtol=2e-6, atol=1e-9)


def test_order0_coef_matches_oracle(pair):
  lib_sde, o = pair
  rev = oc.get_rev_ts(1.0, 1e-3, 2, 19)
  m, e = lib_sde.prepare_order0_coef(rev)
  mo, eo = o.prepare_order0_coef(rev)
  np.testing.assert_allclose(m, mo, rtol=2e-6, atol=1e-9)
  np.testing.assert_allclose(e, eo, rtol=2e-6, atol=1e-9)


def test_euler_R_table_variant_matches_oracle():
  cfg = configs.cld_ddpmpp_cifar10()          # is_R_rk=False, R_dt=1e-5 -> int(1/1e-5) == 99999 grid quirk
  a, b = sde_lib.from_config(cfg), oc.from_config(cfg)
  ts = np.array([1e-3, 0.2, 0.77, 1.0])
  np.testing.assert_allclose(a._R64(ts), b.R(ts), rtol=1e-9, atol=1e-12)


def test_blur_tables_match_oracle():
  a, b = bsde.SDE(sigma_blur_max=1.0), ob.SDE(sigma_blur_max=1.0)
  assert abs(a.sampling_T - b.sampling_T) < 1e-14
  np.testing.assert_allclose(bsampling.get_rev_ts(a, 2, 50), ob.get_rev_ts(b, 2, 50), rtol=1e-6)
  for t in (1e-5, 0.3, 0.9959):
    np.testing.assert_allclose(a.y_mean_coef([t])[0], b.y_mean_coef(t), rtol=2e-6)
    np.testing.assert_allclose(a.y_std_coef([t])[0], b.y_std_coef(t), rtol=2e-6)
    np.testing.assert_allclose(a.get_frequency_scaling([t])[0], b.get_frequency_scaling(t), rtol=2e-6)


def test_factory_dispatch_and_errors():
  cfg = configs.cld_accr_dcifar10()
  sde = sde_lib.from_config(cfg)
  cfg.sampling.method = "no_such_sampler"
  with pytest.raises(RuntimeError):
    sampling.get_sampling_fn(cfg, sde, None, None, lambda x: (x + 1) / 2)
  assert callable(sampling.get_order0_sampler(sde, None, (32, 32, 3), 10, None, is_em=True))
  bcfg = configs.blur_ddpm_deep_cifar10(1.0)
  bcfg.sampling.method = "deis"
  with pytest.raises(RuntimeError):
    bsampling.get_sampling_fn(bcfg, bsde.from_config(bcfg), None, None, None)
  assert sampling._affine_of(lambda x: (x + 1.) / 2.) == (0.5, 0.5, True)
  assert sampling._affine_of(lambda x: x) == (1.0, 0.0, True)
  assert sampling._affine_of(lambda x: x ** 2)[2] is False


@pytest.mark.skipif(_lib.cuda_available(), reason="checks the no-device failure mode")
def test_compute_fails_loudly_without_gpu():
  x = np.zeros((2, 4, 4, 3, 2), np.float32)
  with pytest.raises(RuntimeError, match="CUDA"):
    deis.multistep_ab_step(x, np.zeros((3, 2, 2), np.float32), x, x[None])
  from gddim_b200 import net
  cfg = configs.cld_accr_dcifar10()
  model = net.ScoreNet(cfg)
  fn = sampling.get_deis_sampler(sde_lib.from_config(cfg), model, (32, 32, 3), 10, None, 0, denoising=True)
  with pytest.raises(RuntimeError, match="CUDA"):
    fn(0, model, 2, u=np.zeros((2, 32, 32, 3, 2), np.float32))


def test_sdeis_tables_match_oracle():
  """LambdaSDE (stochastic gDDIM) coefficient tables: hat-Psi table (1e5 RK4 steps), conditional reverse covariance
  (1e4 RK4 steps per interval), order-0 and polynomial paths -- library (C++) vs the numpy oracle."""
  cfg = configs.cld_deep_cifar10()
  cfg.model.R_dt = 1e-4
  a, o = sde_lib.from_config(cfg), oc.from_config(cfg)
  rev = oc.get_rev_ts(1.0, 1e-3, 2, 3)
  lib_l, ora_l = sde_lib.LambdaSDE(a, 0.5, True), oc.LambdaSDE(o, 0.5, True)
  for order in (0, 1):
    got, want = lib_l.get_deis_coef(order, rev), ora_l.get_deis_coef(order, rev)
    assert got.shape == want.shape == (3, order + 4, 2, 2)
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6 * np.abs(want).max())
  for c in (want[0, -1], want[1, -1], np.array([[2.0, 1.0], [1.0, 3.0]]), np.zeros((2, 2))):
    np.testing.assert_allclose(sde_lib.mvn_factor_svd(c), oc.mvn_factor_svd(c), atol=1e-12)
  f = oc.mvn_factor_svd(np.array([[2.0, 1.0], [1.0, 3.0]]))
  np.testing.assert_allclose(f @ f.T, [[2.0, 1.0], [1.0, 3.0]], atol=1e-12)


def test_ldeis_and_mldeis_tables_match_oracle():
  cfg = configs.cld_deep_cifar10()
  cfg.model.R_dt = 1e-4
  a, o = sde_lib.from_config(cfg), oc.from_config(cfg)
  rev = oc.get_rev_ts(1.0, 1e-3, 2, 4)
  np.testing.assert_allclose(sde_lib.LSDE(a).get_deis_coef(2, rev), oc.LSDE(o).get_deis_coef(2, rev), rtol=3e-6, atol=1e-8)
  got = np.empty((4, 4, 2, 2))
  _lib.check(_lib.lib().gddim_cld_mldeis_coef(a._h, 1, rev.ctypes.data, 5, got.ctypes.data))
  want = oc.MLCLD(o, n=100_000).get_deis_coef(1, rev)
  np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12)
  # every sampler name of cld_jax/sampling.py:57-153 dispatches
  for name in ("order0", "deis", "sdeis", "ldeis", "hybdeis", "mldeis", "ode", "sscs", "em"):
    cfg.sampling.method = name
    assert callable(sampling.get_sampling_fn(cfg, a, None, None, lambda x: (x + 1) / 2))
