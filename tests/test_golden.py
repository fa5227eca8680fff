"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the oracle):
CPU: the oracle and the library's host tables still reproduce them; GPU: the CUDA path matches them."""
import os

import numpy as np
import pytest

from conftest import rel_l2
from gddim_b200 import configs
from gddim_b200.blur import sde_lib as bsde
from gddim_b200.cld import sde_lib
from oracle import blur as ob
from oracle import cld as oc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_reproduces_table_fixtures():
  g = np.load(os.path.join(G, "cld_tables.npz"))
  sde = oc.CLD(is_R_rk=True, R_dt=1e-6)
  np.testing.assert_allclose(oc.get_rev_ts(1.0, 1e-3, 2, 49), g["rev_ts"], rtol=0, atol=0)
  np.testing.assert_allclose(sde.R(g["R_at"]), g["R"], rtol=1e-12)
  np.testing.assert_allclose(sde.get_deis_coef(2, g["rev_ts"]), g["coef_o2"], rtol=1e-12, atol=1e-16)
  b = np.load(os.path.join(G, "blur_tables.npz"))
  s = ob.SDE(sigma_blur_max=1.0)
  np.testing.assert_allclose(ob.get_rev_ts(s, 2, 50), b["rev_ts"], rtol=1e-14)
  np.testing.assert_allclose(s.y_mean_coef(b["rev_ts"][0])[..., 0], b["mean_first"], rtol=1e-13)


def test_library_tables_match_table_fixtures():
  g = np.load(os.path.join(G, "cld_tables.npz"))
  sde = sde_lib.from_config(configs.cld_accr_dcifar10())
  np.testing.assert_allclose(sde._R64(g["R_at"]), g["R"], rtol=1e-9)
  np.testing.assert_allclose(sde.get_deis_coef(2, g["rev_ts"]), g["coef_o2"], rtol=2e-6, atol=1e-9)
  np.testing.assert_allclose(sde.get_deis_coef(3, g["rev_ts"]), g["coef_o3"], rtol=2e-6, atol=1e-9)
  m, e = sde.prepare_order0_coef(oc.get_rev_ts(1.0, 1e-3, 2, 9))
  np.testing.assert_allclose(m, g["order0_mean"], rtol=2e-6, atol=1e-9)
  np.testing.assert_allclose(e, g["order0_eps"], rtol=2e-6, atol=1e-9)
  b = np.load(os.path.join(G, "blur_tables.npz"))
  s = bsde.SDE(sigma_blur_max=1.0)
  assert abs(s.sampling_T - float(b["sampling_T"])) < 1e-14
  np.testing.assert_allclose(s.y_mean_coef([b["rev_ts"][-1]])[0, ..., 0], b["mean_last"], rtol=2e-6)
  np.testing.assert_allclose(s.y_std_coef(b["rev_ts"]), b["std"], rtol=2e-6)


@pytest.mark.gpu
def test_gpu_cld_sampler_matches_golden():
  from helpers import build
  from gddim_b200.cld import sampling
  g = np.load(os.path.join(G, "cld_small_sampler.npz"))
  cfg, model, _ = build("cld_deep")
  sde = sde_lib.from_config(cfg)
  fn = sampling.get_deis_sampler(sde, model, (32, 32, 3), int(g["nfe"]), lambda x: (x + 1.) / 2., int(g["order"]),
                                 ts_order=2, denoising=True)
  x, v, _ = fn(0, model, 2, u=g["u"])
  assert rel_l2(x, g["x"]) < 1e-3 and rel_l2(v, g["v"]) < 1e-3
  from gddim_b200 import net
  eps = net.get_eps_fn(sde, model)(g["u"], np.ones(2, np.float32))
  assert rel_l2(eps, g["eps_first"]) < 2e-3


@pytest.mark.gpu
def test_gpu_blur_sampler_matches_golden():
  from helpers import build
  from gddim_b200.blur import sampling as bs
  g = np.load(os.path.join(G, "blur_small_sampler.npz"))
  cfg, model, _ = build("blur_deep")
  fn = bs.get_order0_sampler(bsde.from_config(cfg), model, (32, 32, 3), 2, int(g["nfe"]), lambda x: (x + 1.) / 2.)
  x, n = fn(0, model, 2, u=g["y"])
  assert n == int(g["nfe"]) and rel_l2(x, g["x"]) < 1e-3
