"""Generates the golden fixtures in this directory.

The reference (JAX 0.2.8 / Flax 0.3.1) cannot be imported in this image and ships no golden vectors
(SURVEY.md 4, 8c), so these are outputs of the *oracle* (oracle/*.py, pinned by the analytic identities in
tests/test_oracle_*.py) on seeded inputs.  They guard the oracle against drift (CPU tests) and give the GPU
parity tests a fixed target that does not depend on re-running the oracle.

  python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import build_oracle_only, oracle_blur_sample, oracle_cld_sample, prior_u  # noqa: E402
from oracle import blur as ob  # noqa: E402
from oracle import cld as oc  # noqa: E402


def main():
  # 1. CLD tables of the README evaluation setting (accr_dcifar10: RK4 R table, NFE=50)
  sde = oc.CLD(is_R_rk=True, R_dt=1e-6)
  rev50 = oc.get_rev_ts(1.0, 1e-3, 2, 49)
  np.savez_compressed(os.path.join(HERE, "cld_tables.npz"),
                      rev_ts=rev50,
                      R_at=np.array([1e-3, 0.1, 0.5, 1.0]), R=sde.R(np.array([1e-3, 0.1, 0.5, 1.0])),
                      coef_o2=sde.get_deis_coef(2, rev50), coef_o3=sde.get_deis_coef(3, rev50),
                      order0_mean=sde.prepare_order0_coef(oc.get_rev_ts(1.0, 1e-3, 2, 9))[0],
                      order0_eps=sde.prepare_order0_coef(oc.get_rev_ts(1.0, 1e-3, 2, 9))[1])
  # 2. blur tables
  b = ob.SDE(sigma_blur_max=1.0)
  rev = ob.get_rev_ts(b, 2, 50)
  np.savez_compressed(os.path.join(HERE, "blur_tables.npz"), rev_ts=rev, sampling_T=b.sampling_T,
                      mean_first=b.y_mean_coef(rev[0])[..., 0], mean_last=b.y_mean_coef(rev[-1])[..., 0],
                      std=np.array([b.y_std_coef(t) for t in rev]))
  # 3. small-network sampler outputs (nf=64, one res-block; FIR + pyramid + attention; non-degenerate init)
  cfg, net_fn = build_oracle_only("cld_deep")
  u = prior_u(2, seed=0)
  x, v, _ = oracle_cld_sample(cfg, net_fn, u, 6, 2, denoising=True)
  eps0 = oc.make_eps_fn(oc.from_config(cfg), net_fn)(u, 1.0)
  np.savez_compressed(os.path.join(HERE, "cld_small_sampler.npz"), u=u, x=x.astype(np.float32), v=v.astype(np.float32),
                      eps_first=eps0.astype(np.float32), nfe=6, order=2)
  cfg, net_fn = build_oracle_only("blur_deep")
  y = prior_u(2, seed=2, cld=False)
  xb, _ = oracle_blur_sample(cfg, net_fn, y, 6)
  np.savez_compressed(os.path.join(HERE, "blur_small_sampler.npz"), y=y, x=xb.astype(np.float32), nfe=6)
  print("golden fixtures written to", HERE)


if __name__ == "__main__":
  main()
