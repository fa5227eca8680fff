"""Pins the network oracle (oracle/ncsnpp.py): parameter counts of SURVEY.md 8(c)(8), FIR closed forms, and
agreement of the library's own parameter walk (names, shapes, init kinds) with the oracle's."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gddim_b200 import configs, net, params
from oracle import ncsnpp as on


@pytest.mark.parametrize("mk,cld,count", [
    (configs.cld_accr_dcifar10, True, 107_597_446),
    (configs.blur_ddpm_deep_cifar10, False, 107_587_075),
    (configs.cld_ddpmpp_cifar10, True, 61_811_334),
    (configs.cld_simple_cifar10, True, 3_883_686),
])
def test_parameter_counts(mk, cld, count):
  assert params.count(on.collect_specs(mk(), cld=cld)) == count


@pytest.mark.parametrize("mk,cld", [(configs.cld_accr_dcifar10, True), (configs.blur_ddpm_deep_cifar10, False),
                                    (configs.cld_ddpmpp_cifar10, True)])
def test_library_walk_matches_oracle_walk(mk, cld):
  cfg = mk()
  want = on.collect_specs(cfg, cld=cld)
  got = net.ScoreNet(cfg, cld=cld).specs()
  assert set(got) == set(want)
  for name, (shape, kind, scale) in want.items():
    gs, gk, gsc = got[name]
    assert tuple(gs) == tuple(shape), name
    assert gk == kind, name
    assert gsc == pytest.approx(scale, rel=1e-6), name


def test_fir_closed_forms():
  x = torch.randn(2, 5, 8, 8, dtype=torch.float64)
  k = np.outer([1, 3, 3, 1], [1, 3, 3, 1]) / 64.0
  kt = torch.as_tensor(k)[None, None].repeat(5, 1, 1, 1)
  up = on.upsample_2d(x, (1, 3, 3, 1))
  want_up = F.conv_transpose2d(x, kt * 4, stride=2, padding=1, groups=5)
  assert torch.allclose(up, want_up, atol=1e-12)
  down = on.downsample_2d(x, (1, 3, 3, 1))
  want_down = F.conv2d(x, kt, stride=2, padding=1, groups=5)
  assert torch.allclose(down, want_down, atol=1e-12)


def test_fir_tap_tables_used_by_the_kernel():
  """The per-axis tap weights hard-coded in csrc/norm.cu: down = [1,3,3,1]/8 at 2o-1..2o+2; up: even o=2m ->
  .25 x[m-1] + .75 x[m], odd o=2m+1 -> .75 x[m] + .25 x[m+1]."""
  x = torch.randn(1, 1, 8, 8, dtype=torch.float64)
  xp = F.pad(x, (2, 2, 2, 2))
  up = on.upsample_2d(x, (1, 3, 3, 1))[0, 0]
  down = on.downsample_2d(x, (1, 3, 3, 1))[0, 0]
  xs = xp[0, 0]
  g = lambda y, xx: xs[y + 2, xx + 2]
  kd = [0.125, 0.375, 0.375, 0.125]
  for oy, ox in [(0, 0), (1, 2), (3, 3)]:
    want = sum(kd[i] * kd[j] * g(2 * oy - 1 + i, 2 * ox - 1 + j) for i in range(4) for j in range(4))
    assert abs(float(down[oy, ox]) - float(want)) < 1e-12

  def taps(o):
    return ((o >> 1, 0.75), ((o >> 1) + 1, 0.25)) if o & 1 else (((o >> 1) - 1, 0.25), (o >> 1, 0.75))
  for oy, ox in [(0, 0), (1, 0), (4, 7), (15, 15), (6, 9)]:
    want = sum(wy * wx * g(iy, ix) for iy, wy in taps(oy) for ix, wx in taps(ox))
    assert abs(float(up[oy, ox]) - float(want)) < 1e-12


def test_pyramid_conv_equals_fused_formula():
  """conv_downsample_2d == FIR(pad 2) then 3x3 stride-2 VALID conv == the gather used by im2col_fir_down."""
  cfg = configs.tiny(configs.cld_accr_dcifar10(), nf=16, num_res_blocks=1)
  torch.manual_seed(0)
  x = torch.randn(1, 2, 8, 8, dtype=torch.float64)
  w = torch.randn(3, 3, 2, 4, dtype=torch.float64)

  class S:
    def child(self, _):
      return self
    def param(self, name, shape, kind, scale=1.0):
      return w if name == "weight" else torch.zeros(4, dtype=torch.float64)
  got = on._conv2d_down(S(), x, 4, cfg.model.fir_kernel)[0]
  kf = [0.125, 0.375, 0.375, 0.125]
  xp = F.pad(x, (2, 2, 2, 2))[0]
  out = torch.zeros(4, 4, 4, dtype=torch.float64)
  for oy in range(4):
    for ox in range(4):
      for ky in range(3):
        for kx in range(3):
          py, px = 2 * oy + ky, 2 * ox + kx
          f = sum(kf[i] * kf[j] * xp[:, py + i, px + j] for i in range(4) for j in range(4))   # [cin]
          out[:, oy, ox] += f @ w[ky, kx]
  assert torch.allclose(got, out, atol=1e-12)


def test_forward_runs_and_is_deterministic():
  cfg = configs.tiny(configs.cld_accr_dcifar10(), nf=16, num_res_blocks=1)
  specs = on.collect_specs(cfg)
  p = params.generate(specs, nondegenerate=True)
  x = np.random.default_rng(0).standard_normal((2, 32, 32, 6)).astype(np.float32)
  a = on.forward(p, cfg, x, 999 * 0.3)
  b = on.forward(p, cfg, x, 999 * 0.3)
  assert a.shape == x.shape and np.array_equal(a, b) and np.isfinite(a).all() and np.abs(a).max() > 1e-3
  c = on.forward(p, cfg, x, 999 * 0.3, dtype=torch.float64)
  assert np.abs(a - c).max() < 1e-3 * max(1.0, np.abs(c).max())
