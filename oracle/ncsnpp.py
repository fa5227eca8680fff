"""ORACLE (test infrastructure, NOT product code) -- NCSN++/DDPM++ forward, torch CPU (fp32 or fp64).

Parity status: PINNED to outputs of the reference itself: models/ncsnpp.py + layerspp.py + layers.py +
up_or_down_sampling.py run unmodified under tests/refshim (flax.linen stand-in with compact-module auto-naming);
tests/test_ref_golden.py checks against tests/golden/ref_net.npz / ref_ops.npz / ref_blur_sampler.npz that
(i) the parameter names, shapes, initialisers and creation order of this walk equal what the reference's modules
create, (ii) the fp64 forward equals the reference's NCSNpp.__call__ to 1e-10 (FIR + pyramid + Fourier net) /
1e-5 (DDPM++: the reference builds its positional table in explicit float32), (iii) upsample_2d / downsample_2d /
conv_downsample_2d / naive_* equal the reference functions to 1e-12.  tests/test_oracle_net.py keeps the
parameter counts of SURVEY.md 8(c)(8) (107,597,446 / 107,587,075 / 61,811,334 / 3,883,686) and closed forms.

Restates (file:line relative to /root/reference/cld_jax/models):
  ncsnpp.py:41-243              NCSNpp.__call__ (control flow, skip stack, input pyramid)
  layerspp.py:33-43             GaussianFourierProjection
  layerspp.py:61-83             AttnBlockpp
  layerspp.py:115-143           Downsample (fir + with_conv -> up_or_down_sampling.Conv2d(down=True))
  layerspp.py:180-227           ResnetBlockBigGANpp
  layers.py:30-42,60-107        get_act, default_init, ddpm_conv1x1/3x3
  layers.py:450-478             get_timestep_embedding, NIN
  up_or_down_sampling.py:40-86,168-411  Conv2d, naive_*sample_2d, conv_downsample_2d, upfirdn_2d, upsample_2d, downsample_2d
Third-party behaviour restated from the published flax 0.3.x API (the package itself is not installable here; the
same statement lives once more, independently, in tests/refshim/shim.py): nn.GroupNorm(epsilon=1e-6, contiguous
channel groups, stats over H,W,C/G), nn.Conv padding SAME, auto-naming <Class>_<k> per parent scope.

Parameters are a flat dict  "ResnetBlockBigGANpp_3/Conv_0/kernel" -> array in Flax layout
(conv HWIO, dense (in,out)).  `collect_specs` runs the same walk and records
name -> (shape, kind, scale) where kind in {"vs" (variance_scaling fan_avg uniform), "zeros",
"ones", "normal"}; the numbers are drawn elsewhere (gddim_b200/params.py) so that the oracle and
the library consume the same arrays.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F


class _Scope:
  """Mimics flax.linen compact-module auto naming."""

  def __init__(self, store, prefix=""):
    self.store, self.prefix, self.counts = store, prefix, {}

  def child(self, cls):
    k = self.counts.get(cls, 0)
    self.counts[cls] = k + 1
    return _Scope(self.store, f"{self.prefix}{cls}_{k}/")

  def param(self, name, shape, kind, scale=1.0):
    return self.store.get(self.prefix + name, tuple(int(s) for s in shape), kind, scale)


class _Store:
  def __init__(self, params, dtype):
    self.params, self.dtype = params, dtype
    self.specs = OrderedDict()

  def get(self, name, shape, kind, scale):
    self.specs[name] = (shape, kind, float(scale))
    if self.params is None:                      # spec-collection mode: values are irrelevant
      return torch.zeros(shape, dtype=self.dtype)
    a = np.asarray(self.params[name])
    assert tuple(a.shape) == shape, (name, a.shape, shape)
    return torch.as_tensor(a).to(self.dtype)


def _scale0(s):
  return 1e-10 if s == 0 else s                  # layers.py:62


def swish(x):
  return x * torch.sigmoid(x)


# ---- primitive layers (x is NCHW inside the oracle) ---------------------------------------------
def _conv(scope, x, out_ch, ksize, init_scale=1.0, stride=1):
  """nn.Conv via ddpm_conv1x1/3x3, layers.py:66-107.  Param scope: the CALLER's (helper functions)."""
  s = scope.child("Conv")
  w = s.param("kernel", (ksize, ksize, x.shape[1], out_ch), "vs", _scale0(init_scale))
  b = s.param("bias", (out_ch,), "zeros")
  return F.conv2d(x, w.permute(3, 2, 0, 1), b, stride=stride, padding=ksize // 2)


def _dense(scope, x, out_ch):
  s = scope.child("Dense")
  w = s.param("kernel", (x.shape[-1], out_ch), "vs", 1.0)
  b = s.param("bias", (out_ch,), "zeros")
  return x @ w + b


def _group_norm(scope, x):
  """nn.GroupNorm(num_groups=min(C//4, 32)) -- layerspp.py:69,196,218; ncsnpp.py:236."""
  C = x.shape[1]
  s = scope.child("GroupNorm")
  g = s.param("scale", (C,), "ones")
  b = s.param("bias", (C,), "zeros")
  return F.group_norm(x, min(C // 4, 32), g, b, eps=1e-6)


def _nin(scope, x, out_ch, init_scale=0.1):
  """NIN, layers.py:467-478 (default init_scale 0.1)."""
  s = scope.child("NIN")
  w = s.param("W", (x.shape[1], out_ch), "vs", _scale0(init_scale))
  b = s.param("b", (out_ch,), "zeros")
  return torch.einsum("bchw,co->bohw", x, w) + b[None, :, None, None]


# ---- StyleGAN2 FIR resampling: literal restatement of upfirdn_2d ----------------------------------
def _setup_kernel(k):
  k = np.asarray(k, dtype=np.float64)
  if k.ndim == 1:
    k = np.outer(k, k)
  return k / np.sum(k)


def upfirdn_2d(x, k, up, down, pad0, pad1):
  """up_or_down_sampling.py:212-294, on NCHW: zero-insert, pad, VALID correlate with flipped k, decimate."""
  B, C, H, W = x.shape
  kt = torch.as_tensor(np.ascontiguousarray(k[::-1, ::-1]), dtype=x.dtype)
  z = torch.zeros(B, C, H, up, W, up, dtype=x.dtype)
  z[:, :, :, 0, :, 0] = x
  z = z.reshape(B, C, H * up, W * up)
  z = F.pad(z, (pad0, pad1, pad0, pad1))
  z = F.conv2d(z.reshape(B * C, 1, z.shape[2], z.shape[3]), kt[None, None])
  z = z.reshape(B, C, z.shape[2], z.shape[3])
  return z[:, :, ::down, ::down]


def upsample_2d(x, k, factor=2):
  k = _setup_kernel(k) * (factor ** 2)
  p = k.shape[0] - factor
  return upfirdn_2d(x, k, factor, 1, (p + 1) // 2 + factor - 1, p // 2)


def downsample_2d(x, k, factor=2):
  k = _setup_kernel(k)
  p = k.shape[0] - factor
  return upfirdn_2d(x, k, 1, factor, (p + 1) // 2, p // 2)


def naive_upsample_2d(x, factor=2):
  return x.repeat_interleave(factor, dim=2).repeat_interleave(factor, dim=3)


def naive_downsample_2d(x, factor=2):
  B, C, H, W = x.shape
  return x.reshape(B, C, H // factor, factor, W // factor, factor).mean(dim=(3, 5))


def _conv2d_down(scope, x, out_ch, fir_kernel):
  """layerspp.Downsample(fir=True, with_conv=True) -> Conv2d(down=True) -> conv_downsample_2d
  (up_or_down_sampling.py:40-73,168-209): FIR with pad (2,2), then 3x3 stride-2 VALID conv, then bias."""
  s = scope.child("Conv2d")
  w = s.param("weight", (3, 3, x.shape[1], out_ch), "vs", 1.0)
  b = s.param("bias", (out_ch,), "zeros")
  k = _setup_kernel(fir_kernel)
  p = (k.shape[0] - 2) + (3 - 1)
  x = upfirdn_2d(x, k, 1, 1, (p + 1) // 2, p // 2)
  return F.conv2d(x, w.permute(3, 2, 0, 1), None, stride=2) + b[None, :, None, None]


# ---- blocks ---------------------------------------------------------------------------------------
def _resblock(scope, cfg, x, temb, out_ch=None, up=False, down=False):
  """ResnetBlockBigGANpp, layerspp.py:180-227 (train=False: dropout is the identity)."""
  s = scope.child("ResnetBlockBigGANpp")
  m = cfg.model
  C = x.shape[1]
  out_ch = out_ch if out_ch else C
  h = swish(_group_norm(s, x))
  if up:
    if m.fir:
      h, x = upsample_2d(h, m.fir_kernel), upsample_2d(x, m.fir_kernel)
    else:
      h, x = naive_upsample_2d(h), naive_upsample_2d(x)
  elif down:
    if m.fir:
      h, x = downsample_2d(h, m.fir_kernel), downsample_2d(x, m.fir_kernel)
    else:
      h, x = naive_downsample_2d(h), naive_downsample_2d(x)
  h = _conv(s, h, out_ch, 3)
  if temb is not None:
    h = h + _dense(s, swish(temb), out_ch)[:, :, None, None]
  h = swish(_group_norm(s, h))
  h = _conv(s, h, out_ch, 3, init_scale=m.init_scale)
  if C != out_ch or up or down:
    x = _conv(s, x, out_ch, 1)
  return (x + h) / np.sqrt(2.0) if m.skip_rescale else x + h


def _attnblock(scope, cfg, x):
  """AttnBlockpp, layerspp.py:61-83."""
  s = scope.child("AttnBlockpp")
  B, C, H, W = x.shape
  h = _group_norm(s, x)
  q, k, v = _nin(s, h, C), _nin(s, h, C), _nin(s, h, C)
  w = torch.einsum("bchw,bcHW->bhwHW", q, k) * (int(C) ** (-0.5))
  w = torch.softmax(w.reshape(B, H, W, H * W), dim=-1).reshape(B, H, W, H, W)
  h = torch.einsum("bhwHW,bcHW->bchw", w, v)
  h = _nin(s, h, C, init_scale=cfg.model.init_scale)
  return (x + h) / np.sqrt(2.0) if cfg.model.skip_rescale else x + h


def _time_embedding(scope, cfg, labels, dtype):
  """ncsnpp.py:68-91."""
  m = cfg.model
  nf = m.nf
  et = m.embedding_type.lower()
  if et == "fourier":
    s = scope.child("GaussianFourierProjection")
    W = s.param("W", (nf,), "normal", m.fourier_scale)
    xp = torch.log(labels)[:, None] * W[None, :] * 2 * math.pi
    temb = torch.cat([torch.sin(xp), torch.cos(xp)], dim=-1)
  elif et == "positional":
    half = nf // 2
    e = math.log(10000) / (half - 1)
    f = torch.exp(torch.arange(half, dtype=torch.float32).to(dtype) * -e)
    a = labels[:, None] * f[None, :]
    temb = torch.cat([torch.sin(a), torch.cos(a)], dim=1)
  else:
    raise ValueError(f"embedding type {et} unknown.")
  if m.conditional:
    temb = _dense(scope, temb, nf * 4)
    temb = _dense(scope, swish(temb), nf * 4)
  else:
    temb = None
  return temb


def _forward(store, cfg, x_nhwc, labels):
  """NCSNpp.__call__, ncsnpp.py:41-243, restricted to what the shipped image configs use:
  resblock_type='biggan', progressive='none', progressive_input in {'none','residual'}, swish."""
  m = cfg.model
  assert m.resblock_type.lower() == "biggan" and m.progressive.lower() == "none"
  assert m.progressive_input.lower() in ("none", "residual") and m.nonlinearity.lower() == "swish"
  assert not m.scale_by_sigma
  top = _Scope(store)
  dtype = store.dtype
  x = torch.as_tensor(np.asarray(x_nhwc)).to(dtype).permute(0, 3, 1, 2)
  labels = torch.as_tensor(np.broadcast_to(np.asarray(labels, dtype=np.float64), (x.shape[0],)).copy()).to(dtype)
  nf, ch_mult, nrb = m.nf, tuple(m.ch_mult), m.num_res_blocks
  nres = len(ch_mult)
  temb = _time_embedding(top, cfg, labels, dtype)
  if not cfg.data.centered:
    x = 2 * x - 1.0
  pyramid = x if m.progressive_input.lower() != "none" else None

  hs = [_conv(top, x, nf, 3)]
  for lvl in range(nres):
    for _ in range(nrb):
      h = _resblock(top, cfg, hs[-1], temb, out_ch=nf * ch_mult[lvl])
      if h.shape[2] in tuple(m.attn_resolutions):
        h = _attnblock(top, cfg, h)
      hs.append(h)
    if lvl != nres - 1:
      h = _resblock(top, cfg, hs[-1], temb, down=True)
      if pyramid is not None:                                  # progressive_input == 'residual'
        assert m.fir, "non-FIR strided pyramid conv (TF SAME padding, stride 2) is used by no shipped config"
        pyramid = _conv2d_down(top.child("Downsample"), pyramid, h.shape[1], m.fir_kernel)
        pyramid = (pyramid + h) / np.sqrt(2.0) if m.skip_rescale else pyramid + h
        h = pyramid
      hs.append(h)

  h = hs[-1]
  h = _resblock(top, cfg, h, temb)
  h = _attnblock(top, cfg, h)
  h = _resblock(top, cfg, h, temb)

  for lvl in reversed(range(nres)):
    for _ in range(nrb + 1):
      h = _resblock(top, cfg, torch.cat([h, hs.pop()], dim=1), temb, out_ch=nf * ch_mult[lvl])
    if h.shape[2] in tuple(m.attn_resolutions):
      h = _attnblock(top, cfg, h)
    if lvl != 0:
      h = _resblock(top, cfg, h, temb, up=True)
  assert not hs
  h = swish(_group_norm(top, h))
  h = _conv(top, h, x.shape[1], 3, init_scale=m.init_scale)
  return h.permute(0, 2, 3, 1)


def in_channels(cfg, cld=True):
  return cfg.data.num_channels * (2 if cld else 1)


def collect_specs(cfg, cld=True):
  """Ordered name -> (shape, kind, scale), by running the walk on a batch-1 zero input."""
  store = _Store(None, torch.float32)
  S = cfg.data.image_size
  with torch.no_grad():
    _forward(store, cfg, np.zeros((1, S, S, in_channels(cfg, cld)), np.float32), 1.0)
  return store.specs


def forward(params, cfg, x_nhwc, labels, dtype=torch.float32):
  """x [B,H,W,Cin] (numpy), labels scalar or [B] (= 999 t)  ->  numpy [B,H,W,Cin] in `dtype`."""
  store = _Store(params, dtype)
  with torch.no_grad():
    return _forward(store, cfg, x_nhwc, labels).contiguous().numpy()


def make_net_fn(params, cfg, dtype=torch.float32):
  return lambda x, labels: forward(params, cfg, x, labels, dtype)
