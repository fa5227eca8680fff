"""Operator-level wrappers over libgddim_b200.so (torch.cuda tensors in, torch.cuda tensors out).

These are the building blocks the launch plan in csrc/unet.cpp is made of, exposed so that each kernel can
be checked against the oracle on its own:
  conv2d / nin   <-> cld_jax/models/layers.py:66-107 (ddpm_conv1x1 / ddpm_conv3x3), :467-478 (NIN)
  group_norm     <-> flax nn.GroupNorm (+ swish, + up_or_down_sampling resamplers), layerspp.py:196-213
  attention      <-> layerspp.py:74-78
"""
import ctypes as C

import numpy as np

from . import _lib


def _ptr(t):
  return None if t is None else t.data_ptr()


def pack_conv_weight(kernel_hwio, extra_1x1=None, split=False):
  """Flax HWIO kernel (kh,kw,cin,cout) [+ optional 1x1 shortcut kernel (1,1,cin2,cout)] -> K-major fp16
  [cout, kh*kw*cin (+ cin2)] with k = tap*cin + ci  (the layout csrc/unet.cpp packs).  split=True: [cout, 2K] =
  [fp16(w) | fp16(w - fp16(w))] for conv_gemm(..., wsplit=2) (precise mode)."""
  import torch
  k = torch.as_tensor(np.asarray(kernel_hwio, np.float32))
  kh, kw, cin, cout = k.shape
  w = k.reshape(kh * kw * cin, cout).t().contiguous()
  if extra_1x1 is not None:
    e = torch.as_tensor(np.asarray(extra_1x1, np.float32))
    w = torch.cat([w, e.reshape(-1, cout).t().contiguous()], dim=1)
  hi = w.to(torch.float16)
  if split:
    lo = (w - hi.float()).to(torch.float16)
    return torch.cat([hi, lo], dim=1).contiguous().cuda()
  return hi.contiguous().cuda()


def conv_gemm(a0, w, N, taps0=9, a1=None, taps1=1, bias=None, bias2=None, residual=None, rowscale=None, scale=1.0,
              out_fp32=True, out_fp16=False, impl=0, force_block_n=0, force_m_sub=0, force_cta_pairs=0, epi=0, n_store=0, a0_coff=0, a0_c=None, w_ld=None,
              w_koff=0, w_batch_stride=0, w_rows_per_batch=0, reverse=0, gn=None, wsplit=0):
  """a0 (and a1): fp16 [B,H,W,C]; w: fp16 K-major.  Returns (out32 or None, out16 or None[, row_out]).
  gn = (gamma, beta, groups, silu[, eps[, dual]]): epi 2, the GroupNorm (+ swish) of the output applied by the epilogue
  -> out16; dual = True also returns the un-normalised fp32 result (bias / residual / scale as usual) in out32."""
  import torch
  _lib.require_cuda("conv_gemm")
  B, H, W, C0 = a0.shape
  d = _lib.GemmDesc()
  d.a0, d.a0_ctot, d.a0_coff, d.a0_c, d.a0_taps = a0.data_ptr(), C0, a0_coff, (a0_c or C0), taps0
  if a1 is not None:
    d.a1, d.a1_ctot, d.a1_coff, d.a1_c, d.a1_taps = a1.data_ptr(), a1.shape[3], 0, a1.shape[3], taps1
  d.B, d.H, d.W = B, H, W
  d.w, d.N, d.w_ld, d.w_koff = w.data_ptr(), N, (w_ld or w.shape[-1]), w_koff
  d.w_batch_stride, d.w_rows_per_batch = w_batch_stride, w_rows_per_batch
  d.bias, d.bias2, d.residual, d.rowscale = _ptr(bias), _ptr(bias2), _ptr(residual), _ptr(rowscale)
  d.scale = scale
  No = n_store or N
  gn_dual = gn is not None and len(gn) > 5 and bool(gn[5])     # dual mode: fp32 linear result AND its normalised fp16 copy
  if gn is not None:
    epi, out_fp32, out_fp16 = 2, gn_dual, True
    d.gn_gamma, d.gn_beta, d.gn_groups, d.gn_silu = gn[0].data_ptr(), gn[1].data_ptr(), int(gn[2]), int(bool(gn[3]))
    d.gn_eps = float(gn[4]) if len(gn) > 4 else 1e-6
  o32 = torch.empty((B, H, W, No), dtype=torch.float32, device="cuda") if (out_fp32 and (epi == 0 or gn_dual)) else None
  o16 = torch.empty((B, H, W, N), dtype=torch.float16, device="cuda") if (out_fp16 or epi == 1) else None
  row = torch.empty((B, H, W), dtype=torch.float32, device="cuda") if epi == 1 else None
  d.out32, d.out16, d.row_out, d.ldo = _ptr(o32), _ptr(o16), _ptr(row), No
  d.n_store = n_store
  d.epi, d.impl, d.force_block_n, d.force_m_sub = epi, impl, force_block_n, force_m_sub
  d.force_cta_pairs = force_cta_pairs
  d.reverse = reverse
  d.wsplit = wsplit
  st = torch.cuda.current_stream().cuda_stream
  _lib.check(_lib.lib().gddim_conv_gemm(C.byref(d), st), "gddim_conv_gemm")
  if epi == 1:
    return o32, o16, row
  return o32, o16


def group_norm(x, gamma, beta, groups=None, silu=True, resample=0, x2=None, want_raw=False, want_norm=True,
               eps=1e-6, raw_scale=1.0, reverse=0):
  """x (and x2): fp32 [B,H,W,C].  Returns (dst16 or None, raw16 or None)."""
  import torch
  _lib.require_cuda("group_norm")
  B, H, W, C1 = x.shape
  C2 = 0 if x2 is None else x2.shape[3]
  Ct = C1 + C2
  Ho, Wo = (H // 2, W // 2) if resample in (1, 3) else ((H * 2, W * 2) if resample in (2, 4) else (H, W))
  d = _lib.NormDesc()
  d.src1, d.c1, d.src2, d.c2 = x.data_ptr(), C1, _ptr(x2), C2
  d.B, d.H, d.W = B, H, W
  d.groups = groups or min(Ct // 4, 32)
  d.gamma, d.beta, d.eps = _ptr(gamma), _ptr(beta), eps
  d.silu, d.resample = int(silu), resample
  dst = torch.empty((B, Ho, Wo, Ct), dtype=torch.float16, device="cuda") if want_norm else None
  raw = torch.empty((B, Ho, Wo, Ct), dtype=torch.float16, device="cuda") if want_raw else None
  d.dst16, d.raw16 = _ptr(dst), _ptr(raw)
  d.raw_scale = raw_scale
  d.reverse = reverse
  st = torch.cuda.current_stream().cuda_stream
  _lib.check(_lib.lib().gddim_group_norm(C.byref(d), st), "gddim_group_norm")
  return dst, raw


def attention(qkv, scale=None, reverse=0):
  """qkv: fp16 [B,H,W,3C] (q, k, v channel thirds).  Returns softmax(q k^T * scale) v as fp16 [B,H,W,C]
  (layerspp.py:74-78).  scale defaults to C^-0.5."""
  import torch
  _lib.require_cuda("attention")
  B, H, W, C3 = qkv.shape
  Cc = C3 // 3
  out = torch.empty((B, H, W, Cc), dtype=torch.float16, device="cuda")
  st = torch.cuda.current_stream().cuda_stream
  _lib.check(_lib.lib().gddim_attention(_ptr(qkv), _ptr(out), B, H * W, Cc, float(Cc) ** -0.5 if scale is None else scale,
                                        reverse, st), "gddim_attention")
  return out


def attention_proj(qkv, w3, bias3, residual, out_scale=1.0, scale=None, reverse=0):
  """AttnBlockpp from qkv to the block output in one kernel (T = 256, C = 256): returns (out32 [B,H,W,C],
  colstats [B*H*W/32, 2, C]) with out32 = (attention(qkv) @ w3^T + residual) * out_scale + bias3 * out_scale.
  w3: fp16 [C_out, C_in]."""
  import torch
  _lib.require_cuda("attention_proj")
  B, H, W, C3 = qkv.shape
  Cc = C3 // 3
  out = torch.empty((B, H, W, Cc), dtype=torch.float32, device="cuda")
  stats = torch.empty((B * H * W // 32, 2, Cc), dtype=torch.float32, device="cuda")
  st = torch.cuda.current_stream().cuda_stream
  _lib.check(_lib.lib().gddim_attention_proj(_ptr(qkv), _ptr(w3), _ptr(bias3), _ptr(residual), _ptr(out), _ptr(stats), B,
                                             H * W, Cc, float(Cc) ** -0.5 if scale is None else scale, out_scale, reverse, st),
             "gddim_attention_proj")
  return out, stats


def gn_qkv(x, gamma, beta, w, bias, groups=None, eps=1e-6, reverse=0):
  """fp16(GroupNorm(x)) @ w^T + bias in one kernel (layerspp.py:69-72).  x fp32 [B,H,W,256], w fp16 [768,256]."""
  import torch
  _lib.require_cuda("gn_qkv")
  B, H, W, Cc = x.shape
  N = w.shape[0]
  out = torch.empty((B, H, W, N), dtype=torch.float16, device="cuda")
  st = torch.cuda.current_stream().cuda_stream
  _lib.check(_lib.lib().gddim_gn_qkv(_ptr(x), _ptr(gamma), _ptr(beta), min(Cc // 4, 32) if groups is None else groups, eps,
                                     _ptr(w), _ptr(bias), _ptr(out), B, H * W, Cc, N, reverse, st), "gddim_gn_qkv")
  return out
