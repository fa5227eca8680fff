"""Kernel-level parity on the GPU, through the C ABI (gddim_b200.ops / deis / blur wrappers): the tcgen05
implicit-GEMM convolution (impl 0) and its CUDA-core twin (impl 1) against torch conv2d on the same
fp16-rounded operands; GroupNorm(+swish)(+FIR) against the oracle's literal upfirdn restatement; the
update / relayout / DCT kernels against the numpy oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2
from gddim_b200 import ops
from gddim_b200.blur import blur as gblur
from gddim_b200.blur import multistep as gms
from gddim_b200.cld import deis as gdeis
from oracle import blur as ob
from oracle import cld as oc
from oracle import ncsnpp as on

pytestmark = pytest.mark.gpu


def _conv_ref(a16, k_hwio, taps):
  """fp64 conv on the fp16-rounded operands; a16 [B,H,W,C] fp16 (cpu), k_hwio fp32."""
  x = a16.double().permute(0, 3, 1, 2)
  w = torch.as_tensor(k_hwio).to(torch.float16).double().permute(3, 2, 0, 1)
  return F.conv2d(x, w, padding=1 if taps == 9 else 0).permute(0, 2, 3, 1)


CASES = [  # B, H, W, Cin, Cout, two_seg
    (2, 32, 32, 128, 128, False),
    (3, 16, 16, 256, 256, False),
    (4, 8, 8, 256, 256, True),
    (16, 4, 4, 512, 256, True),
    (3, 4, 4, 256, 256, False),        # M = 48 < one tile: tail rows masked
    (1, 32, 32, 384, 128, True),
    (5, 16, 16, 64, 64, False),
    # operator ABI with channel counts that are odd multiples of 32: no 64-wide K block -> CUDA-core kernel on both
    # `impl`s (the network planner never gets here: it pairs pixels, unet.cpp pack_conv_paired)
    (3, 32, 32, 32, 32, False),
    (2, 32, 32, 96, 32, True),
    (3, 16, 16, 96, 64, True),
]


@pytest.mark.parametrize("impl", [1, 0], ids=["ref", "umma"])
@pytest.mark.parametrize("case", CASES, ids=[f"b{c[0]}_{c[1]}x{c[2]}_{c[3]}to{c[4]}{'_sc' if c[5] else ''}" for c in CASES])
def test_conv3x3_bias_residual_scale(impl, case):
  B, H, W, Cin, Cout, two = case
  g = torch.Generator().manual_seed(B * 1000 + Cin)
  a = (torch.randn(B, H, W, Cout if two else Cin, generator=g)).to(torch.float16)
  k = (torch.randn(3, 3, a.shape[3], Cout, generator=g) / np.sqrt(9 * a.shape[3])).numpy()
  bias = torch.randn(Cout, generator=g)
  bias2 = torch.randn(Cout, generator=g)
  want = _conv_ref(a, k, 9) + (bias + bias2).double()
  a1 = k1 = res = None
  if two:
    a1 = torch.randn(B, H, W, Cin, generator=g).to(torch.float16)
    k1 = (torch.randn(1, 1, Cin, Cout, generator=g) / np.sqrt(Cin)).numpy()
    want = want + _conv_ref(a1, k1, 1)
  else:
    res = torch.randn(B, H, W, Cout, generator=g)
    want = want + res.double()
  want = want / np.sqrt(2.0)
  w = ops.pack_conv_weight(k, k1)
  o32, o16 = ops.conv_gemm(a.cuda(), w, Cout, taps0=9, a1=None if a1 is None else a1.cuda(), bias=bias.cuda(),
                           bias2=bias2.cuda(), residual=None if res is None else res.cuda(),
                           scale=float(1 / np.sqrt(2.0)), out_fp32=True, out_fp16=True, impl=impl)
  torch.cuda.synchronize()
  e32 = rel_l2(o32.cpu().numpy(), want.numpy())
  print(f"conv {case} impl={impl}: fp32-out rel_l2={e32:.2e}")
  assert e32 < 2e-5
  assert rel_l2(o16.float().cpu().numpy(), want.numpy()) < 1e-3


@pytest.mark.parametrize("impl", [1, 0], ids=["ref", "umma"])
@pytest.mark.parametrize("bn", [0, 32, 64, 128, 256])
def test_gemm_block_n_variants(impl, bn):
  if impl == 1 and bn != 0:
    pytest.skip("block_n only exists on the tcgen05 path")
  g = torch.Generator().manual_seed(7 + bn)
  a = torch.randn(2, 16, 16, 128, generator=g).to(torch.float16)
  k = (torch.randn(1, 1, 128, 256, generator=g) / np.sqrt(128)).numpy()
  want = _conv_ref(a, k, 1)
  o32, _ = ops.conv_gemm(a.cuda(), ops.pack_conv_weight(k), 256, taps0=1, impl=impl, force_block_n=bn)
  assert rel_l2(o32.cpu().numpy(), want.numpy()) < 2e-5


@pytest.mark.parametrize("bn", [64, 128])
@pytest.mark.parametrize("B", [2, 3, 5])
def test_umma_256_row_tiles(bn, B):
  """m_sub = 2: one weight tile feeds two 128-row accumulators (odd B: the last CTA tile is half out of range)."""
  g = torch.Generator().manual_seed(40 + bn + B)
  a = torch.randn(B, 8, 16, 128, generator=g).to(torch.float16)       # M = B*128
  k = (torch.randn(3, 3, 128, 128, generator=g) / np.sqrt(9 * 128)).numpy()
  res = torch.randn(B, 8, 16, 128, generator=g)
  want = (_conv_ref(a, k, 9) + res.double()) * 0.5
  o32, _ = ops.conv_gemm(a.cuda(), ops.pack_conv_weight(k), 128, taps0=9, residual=res.cuda(), scale=0.5, impl=0,
                         force_block_n=bn, force_m_sub=2)
  assert rel_l2(o32.cpu().numpy(), want.numpy()) < 2e-5


@pytest.mark.parametrize("B", [2, 3, 9])
@pytest.mark.parametrize("two_seg", [False, True])
def test_umma_cta_pairs(B, two_seg):
  """cta_group::2: a cluster of two CTAs issues M = 256 MMAs, each CTA staging its own 128 rows and half of the
  weight tile.  Odd B leaves the follower CTA of the last pair without rows (fully out-of-range tile); B = 9 gives
  more pair tiles than one wave on small grids.  The heuristic only picks pairs for large layers, so force them."""
  g = torch.Generator().manual_seed(90 + B + int(two_seg))
  a = torch.randn(B, 8, 16, 128, generator=g).to(torch.float16)       # M = B*128 -> B CTA tiles
  k = (torch.randn(3, 3, 128, 256, generator=g) / np.sqrt(9 * 128)).numpy()
  a1 = torch.randn(B, 8, 16, 64, generator=g).to(torch.float16) if two_seg else None
  k1 = (torch.randn(1, 1, 64, 256, generator=g) / 8).numpy() if two_seg else None
  res = torch.randn(B, 8, 16, 256, generator=g)
  bias = torch.randn(256, generator=g)
  want = _conv_ref(a, k, 9)
  if two_seg:
    want = want + _conv_ref(a1, k1, 1)
  want = (want + bias.double() + res.double()) * 0.5
  o32, o16 = ops.conv_gemm(a.cuda(), ops.pack_conv_weight(k, k1), 256, taps0=9, a1=None if a1 is None else a1.cuda(),
                           bias=bias.cuda(), residual=res.cuda(), scale=0.5, out_fp16=True, impl=0, force_block_n=256,
                           force_cta_pairs=2)
  assert rel_l2(o32.cpu().numpy(), want.numpy()) < 2e-5
  assert rel_l2(o16.float().cpu().numpy(), want.numpy()) < 1e-3
  single, _ = ops.conv_gemm(a.cuda(), ops.pack_conv_weight(k, k1), 256, taps0=9, a1=None if a1 is None else a1.cuda(),
                            bias=bias.cuda(), residual=res.cuda(), scale=0.5, impl=0, force_block_n=256,
                            force_cta_pairs=1)
  # same operands, fp32 accumulation; only the order of the taps differs (the pair path uses halo tiles here)
  assert rel_l2(single.cpu().numpy(), o32.cpu().numpy()) < 1e-6


HALO_CASES = [  # B, H, W, Cin, Cout, C1 (1x1 shortcut segment), block_n, m_sub, pairs
    (2, 32, 32, 128, 128, 0, 128, 2, 1), (3, 32, 32, 128, 128, 256, 128, 2, 1), (3, 32, 32, 128, 128, 0, 128, 2, 2),
    (2, 16, 16, 256, 256, 0, 256, 1, 2), (3, 16, 16, 256, 256, 128, 256, 1, 2), (1, 32, 32, 64, 256, 0, 256, 1, 2),
    (2, 64, 64, 64, 128, 0, 128, 2, 1), (1, 16, 16, 128, 256, 0, 256, 1, 2)]


@pytest.mark.parametrize("case", HALO_CASES, ids=[f"b{c[0]}_{c[1]}x{c[2]}_{c[3]}to{c[4]}+{c[5]}_bn{c[6]}x{c[7]}_cg{c[8]}" for c in HALO_CASES])
def test_umma_halo_tiles(case):
  """3x3 convolutions whose CTA tile is a block of whole image rows: one (rows + 2)-row TMA box per x-shift, the three
  y-shifts are descriptor offsets into it (incl. CTA pairs, a 1x1 shortcut segment, odd tile counts, image borders)."""
  B, H, W, Cin, Cout, C1, bn, ms, cg = case
  g = torch.Generator().manual_seed(sum(case))
  a = torch.randn(B, H, W, Cin, generator=g).to(torch.float16)
  k = (torch.randn(3, 3, Cin, Cout, generator=g) / np.sqrt(9 * Cin)).numpy()
  a1 = torch.randn(B, H, W, C1, generator=g).to(torch.float16) if C1 else None
  k1 = (torch.randn(1, 1, C1, Cout, generator=g) / np.sqrt(C1)).numpy() if C1 else None
  res = torch.randn(B, H, W, Cout, generator=g)
  bias = torch.randn(Cout, generator=g)
  want = _conv_ref(a, k, 9)
  if C1:
    want = want + _conv_ref(a1, k1, 1)
  want = (want + bias.double() + res.double()) * 0.5
  o32, o16 = ops.conv_gemm(a.cuda(), ops.pack_conv_weight(k, k1), Cout, taps0=9, a1=None if a1 is None else a1.cuda(),
                           bias=bias.cuda(), residual=res.cuda(), scale=0.5, out_fp16=True, impl=0, force_block_n=bn,
                           force_m_sub=ms, force_cta_pairs=cg)
  e = rel_l2(o32.cpu().numpy(), want.numpy())
  print(f"halo {case}: {e:.2e}")
  assert e < 2e-5
  assert rel_l2(o16.float().cpu().numpy(), want.numpy()) < 1e-3


@pytest.mark.parametrize("impl", [1, 0], ids=["ref", "umma"])
def test_attention_chain(impl):
  """qk^T -> row softmax epilogue -> P V with per-image B operands == softmax(q k^T / sqrt(C)) v."""
  B, H, W, Cc = 3, 16, 16, 256
  T = H * W
  g = torch.Generator().manual_seed(11)
  qkv = (torch.randn(B, H, W, 3 * Cc, generator=g) * 0.5).to(torch.float16)
  qd = qkv.cuda()
  scale = float(Cc) ** -0.5
  _, p16, rowinv = ops.conv_gemm(qd, qd, T, taps0=1, a0_coff=0, a0_c=Cc, w_ld=3 * Cc, w_koff=Cc,
                                 w_batch_stride=T * 3 * Cc, w_rows_per_batch=T, scale=scale, epi=1, impl=impl)
  q, k, v = qkv[..., :Cc].double().reshape(B, T, Cc), qkv[..., Cc:2 * Cc].double().reshape(B, T, Cc), \
      qkv[..., 2 * Cc:].double().reshape(B, T, Cc)
  s = torch.einsum("btc,bsc->bts", q, k) * scale
  p_want = torch.softmax(s, dim=-1)
  p_got = p16.float().reshape(B, T, T).cpu().double() * rowinv.reshape(B, T, 1).cpu().double()
  assert rel_l2(p_got.numpy(), p_want.numpy()) < 2e-3
  vT = v.transpose(1, 2).contiguous().to(torch.float16).cuda()             # [B, C, T]
  _, o16 = ops.conv_gemm(p16, vT, Cc, taps0=1, w_ld=T, w_batch_stride=Cc * T, w_rows_per_batch=Cc,
                         rowscale=rowinv, out_fp32=False, out_fp16=True, impl=impl)
  o_want = torch.einsum("bts,bsc->btc", p_want, v)
  assert rel_l2(o16.float().reshape(B, T, Cc).cpu().numpy(), o_want.numpy()) < 3e-3


@pytest.mark.parametrize("B", [1, 3, 8])
@pytest.mark.parametrize("reverse", [0, 1])
def test_fused_attention_256(B, reverse):
  """attn256_kernel (QK^T -> softmax -> P.V, V as an MN-major operand) == softmax(q k^T / sqrt(C)) v, and agrees with
  the unfused GEMM chain it replaces."""
  H = W = 16; Cc = 256; T = H * W
  g = torch.Generator().manual_seed(23 + B)
  qkv = (torch.randn(B, H, W, 3 * Cc, generator=g) * 0.5).to(torch.float16)
  qkv[..., :Cc] *= 2.0                                      # peaked rows as well as flat ones
  qd = qkv.cuda()
  o = ops.attention(qd, reverse=reverse)
  q, k, v = [qkv[..., i * Cc:(i + 1) * Cc].double().reshape(B, T, Cc) for i in range(3)]
  want = torch.einsum("bts,bsc->btc", torch.softmax(torch.einsum("btc,bsc->bts", q, k) * Cc ** -0.5, dim=-1), v)
  e = rel_l2(o.float().reshape(B, T, Cc).cpu().numpy(), want.numpy())
  print(f"fused attention B={B}: {e:.2e}")
  assert e < 2e-3
  scale = float(Cc) ** -0.5
  _, p16, rowinv = ops.conv_gemm(qd, qd, T, taps0=1, a0_coff=0, a0_c=Cc, w_ld=3 * Cc, w_koff=Cc,
                                 w_batch_stride=T * 3 * Cc, w_rows_per_batch=T, scale=scale, epi=1, impl=0)
  vT = qd[..., 2 * Cc:].reshape(B, T, Cc).transpose(1, 2).contiguous()
  _, o_chain = ops.conv_gemm(p16, vT, Cc, taps0=1, w_ld=T, w_batch_stride=Cc * T, w_rows_per_batch=Cc,
                             rowscale=rowinv, out_fp32=False, out_fp16=True, impl=0)
  assert rel_l2(o.float().cpu().numpy(), o_chain.float().cpu().numpy()) < 5e-4


@pytest.mark.parametrize("B,reverse", [(2, 0), (5, 1)])
def test_fused_attention_projection(B, reverse):
  """attn256_kernel<PROJ>: attention + NIN_3 + residual + 1/sqrt(2) + column statistics in one kernel == the fused
  attention kernel followed by the projection GEMM, bit for bit (it runs the same epilogue code)."""
  H = W = 16; Cc = 256; T = H * W
  g = torch.Generator().manual_seed(31 + B)
  qkv = (torch.randn(B, H, W, 3 * Cc, generator=g) * 0.5).to(torch.float16).cuda()
  k3 = (torch.randn(1, 1, Cc, Cc, generator=g) / np.sqrt(Cc)).numpy()
  w3 = ops.pack_conv_weight(k3)                               # fp16 [C_out, C_in]
  b3 = torch.randn(Cc, generator=g).cuda()
  x = torch.randn(B, H, W, Cc, generator=g).cuda()
  sc = float(1.0 / np.sqrt(2.0))
  out, stats = ops.attention_proj(qkv, w3, b3, x, out_scale=sc, reverse=reverse)
  o16 = ops.attention(qkv)
  want, _ = ops.conv_gemm(o16, w3, Cc, taps0=1, bias=b3, residual=x, scale=sc, impl=0)
  assert torch.equal(out, want)
  slabs = out.reshape(B * T // 32, 32, Cc).double()
  assert rel_l2(stats[:, 0].cpu().numpy(), slabs.sum(1).cpu().numpy()) < 1e-6
  assert rel_l2(stats[:, 1].cpu().numpy(), (slabs * slabs).sum(1).cpu().numpy()) < 1e-6
  # and against the definition
  q, k, v = [qkv[..., i * Cc:(i + 1) * Cc].double().reshape(B, T, Cc).cpu() for i in range(3)]
  h = torch.einsum("bts,bsc->btc", torch.softmax(torch.einsum("btc,bsc->bts", q, k) * Cc ** -0.5, dim=-1), v)
  ref = (h @ torch.as_tensor(k3[0, 0]).to(torch.float16).double() + b3.cpu().double() + x.reshape(B, T, Cc).cpu().double()) * sc
  assert rel_l2(out.reshape(B, T, Cc).cpu().numpy(), ref.numpy()) < 1e-3


@pytest.mark.parametrize("B,reverse,H", [(1, 0, 16), (3, 1, 16), (8, 0, 16), (2, 1, 32),
                                         # more tiles than SMs: the persistent kernel's CTAs walk 2..5 tiles each
                                         (150, 0, 16), (77, 1, 16), (45, 0, 32)])
def test_gn_fused_into_qkv_projection(B, reverse, H):
  """gn_qkv_kernel: GroupNorm apply in the A-operand path of the q/k/v projection == gn_apply followed by the N = 768
  GEMM, bit for bit (same x*a+b, same fp16 rounding, same accumulation order, same epilogue).  H = 32: the 1024-token
  attention of the 256x256 configuration."""
  W = H; Cc = 256
  g = torch.Generator().manual_seed(41 + B)
  x = (torch.randn(B, H, W, Cc, generator=g) * 1.7 + 0.2).cuda()
  gamma, beta = (1 + 0.1 * torch.randn(Cc, generator=g)).cuda(), (0.1 * torch.randn(Cc, generator=g)).cuda()
  k = (torch.randn(1, 1, Cc, 3 * Cc, generator=g) / np.sqrt(Cc)).numpy()
  w = ops.pack_conv_weight(k)
  bias = torch.randn(3 * Cc, generator=g).cuda()
  got = ops.gn_qkv(x, gamma, beta, w, bias, reverse=reverse)
  h16, _ = ops.group_norm(x, gamma, beta, silu=False)
  _, want = ops.conv_gemm(h16, w, 3 * Cc, taps0=1, bias=bias, out_fp32=False, out_fp16=True, impl=0)
  assert torch.equal(got, want)
  hn = F.group_norm(x.double().permute(0, 3, 1, 2), 32, gamma.double(), beta.double(), eps=1e-6).permute(0, 2, 3, 1)
  ref = hn.cpu() @ torch.as_tensor(k[0, 0]).double() + bias.cpu().double()
  assert rel_l2(got.float().cpu().numpy(), ref.numpy()) < 1e-3


def test_attention_small_and_unsupported():
  g = torch.Generator().manual_seed(5)
  qkv = (torch.randn(2, 4, 4, 3 * 256, generator=g) * 0.5).to(torch.float16)
  o = ops.attention(qkv.cuda())
  q, k, v = [qkv[..., i * 256:(i + 1) * 256].double().reshape(2, 16, 256) for i in range(3)]
  want = torch.einsum("bts,bsc->btc", torch.softmax(torch.einsum("btc,bsc->bts", q, k) / 16.0, dim=-1), v)
  assert rel_l2(o.float().reshape(2, 16, 256).cpu().numpy(), want.numpy()) < 2e-3
  with pytest.raises(RuntimeError):
    ops.attention(torch.zeros(1, 16, 8, 3 * 256, dtype=torch.float16, device="cuda"))


NORM_CASES = [(2, 32, 32, 128, 0, 0), (2, 16, 16, 256, 128, 0), (2, 16, 16, 128, 64, 0), (3, 8, 8, 256, 0, 1),
              (3, 8, 8, 256, 0, 2), (2, 16, 16, 128, 0, 3), (2, 4, 4, 256, 0, 4), (5, 4, 4, 256, 256, 0),
              (2, 32, 32, 64, 0, 1),
              # single-kernel path for images of <= 64 pixels (8x8 / 4x4 levels), incl. concatenated sources
              (3, 8, 8, 256, 0, 0), (2, 8, 8, 256, 256, 0), (3, 4, 4, 256, 0, 0), (2, 8, 8, 128, 0, 0), (2, 4, 4, 64, 64, 0),
              (2, 8, 8, 192, 64, 0), (2, 2, 2, 256, 0, 0),
              # nf = 32 networks: 8 / 16 / 24 groups of four channels
              (2, 32, 32, 32, 0, 0), (2, 32, 32, 32, 32, 0), (3, 32, 32, 64, 32, 0), (2, 32, 32, 32, 0, 3), (2, 16, 16, 64, 0, 4),
              (2, 16, 16, 64, 64, 0), (3, 8, 8, 64, 0, 0), (2, 4, 4, 64, 64, 0)]


@pytest.mark.parametrize("case", NORM_CASES, ids=[f"b{c[0]}_{c[1]}_{c[3]}+{c[4]}_rs{c[5]}" for c in NORM_CASES])
@pytest.mark.parametrize("silu", [True, False])
def test_group_norm_swish_resample(case, silu):
  B, H, W, C1, C2, rs = case
  g = torch.Generator().manual_seed(H * C1 + rs)
  x1 = torch.randn(B, H, W, C1, generator=g) * 2 + 0.3
  x2 = torch.randn(B, H, W, C2, generator=g) if C2 else None
  Ct = C1 + C2
  gamma, beta = 1 + 0.1 * torch.randn(Ct, generator=g), 0.1 * torch.randn(Ct, generator=g)
  dst, raw = ops.group_norm(x1.cuda(), gamma.cuda(), beta.cuda(), silu=silu, resample=rs,
                            x2=None if x2 is None else x2.cuda(), want_raw=True)
  x = (x1 if x2 is None else torch.cat([x1, x2], -1)).double().permute(0, 3, 1, 2)
  h = F.group_norm(x, min(Ct // 4, 32), gamma.double(), beta.double(), eps=1e-6)
  if silu:
    h = on.swish(h)
  fn = {0: lambda t: t, 1: lambda t: on.downsample_2d(t, (1, 3, 3, 1)), 2: lambda t: on.upsample_2d(t, (1, 3, 3, 1)),
        3: on.naive_downsample_2d, 4: on.naive_upsample_2d}[rs]
  want_n, want_r = fn(h).permute(0, 2, 3, 1), fn(x).permute(0, 2, 3, 1)
  assert dst.shape == want_n.shape
  assert rel_l2(dst.float().cpu().numpy(), want_n.numpy()) < 6e-4       # fp16 output rounding ~ 2^-11 / sqrt(3)
  assert rel_l2(raw.float().cpu().numpy(), want_r.numpy()) < 6e-4


REV_CASES = [  # B, H, W, Cin, Cout, block_n, m_sub, pairs
    (3, 32, 32, 128, 128, 128, 2, 2), (5, 16, 16, 256, 256, 256, 1, 2), (3, 16, 16, 128, 256, 256, 1, 1),
    (7, 8, 8, 64, 64, 64, 1, 1), (9, 8, 16, 128, 128, 128, 2, 1)]


@pytest.mark.parametrize("case", REV_CASES, ids=[f"b{c[0]}_{c[1]}x{c[2]}_{c[3]}to{c[4]}_bn{c[5]}x{c[6]}_cg{c[7]}" for c in REV_CASES])
def test_umma_reverse_tile_order_is_bit_identical(case):
  """`reverse` only changes the order in which a CTA visits its tiles (L2 reuse along producer -> consumer chains,
  unet.cpp finalize): the outputs must be bit-identical, including ragged last tiles, pairs and halo tiles."""
  B, H, W, Cin, Cout, bn, ms, cg = case
  g = torch.Generator().manual_seed(7 + sum(case))
  a = torch.randn(B, H, W, Cin, generator=g).to(torch.float16).cuda()
  k = (torch.randn(3, 3, Cin, Cout, generator=g) / np.sqrt(9 * Cin)).numpy()
  res = torch.randn(B, H, W, Cout, generator=g).cuda()
  bias = torch.randn(Cout, generator=g).cuda()
  w = ops.pack_conv_weight(k)
  outs = [ops.conv_gemm(a, w, Cout, taps0=9, bias=bias, residual=res, scale=0.5, out_fp16=True, impl=0,
                        force_block_n=bn, force_m_sub=ms, force_cta_pairs=cg, reverse=r) for r in (0, 1)]
  assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
  want = (_conv_ref(a.cpu(), k, 9) + bias.cpu().double() + res.cpu().double()) * 0.5
  assert rel_l2(outs[1][0].cpu().numpy(), want.numpy()) < 2e-5


@pytest.mark.parametrize("case", [(5, 32, 32, 128, 0, 0), (3, 16, 16, 256, 128, 1), (4, 8, 8, 256, 0, 0), (3, 16, 16, 64, 0, 2)],
                         ids=lambda c: f"b{c[0]}_{c[1]}_{c[3]}+{c[4]}_rs{c[5]}")
def test_group_norm_reverse_sweep_is_bit_identical(case):
  B, H, W, C1, C2, rs = case
  g = torch.Generator().manual_seed(11 + sum(case))
  x1 = (torch.randn(B, H, W, C1, generator=g) * 2 + 0.3).cuda()
  x2 = torch.randn(B, H, W, C2, generator=g).cuda() if C2 else None
  Ct = C1 + C2
  gamma, beta = (1 + 0.1 * torch.randn(Ct, generator=g)).cuda(), (0.1 * torch.randn(Ct, generator=g)).cuda()
  outs = [ops.group_norm(x1, gamma, beta, silu=True, resample=rs, x2=x2, want_raw=True, reverse=r) for r in (0, 1)]
  assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("order", [0, 1, 2, 3])
@pytest.mark.parametrize("as_numpy", [True, False])
def test_multistep_ab_step(order, as_numpy):
  rng = np.random.default_rng(order)
  x = rng.standard_normal((3, 8, 8, 3, 2)).astype(np.float32)
  e = rng.standard_normal(x.shape).astype(np.float32)
  hist = rng.standard_normal((order + 1,) + x.shape).astype(np.float32)
  coef = rng.standard_normal((order + 3, 2, 2)).astype(np.float32)
  wx, wh = oc.multistep_ab_step(x.astype(np.float64), coef.astype(np.float64), e.astype(np.float64), hist.astype(np.float64))
  if as_numpy:
    gx, gh = gdeis.multistep_ab_step(x, coef, e, hist)
  else:
    gx, gh = gdeis.multistep_ab_step(torch.as_tensor(x).cuda(), coef, torch.as_tensor(e).cuda(), torch.as_tensor(hist).cuda())
    gx, gh = gx.cpu().numpy(), gh.cpu().numpy()
  np.testing.assert_allclose(gx, wx, rtol=0, atol=2e-5)
  np.testing.assert_array_equal(gh, wh.astype(np.float32))           # history shift is a pure copy: bit exact


def test_scalar_ab_step():
  rng = np.random.default_rng(5)
  x = rng.standard_normal((4, 32, 32, 3)).astype(np.float32)
  e = rng.standard_normal(x.shape).astype(np.float32)
  h = rng.standard_normal((2,) + x.shape).astype(np.float32)
  c = rng.standard_normal(4).astype(np.float32)
  wx, wh = ob.ab_step(x.astype(np.float64), c.astype(np.float64), e.astype(np.float64), h.astype(np.float64))
  gx, gh = gms.ab_step(x, c, e, h)
  np.testing.assert_allclose(gx, wx, atol=1e-5)
  np.testing.assert_array_equal(gh, wh.astype(np.float32))


@pytest.mark.parametrize("C", [1, 3])
def test_dct_matches_oracle_and_roundtrips(C):
  x = np.random.default_rng(C).standard_normal((5, 32, 32, C)).astype(np.float32)
  y = gblur.batch_img_dct(x)
  np.testing.assert_allclose(y, ob.batch_img_dct(x.astype(np.float64)), atol=2e-5)
  np.testing.assert_allclose(gblur.batch_img_idct(y), x, atol=2e-5)
  np.testing.assert_allclose(gblur.batch_img_idct(x), ob.batch_img_idct(x.astype(np.float64)), atol=2e-5)


@pytest.mark.parametrize("impl", [1, 0], ids=["ref", "umma"])
def test_few_channel_output_umma_or_ref(impl):
  """Head conv: C_out = 6 padded to a 32-wide N tile, only 6 columns stored densely ([.., 6] fp32)."""
  g = torch.Generator().manual_seed(77)
  a = torch.randn(3, 32, 32, 128, generator=g).to(torch.float16)
  k = (torch.randn(3, 3, 128, 6, generator=g) / np.sqrt(9 * 128)).numpy()
  bias = torch.randn(6, generator=g)
  want = _conv_ref(a, k, 9) + bias.double()
  kp = np.zeros((3, 3, 128, 32), np.float32); kp[..., :6] = k
  bp = torch.zeros(32); bp[:6] = bias
  o32, _ = ops.conv_gemm(a.cuda(), ops.pack_conv_weight(kp), 32, taps0=9, bias=bp.cuda(), impl=impl, n_store=6)
  assert o32.shape == (3, 32, 32, 6)
  assert rel_l2(o32.cpu().numpy(), want.numpy()) < 2e-5


GNF_CASES = [  # B, H, W, Cin, N, taps, force_block_n, force_m_sub, force_pairs     (what it exercises)
    (3, 16, 16, 128, 256, 9, 0, 0, 0),      # image = 2 CTAs: cluster of two single-CTA MMAs (DSMEM exchange)
    (4, 16, 16, 128, 256, 9, 256, 1, 2),    # image = one cta_group::2 pair
    (2, 32, 32, 64, 128, 9, 0, 0, 0),       # image = 4 CTAs, 256-row halo tiles
    (3, 32, 32, 64, 64, 9, 0, 0, 0),        # image = 4 CTAs, plain 256-row tiles, 4 channels per group
    (5, 8, 8, 128, 256, 9, 0, 0, 0),        # two images per tile, ragged last tile
    (3, 4, 4, 256, 256, 9, 0, 0, 0),        # 16-row images: a warp's 32 rows split in two, M = 48 < one tile
    (9, 4, 4, 128, 128, 9, 128, 2, 0),      # 256-row tiles of 16 images, 4 channels per group
    (2, 16, 16, 64, 128, 9, 128, 2, 0),     # tile = exactly one image
    (2, 16, 16, 64, 128, 1, 128, 1, 0),     # 1x1 convolution, image = 2 CTAs
    (6, 8, 8, 256, 256, 9, 64, 1, 0),       # narrow N tiles: groups never straddle tiles
    (40, 16, 16, 128, 256, 9, 0, 0, 0),     # more tiles than SMs: persistent loop, exchange parity alternates
    (37, 32, 32, 128, 128, 9, 0, 0, 0),     # more 4-clusters than fit at once
]


@pytest.mark.parametrize("impl", [1, 0], ids=["ref", "umma"])
@pytest.mark.parametrize("silu", [True, False])
@pytest.mark.parametrize("case", GNF_CASES, ids=[f"b{c[0]}_{c[1]}x{c[2]}_{c[3]}to{c[4]}_t{c[5]}_bn{c[6]}x{c[7]}_cg{c[8]}" for c in GNF_CASES])
def test_conv_with_groupnorm_epilogue(impl, silu, case):
  """epi = 2: out16 = act(GroupNorm(conv(a) + bias + bias2)) in ONE kernel == the reference order conv -> GroupNorm_1 ->
  swish of layerspp.py:214-219 (torch fp64 on the same fp16 operands; flax eps 1e-6, min(C/4, 32) groups)."""
  B, H, W, Cin, N, taps, bn, ms, pairs = case
  if impl == 1 and (bn or B > 9):
    pytest.skip("tile variants do not exist on the CUDA-core path")
  g = torch.Generator().manual_seed(B * 131 + H + N)
  a = torch.randn(B, H, W, Cin, generator=g).to(torch.float16)
  kk = 3 if taps == 9 else 1
  k = (torch.randn(kk, kk, Cin, N, generator=g) / np.sqrt(taps * Cin)).numpy()
  bias, bias2 = torch.randn(N, generator=g), 0.5 * torch.randn(N, generator=g)
  gamma, beta = 1 + 0.2 * torch.randn(N, generator=g), 0.3 * torch.randn(N, generator=g)
  groups = min(N // 4, 32)
  assert ops._lib.lib().gddim_gemm_gnf_supported(H, W, N, groups) == 1
  v = _conv_ref(a, k, taps) + (bias + bias2).double()
  # per-image offsets so that the statistics really differ between the images of one tile
  want = F.group_norm(v.permute(0, 3, 1, 2), groups, gamma.double(), beta.double(), eps=1e-6)
  if silu:
    want = on.swish(want)
  want = want.permute(0, 2, 3, 1)
  _, o16 = ops.conv_gemm(a.cuda(), ops.pack_conv_weight(k), N, taps0=taps, bias=bias.cuda(), bias2=bias2.cuda(), impl=impl,
                         force_block_n=bn, force_m_sub=ms, force_cta_pairs=pairs, gn=(gamma.cuda(), beta.cuda(), groups, silu))
  torch.cuda.synchronize()
  e = rel_l2(o16.float().cpu().numpy(), want.numpy())
  print(f"gnf {case} impl={impl} silu={silu}: rel_l2={e:.2e}")
  assert e < 6e-4                                    # fp16 output rounding ~ 2^-11 / sqrt(3)
  # per image: a wrong statistic (mixed-up images inside a tile) shows up as a per-image error, not in the global norm
  per = [rel_l2(o16[b].float().cpu().numpy(), want[b].numpy()) for b in range(B)]
  assert max(per) < 8e-4, per


def test_groupnorm_epilogue_is_deterministic_and_reverse_invariant():
  B, H, W, Cin, N = 6, 16, 16, 128, 256
  g = torch.Generator().manual_seed(7)
  a = torch.randn(B, H, W, Cin, generator=g).to(torch.float16).cuda()
  w = ops.pack_conv_weight((torch.randn(3, 3, Cin, N, generator=g) / np.sqrt(9 * Cin)).numpy())
  gamma, beta = (1 + 0.2 * torch.randn(N, generator=g)).cuda(), (0.3 * torch.randn(N, generator=g)).cuda()
  outs = [ops.conv_gemm(a, w, N, gn=(gamma, beta, 32, True), reverse=r)[1] for r in (0, 1, 0)]
  torch.cuda.synchronize()
  assert torch.equal(outs[0], outs[2]) and torch.equal(outs[0], outs[1])


DUAL_CASES = [  # B, H, W, C (= Cin = N), residual, force_block_n, force_m_sub, force_pairs
    (2, 32, 32, 128, True, 0, 0, 0),        # 4-CTA cluster, halo tiles
    (3, 16, 16, 256, True, 0, 0, 0),        # 2-CTA cluster
    (4, 16, 16, 256, True, 256, 1, 2),      # cta_group::2 pair = one image
    (5, 8, 8, 256, True, 0, 0, 0),          # two images per tile, ragged
    (3, 4, 4, 256, False, 0, 0, 0),         # 16-row images, no residual (shortcut-conv blocks, pyramid)
    (9, 4, 4, 128, True, 128, 2, 0),
    (38, 32, 32, 128, True, 0, 0, 0),
]


@pytest.mark.parametrize("impl", [1, 0], ids=["ref", "umma"])
@pytest.mark.parametrize("case", DUAL_CASES, ids=[f"b{c[0]}_{c[1]}x{c[2]}_{c[3]}{'_res' if c[4] else ''}_bn{c[5]}x{c[6]}_cg{c[7]}" for c in DUAL_CASES])
def test_conv_with_dual_groupnorm_epilogue(impl, case):
  """epi = 2 with out32: the conv2 of a ResBlock keeps its trunk result (x + h)/sqrt(2) in fp32 (+ column statistics for
  other consumers) and ALSO emits act(GroupNorm_0(.)) of the next block (layerspp.py:196) as fp16."""
  B, H, W, Cc, with_res, bn, ms, pairs = case
  if impl == 1 and (bn or B > 9):
    pytest.skip("tile variants do not exist on the CUDA-core path")
  g = torch.Generator().manual_seed(B * 17 + H)
  a = torch.randn(B, H, W, Cc, generator=g).to(torch.float16)
  k = (torch.randn(3, 3, Cc, Cc, generator=g) / np.sqrt(9 * Cc)).numpy()
  bias = torch.randn(Cc, generator=g)
  res = 2.0 * torch.randn(B, H, W, Cc, generator=g) if with_res else None
  gamma, beta = 1 + 0.2 * torch.randn(Cc, generator=g), 0.3 * torch.randn(Cc, generator=g)
  groups = min(Cc // 4, 32)
  sc = float(1 / np.sqrt(2.0))
  v = _conv_ref(a, k, 9) + bias.double()
  if with_res:
    v = v + res.double()
  v = v * sc
  want = on.swish(F.group_norm(v.permute(0, 3, 1, 2), groups, gamma.double(), beta.double(), eps=1e-6)).permute(0, 2, 3, 1)
  o32, o16 = ops.conv_gemm(a.cuda(), ops.pack_conv_weight(k), Cc, bias=bias.cuda(), residual=None if res is None else res.cuda(),
                           scale=sc, impl=impl, force_block_n=bn, force_m_sub=ms, force_cta_pairs=pairs,
                           gn=(gamma.cuda(), beta.cuda(), groups, True, 1e-6, True))
  torch.cuda.synchronize()
  e32, e16 = rel_l2(o32.cpu().numpy(), v.numpy()), rel_l2(o16.float().cpu().numpy(), want.numpy())
  print(f"dual gnf {case} impl={impl}: out32 {e32:.2e} out16 {e16:.2e}")
  assert e32 < 2e-5 and e16 < 6e-4
  per = [rel_l2(o16[b].float().cpu().numpy(), want[b].numpy()) for b in range(B)]
  assert max(per) < 8e-4, per


@pytest.mark.parametrize("dual", [False, True])
@pytest.mark.parametrize("shape", [(3, 16, 16), (5, 8, 8), (3, 4, 4), (2, 32, 32)])
def test_groupnorm_epilogue_16_channels_per_group(shape, dual):
  """The h half of a concatenated [h, skip] GroupNorm input (512 channels, 32 groups: 16 channels per group) normalised
  by the producer: N = 256 with 16 groups (N = 128 with 8 groups at 32x32)."""
  B, H, W = shape
  N = 128 if H == 32 else 256
  groups = N // 16
  g = torch.Generator().manual_seed(B + H)
  a = torch.randn(B, H, W, 128, generator=g).to(torch.float16)
  k = (torch.randn(3, 3, 128, N, generator=g) / np.sqrt(9 * 128)).numpy()
  bias = torch.randn(N, generator=g)
  gamma, beta = 1 + 0.2 * torch.randn(N, generator=g), 0.3 * torch.randn(N, generator=g)
  v = _conv_ref(a, k, 9) + bias.double()
  want = on.swish(F.group_norm(v.permute(0, 3, 1, 2), groups, gamma.double(), beta.double(), eps=1e-6)).permute(0, 2, 3, 1)
  o32, o16 = ops.conv_gemm(a.cuda(), ops.pack_conv_weight(k), N, bias=bias.cuda(), gn=(gamma.cuda(), beta.cuda(), groups, True, 1e-6, dual))
  torch.cuda.synchronize()
  assert rel_l2(o16.float().cpu().numpy(), want.numpy()) < 6e-4
  if dual:
    assert rel_l2(o32.cpu().numpy(), v.numpy()) < 2e-5


@pytest.mark.parametrize("B", [2, 37, 75, 150])
@pytest.mark.parametrize("dual", [False, True])
def test_groupnorm_epilogue_two_pairs_per_image(B, dual):
  """32x32 images on cta_group::2 pairs: two pairs per image, statistics exchanged through tagged global words
  (all SMs busy; clusters of four only fit 132).  Several rounds, repeated launches (tags restart), odd / even counts."""
  H = W = 32
  C_ = 128
  g = torch.Generator().manual_seed(B)
  a = torch.randn(B, H, W, C_, generator=g).to(torch.float16)
  k = (torch.randn(3, 3, C_, C_, generator=g) / np.sqrt(9 * C_)).numpy()
  bias = torch.randn(C_, generator=g)
  res = torch.randn(B, H, W, C_, generator=g) if dual else None
  gamma, beta = 1 + 0.2 * torch.randn(C_, generator=g), 0.3 * torch.randn(C_, generator=g)
  v = _conv_ref(a, k, 9) + bias.double()
  if dual:
    v = v + res.double()
  want = on.swish(F.group_norm(v.permute(0, 3, 1, 2), 32, gamma.double(), beta.double(), eps=1e-6)).permute(0, 2, 3, 1)
  ad, wd, bd, gd, bed = a.cuda(), ops.pack_conv_weight(k), bias.cuda(), gamma.cuda(), beta.cuda()
  rd = None if res is None else res.cuda()
  outs = []
  for rep in range(3):
    o32, o16 = ops.conv_gemm(ad, wd, C_, bias=bd, residual=rd, force_block_n=128, force_m_sub=2, force_cta_pairs=2,
                             gn=(gd, bed, 32, True, 1e-6, dual))
    outs.append(o16)
  torch.cuda.synchronize()
  per = [rel_l2(outs[0][b].float().cpu().numpy(), want[b].numpy()) for b in range(B)]
  assert max(per) < 8e-4, (max(per), int(np.argmax(per)))
  assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("impl", [1, 0], ids=["ref", "umma"])
@pytest.mark.parametrize("case", [(2, 32, 32, 128, 128), (3, 16, 16, 256, 256), (3, 4, 4, 256, 256), (2, 16, 16, 64, 128)],
                         ids=lambda c: f"b{c[0]}_{c[1]}x{c[2]}_{c[3]}to{c[4]}")
def test_conv_split_weights_removes_weight_rounding(impl, case):
  """wsplit = 2 (precise mode): W = fp16(W) + fp16(W - fp16(W)), two K passes.  Against a conv with the SAME fp16
  activations but the full fp32 weights the error drops from ~2e-4 (weight rounding) to accumulation-order level."""
  B, H, W, Cin, Cout = case
  g = torch.Generator().manual_seed(Cin + H)
  a = torch.randn(B, H, W, Cin, generator=g).to(torch.float16)
  k = (torch.randn(3, 3, Cin, Cout, generator=g) / np.sqrt(9 * Cin)).numpy()
  x = a.double().permute(0, 3, 1, 2)
  want = F.conv2d(x, torch.as_tensor(k).double().permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)     # fp32 weights, not rounded
  o_split, _ = ops.conv_gemm(a.cuda(), ops.pack_conv_weight(k, split=True), Cout, impl=impl, wsplit=2, w_ld=2 * 9 * Cin)
  o_plain, _ = ops.conv_gemm(a.cuda(), ops.pack_conv_weight(k), Cout, impl=impl)
  torch.cuda.synchronize()
  e_split, e_plain = rel_l2(o_split.cpu().numpy(), want.numpy()), rel_l2(o_plain.cpu().numpy(), want.numpy())
  print(f"split weights {case} impl={impl}: plain {e_plain:.2e} -> split {e_split:.2e}")
  assert e_split < 1e-5 and e_plain > 5e-5          # (tensor-core accumulation order ~4e-6)
