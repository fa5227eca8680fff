"""Host-side launch plan of the score network (csrc/unet.cpp walk), inspected through the C ABI without a GPU:
which kernels an NCSNpp.__call__ (cld_jax/models/ncsnpp.py:41-243) turns into for the BASELINE configurations."""
import collections

import pytest

from gddim_b200 import configs, net

KIND = {0: "stem", 1: "groupnorm", 2: "gemm", 3: "head", 4: "im2col", 5: "transpose_v", 6: "small_attn",
        7: "softmax_rows", 8: "attn_fused", 9: "gn_qkv"}


def _plan(cfg, cld=True, batch=4):
  return net.ScoreNet(cfg, cld=cld).plan(batch)


def test_deep_cifar_plan_uses_the_fused_attention_block():
  plan = _plan(configs.cld_accr_dcifar10(), batch=256)
  kinds = collections.Counter(KIND[k] for _, k in plan)
  # attention blocks in creation order (ncsnpp.py:150-226): eight in the encoder at 16x16 (one per res-block), the
  # middle block at 4x4, one in the decoder at 16x16.  16x16: GroupNorm coefficients -> gn_qkv -> attention + projection,
  # nothing of the unfused chain; 4x4 (16 tokens): the CUDA-core attention kernel between plain GEMMs
  for i in range(10):
    tags = [t.split("/", 1)[1] for t, _ in plan if t.startswith(f"AttnBlockpp_{i}/")]
    # blocks 0..6 are followed by a ResBlock at the same resolution: the fused kernel also emits that block's
    # act(GroupNorm_0(.)) ("+gn0"); block 7 feeds the FIR-downsampling block, block 9 the upsampling block
    fused = "attn_proj_fused+gn0" if i < 7 else "attn_proj_fused"
    assert tags == (["gn", "qkv", "attn_small", "proj"] if i == 8 else ["gn_coef", "gn_qkv", fused]), (i, tags)
  assert kinds["gn_qkv"] == 9 and kinds["attn_fused"] == 9 and kinds["transpose_v"] == 0 and kinds["softmax_rows"] == 0
  assert kinds["small_attn"] == 1
  # every convolution / NIN is a tcgen05 GEMM op; stem and head are GEMM ops too (no CUDA-core conv kernels in the plan)
  assert kinds["stem"] == 0 and kinds["head"] == 0 and kinds["gemm"] > 150
  # GroupNorm + swish between conv1 and conv2 of a ResBlock is applied by conv1's epilogue wherever the measured
  # per-layer rule (unet.cpp gnf_pays) allows: everywhere but the nine single-wave 8x8 K = 4608 layers (+ the 32x32
  # N = 256 up-sampling block, whose image spans eight tiles); GroupNorm_0 of 32 blocks comes from the producer of
  # their input (conv2 / pyramid conv / stem / fused attention)
  tags = collections.Counter(t.split("/", 1)[-1] for t, _ in plan)
  assert tags["conv1_gn1"] == 66 and tags["gn1"] == 10
  assert tags["conv2+gn0"] + tags["conv+gn0"] + tags["attn_proj_fused+gn0"] + tags["stem+gn0"] == 32
  # the plan is a property of the architecture, not of the batch
  assert [t for t, _ in _plan(configs.cld_accr_dcifar10(), batch=8)] == [t for t, _ in plan]


def test_blur_and_ddpmpp_plans():
  blur = _plan(configs.blur_ddpm_deep_cifar10(1.0), cld=False)
  assert sum(1 for _, k in blur if KIND[k] == "attn_fused") == 9
  ddpmpp = _plan(configs.cld_ddpmpp_cifar10())
  kinds = collections.Counter(KIND[k] for _, k in ddpmpp)
  assert kinds["attn_fused"] == kinds["gn_qkv"] and kinds["attn_fused"] >= 1


def test_256_geometry_keeps_the_long_sequence_attention_path():
  cfg = configs.cld_accr_dcifar10()
  cfg.data.image_size = 256
  plan = _plan(cfg, batch=2)
  tags = [t.split("/", 1)[1] for t, _ in plan if t.startswith("AttnBlockpp_0/")]
  # 32x32 = 1024 tokens: scores through HBM in fp32, row softmax kernel, V transpose, P.V, projection
  assert tags == ["gn_coef", "gn_qkv", "qk", "softmax", "vT", "pv", "proj"], tags


def test_unsupported_width_is_rejected_at_context_creation():
  cfg = configs.cld_accr_dcifar10()
  cfg.model.nf = 48                                   # nf must be a multiple of 32
  with pytest.raises(RuntimeError):
    net.ScoreNet(cfg, cld=True).plan(1)


def test_simple_cifar10_nf32_plans():
  """simple_cifar10 (cld_jax/configs/simple_cifar10_config.py:45: nf = 32, ch_mult (1, 2, 2, 2), four res-blocks, naive
  resampling, positional embedding): 32 / 64 / 96 / 128-channel layers.  Layers whose input has 32 or 96 channels are
  planned pixel-paired (unet.cpp pack_conv_paired: [B,H,W,c] read as [B,H,W/2,2c], so that a K block is 64 channels wide)
  and keep their GroupNorms as separate passes; the 64- / 128-channel ones take the usual path, GroupNorm epilogue included."""
  model = net.ScoreNet(configs.cld_simple_cifar10(), cld=True)
  plan = model.plan(8)
  tags = collections.Counter(t.split("/", 1)[-1] for t, _ in plan)
  kinds = collections.Counter(KIND[k] for _, k in plan)
  assert kinds["gemm"] > 60 and kinds["gn_qkv"] == 0 and kinds["attn_fused"] == 0      # C = 64 attention: unfused chain
  assert tags["qk_softmax"] == tags["pv"] == tags["vT"] and tags["qk_softmax"] >= 1
  assert tags["conv1"] > 0 and tags["conv1_gn1"] > 0                                   # both kinds of conv1 are present
  n_params = sum(int(__import__("numpy").prod(shape)) for shape, _, _ in model.specs().values())
  assert n_params == 3_883_686


def test_blur_simple_cifar10_nf32_plans_like_the_cld_one():
  """blur_jax/configs/simple_cifar10_config.py: same model section, three data channels instead of six."""
  cld = [t for t, _ in net.ScoreNet(configs.cld_simple_cifar10(), cld=True).plan(4)]
  blur = [t for t, _ in net.ScoreNet(configs.blur_simple_cifar10(1.0), cld=False).plan(4)]
  assert cld == blur and blur[0] == "stem_im2col" and blur[-1] == "head"

