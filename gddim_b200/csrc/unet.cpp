// NCSN++ / DDPM++ forward as a static launch plan over the kernels of kernels.h.
//
// Control flow follows cld_jax/models/ncsnpp.py:41-243; blocks follow layerspp.py:61-83 (AttnBlockpp),
// :115-143 (Downsample), :180-227 (ResnetBlockBigGANpp); parameter names follow flax.linen compact-module
// auto naming (<Class>_<k> per parent scope) so that a Flax checkpoint tree can be loaded by name.
//
// Fusions relative to the reference graph:
//   * Dense_0(act(temb)) of every ResBlock depends only on t -> one [temb_dim x sum(C_out)] GEMV per time
//     value, added as a second bias in the conv1 epilogue (conv1's own bias is folded into it).
//   * the 1x1 shortcut conv (Conv_2) is appended to conv2's K loop as a second A segment; (x + h)/sqrt(2)
//     is the epilogue scale.  Identity shortcuts are a residual read in the epilogue.
//   * channel concat [h, skip] is never materialised: GroupNorm reads two sources.
//   * FIR / naive resampling is fused into the GroupNorm+swish apply pass (both branches).
//   * q, k, v projections are one GEMM (N = 3C); softmax is the epilogue of the QK^T GEMM.
#include "unet.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace gddim {

struct UNet::Scope {
  std::string prefix;
  std::map<std::string, int> counts;
  Scope child(const std::string& cls) {
    int k = counts[cls]++;
    Scope s;
    s.prefix = prefix + cls + "_" + std::to_string(k) + "/";
    return s;
  }
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

UNet::UNet(const gddim_model_cfg& cfg, int max_batch, bool precise) : cfg_(cfg), max_batch_(max_batch), precise_(precise) {
  dry_ = true;
  arena_ = reinterpret_cast<char*>(uintptr_t(1) << 20);   // fake non-null base for the dry planning pass
  wts_ = reinterpret_cast<char*>(uintptr_t(1) << 20);
  if (walk() != 0) return;
  arena_peak_ = arena_top_;
}

UNet::~UNet() {
  if (finalized_ || !dry_) {
    if (arena_) cudaFree(arena_);
    if (wts_) cudaFree(wts_);
  }
}

// ---- arenas -------------------------------------------------------------------------------------------
void UNet::flush_pending() {
  if (pending_free_.empty()) return;
  last_flush_at_ = ops_.size();
  for (auto& pf : pending_free_) {
    size_t bytes = pf.second;
    size_t off = (char*)pf.first - arena_;
    size_t i = 0;
    while (i < free_.size() && free_[i].first < off) ++i;
    free_.insert(free_.begin() + i, {off, bytes});
    if (i + 1 < free_.size() && free_[i].first + free_[i].second == free_[i + 1].first) {
      free_[i].second += free_[i + 1].second;
      free_.erase(free_.begin() + i + 1);
    }
    if (i > 0 && free_[i - 1].first + free_[i - 1].second == free_[i].first) {
      free_[i - 1].second += free_[i].second;
      free_.erase(free_.begin() + i);
    }
  }
  pending_free_.clear();
}

void* UNet::a_alloc(size_t bytes) {
  bytes = align_up(bytes, 1024);
  if (!hold_pending_) flush_pending();
  for (size_t i = 0; i < free_.size(); ++i) {
    if (free_[i].second >= bytes) {
      const size_t off = free_[i].first;
      if (free_[i].second == bytes) free_.erase(free_.begin() + i);
      else { free_[i].first += bytes; free_[i].second -= bytes; }
      return arena_ + off;
    }
  }
  const size_t off = arena_top_;
  arena_top_ += bytes;
  return arena_ + off;
}

void UNet::a_free(void* p, size_t bytes) {
  pending_free_.push_back({p, align_up(bytes, 1024)});      // see flush_pending()
}

void* UNet::w_alloc(size_t bytes) {
  bytes = align_up(bytes, 256);
  const size_t off = weight_top_;
  weight_top_ += bytes;
  return wts_ + off;
}

float* UNet::upload_f32(const std::vector<float>& v) {
  float* d = (float*)w_alloc(v.size() * sizeof(float));
  if (!dry_) cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice);
  return d;
}
__half* UNet::upload_f16(const std::vector<__half>& v) {
  __half* d = (__half*)w_alloc(v.size() * sizeof(__half));
  if (!dry_) cudaMemcpy(d, v.data(), v.size() * sizeof(__half), cudaMemcpyHostToDevice);
  return d;
}

UNet::T32 UNet::new32(int C, int H, int W) {
  T32 t;
  t.C = C; t.H = H; t.W = W;
  t.bytes = (size_t)max_batch_ * H * W * C * sizeof(float);
  t.p = (float*)a_alloc(t.bytes);
  // column statistics written by the producing GEMM's epilogue (one (sum, sumsq) pair per 32-row slab and channel)
  t.stats = nullptr;
  t.stats_bytes = 0;
  t.stats_valid = false;
  if ((H * W) % 32 == 0) {
    t.stats_bytes = (size_t)max_batch_ * H * W / 32 * 2 * C * sizeof(float);
    t.stats = (float*)a_alloc(t.stats_bytes);
  }
  return t;
}
UNet::T16 UNet::new16(int C, int H, int W) {
  T16 t;
  t.C = C; t.H = H; t.W = W;
  t.bytes = (size_t)max_batch_ * H * W * C * sizeof(__half);
  t.p = (__half*)a_alloc(t.bytes);
  return t;
}
void UNet::rel(T32& t) {
  if (t.p) a_free(t.p, t.bytes);
  if (t.stats) a_free(t.stats, t.stats_bytes);
  t.p = nullptr;
  t.stats = nullptr;
}
void UNet::rel(T16& t) { if (t.p) a_free(t.p, t.bytes); t.p = nullptr; }

// ---- parameters -----------------------------------------------------------------------------------------
const std::vector<float>* UNet::param(Scope& s, const std::string& name, std::vector<int> shape, int kind, float scale) {
  const std::string full = s.prefix + name;
  if (dry_) {
    specs_.push_back({full, shape, kind, scale});
    return nullptr;
  }
  auto it = host_params_.find(full);
  if (it == host_params_.end()) { err_ = "parameter not loaded: " + full; return nullptr; }
  size_t n = 1;
  for (int d : shape) n *= d;
  if (it->second.size() != n) { err_ = "parameter size mismatch: " + full; return nullptr; }
  return &it->second;
}

int UNet::set_param(const std::string& name, const float* host, size_t n) {
  for (const auto& sp : specs_)
    if (sp.name == name) {
      size_t need = 1;
      for (int d : sp.shape) need *= d;
      if (need != n) return fail("gddim_param_set: size mismatch for " + name);
      host_params_[name].assign(host, host + n);
      return 0;
    }
  return fail("gddim_param_set: unknown parameter " + name);
}

static float scale0(float s) { return s == 0.f ? 1e-10f : s; }   // layers.py:62

// Raw (un-normalised) activations that feed a GEMM in fp16 -- the shortcut branch and the input pyramid -- are
// stored divided by 64 and their weights multiplied by 64 (exact power-of-two rescaling): headroom up to
// |x| ~ 4e6 before fp16 overflows, which matters when a random-weight network lets the reverse-time state grow.
constexpr float kRawScale = 1.0f / 64.0f;

// conv kernel HWIO (kh,kw,cin,cout) -> K-major [cout][koff + tap*cin + ci] inside rows of length ld
// lo_off > 0 (precise mode): the rounding remainder w - fp16(w) goes, again as fp16, lo_off columns further -- the GEMM
// then runs its K loop twice (A x W_hi + A x W_lo): the weights enter with ~22 significant bits
static void pack_conv(const std::vector<float>& k, int taps, int cin, int cout, std::vector<__half>& dst, int ld,
                      int koff, float mul = 1.f, int lo_off = 0) {
  for (int tap = 0; tap < taps; ++tap)
    for (int ci = 0; ci < cin; ++ci) {
      const float* src = k.data() + ((size_t)tap * cin + ci) * cout;
      for (int co = 0; co < cout; ++co) {
        const float w = src[co] * mul;
        const __half hi = __float2half_rn(w);
        dst[(size_t)co * ld + koff + tap * cin + ci] = hi;
        if (lo_off > 0) dst[(size_t)co * ld + lo_off + koff + tap * cin + ci] = __float2half_rn(w - __half2float(hi));
      }
    }
}

// Pixel pairing for channel counts that are odd multiples of 32 (nf = 32 networks: 32 / 96 channels): a tcgen05 K block is
// 64 fp16 = 128 bytes, and in NHWC two horizontally adjacent pixels of a 32-channel tensor ARE 64 contiguous channels.  A
// [B, H, W, c] tensor is therefore read as [B, H, W/2, 2c] and the convolution becomes a 3x3 (or 1x1) convolution on the
// half-width grid of "super-pixels" with 2c input and 2n output channels (the [B, H, W, n] result in the same memory):
//   W'[dy, dX, (p_in, ci), (p_out, co)] = W[dy, dx, ci, co],  dx = 2 dX + p_in - p_out,  zero where |dx| > 1
// SAME padding of the super-pixel grid is SAME padding of the pixel grid (whole super-pixels of zeros).  A third of the MMA
// work multiplies structural zeros -- on layers that are a percent of a network's FLOPs, and in exchange every layer of the
// network runs on the tensor cores through the one kernel.  Layout as pack_conv: K-major rows [p_out*cout + co][koff + ...].
static void pack_conv_paired(const std::vector<float>& k, int taps, int cin, int cout, std::vector<__half>& dst, int ld,
                             int koff, float mul = 1.f, int lo_off = 0) {
  const int n3 = taps == 9 ? 3 : 1;
  for (int dyi = 0; dyi < n3; ++dyi)
    for (int dXi = 0; dXi < n3; ++dXi)
      for (int pi = 0; pi < 2; ++pi)
        for (int po = 0; po < 2; ++po) {
          const int dx = taps == 9 ? 2 * (dXi - 1) + pi - po : (pi == po ? 0 : 2);
          if (dx < -1 || dx > 1) continue;                                    // structural zero (dst is zero-initialised)
          const int tap_src = taps == 9 ? dyi * 3 + dx + 1 : 0, tap_dst = taps == 9 ? dyi * 3 + dXi : 0;
          for (int ci = 0; ci < cin; ++ci) {
            const float* src = k.data() + ((size_t)tap_src * cin + ci) * cout;
            for (int co = 0; co < cout; ++co) {
              const float w = src[co] * mul;
              const __half hi = __float2half_rn(w);
              const size_t at = (size_t)(po * cout + co) * ld + koff + (size_t)tap_dst * 2 * cin + pi * cin + ci;
              dst[at] = hi;
              if (lo_off > 0) dst[at + lo_off] = __float2half_rn(w - __half2float(hi));
            }
          }
        }
}
static bool needs_pairing(int c) { return c % 64 != 0; }       // (c % 32 == 0 is guaranteed by the nf check in walk())

std::pair<const float*, const float*> UNet::gn_params(Scope& s, int C) {
  Scope g = s.child("GroupNorm");
  const auto* sc = param(g, "scale", {C}, 2, 1.f);
  const auto* bi = param(g, "bias", {C}, 1, 1.f);
  if (dry_) { w_alloc(C * 4); w_alloc(C * 4); return {nullptr, nullptr}; }
  if (!sc || !bi) return {nullptr, nullptr};
  return {upload_f32(*sc), upload_f32(*bi)};
}

void UNet::add_norm(const T32& in1, const T32* in2, const float* gamma, const float* beta, bool silu, int resample,
                    T16* dst, T16* raw, const std::string& tag) {
  Op op;
  op.kind = OP_NORM;
  op.tag = tag;
  NormOp& n = op.norm;
  memset(&n, 0, sizeof(n));
  n.src1 = in1.p; n.c1 = in1.C;
  n.src2 = in2 ? in2->p : nullptr; n.c2 = in2 ? in2->C : 0;
  n.B = max_batch_; n.H = in1.H; n.W = in1.W;
  const int C = n.c1 + n.c2;
  n.groups = std::min(C / 4, 32);
  n.gamma = gamma; n.beta = beta;
  n.eps = 1e-6f;
  n.silu = silu ? 1 : 0;
  n.resample = resample;
  n.colstats1 = in1.stats_valid ? in1.stats : nullptr;
  n.colstats2 = (in2 && in2->stats_valid) ? in2->stats : nullptr;
  n.partial = gn_partial_;
  n.coef = gn_coef_;
  n.ticket = gn_ticket_;
  n.splits = norm_splits(max_batch_, in1.H, in1.W);
  n.dst16 = dst ? dst->p : nullptr;
  n.raw16 = raw ? raw->p : nullptr;
  n.raw_scale = kRawScale;
  ops_.push_back(op);
}

int UNet::add_temb_proj(const std::vector<float>* w, const std::vector<float>* b, const std::vector<float>* conv_b,
                        int out_ch, int copies) {
  const int off = temb_total_;
  temb_total_ += out_ch * copies;
  if (!dry_) {
    // proj_w_host_ is [temb_dim][total] assembled column-block by column-block; total known from the dry pass
    const int total = (int)proj_b_host_.size();
    for (int c = 0; c < copies; ++c) {
      for (int k = 0; k < temb_dim_; ++k)
        for (int n = 0; n < out_ch; ++n) proj_w_host_[(size_t)k * total + off + c * out_ch + n] = (*w)[(size_t)k * out_ch + n];
      for (int n = 0; n < out_ch; ++n) proj_b_host_[off + c * out_ch + n] = (*b)[n] + (*conv_b)[n];
    }
  }
  return off;
}

static GemmOp make_gemm(int B, int H, int W) {
  GemmOp g;
  memset(&g, 0, sizeof(g));
  g.B = B; g.H = H; g.W = W;
  g.scale = 1.f;
  g.epi = EPI_LINEAR;
  return g;
}

// Where the GroupNorm-fused epilogue pays (measured per layer on B200 at batch 256, profiles/r02_gnf_per_op.txt): the
// epilogue gets two passes and a swish.  That is hidden behind the next tile's main loop where a CTA runs several tiles
// (+3 us on a 64 us convolution at 16x16, +17 .. +19 us at 32x32) and always cheaper than the 45 - 53 us GroupNorm pass
// it removes; the single-wave 8x8 layers (one tile per CTA: the epilogue is fully exposed) gain 4 us at K = 2304 and lose
// 2 us at K = 4608, which therefore keep the separate 12 us pass.  GDDIM_GNF_ALL=1 fuses wherever the geometry allows.
static bool gnf_pays(int H, int W, int K, bool dual) {
  static const bool all = [] { const char* e = getenv("GDDIM_GNF_ALL"); return e && e[0] == '1'; }();
  if (all) return true;
  return !(H * W == 64 && !dual && K >= 4608);
}

// ---- ResnetBlockBigGANpp (layerspp.py:180-227) ------------------------------------------------------------
UNet::T32 UNet::resblock(Scope& top, const T32& in1, const T32* in2, int out_ch, bool up, bool down) {
  Scope s = top.child("ResnetBlockBigGANpp");
  const int Cin = in1.C + (in2 ? in2->C : 0);
  if (out_ch == 0) out_ch = Cin;
  const int H = in1.H, W = in1.W;
  const int Ho = up ? H * 2 : (down ? H / 2 : H), Wo = up ? W * 2 : (down ? W / 2 : W);
  const bool need_sc = (Cin != out_ch) || up || down;
  int rs = RS_NONE;
  if (up) rs = cfg_.fir ? RS_FIR_UP : RS_NAIVE_UP;
  if (down) rs = cfg_.fir ? RS_FIR_DOWN : RS_NAIVE_DOWN;
  const float out_scale = cfg_.skip_rescale ? (float)(1.0 / std::sqrt(2.0)) : 1.f;

  auto gn0 = gn_params(s, Cin);
  // GroupNorm_0 + swish of THIS block applied by the epilogue of the GEMM that produced the input (dual GroupNorm
  // epilogue, conv_gemm.cu): possible when that GEMM is the op pushed last (nothing ran in between), the input is a single
  // un-resampled tensor consumed without a raw shortcut copy, and the geometry qualifies.  The producer then writes its
  // fp32 result (trunk / residual / skip) AND this block's fp16 A operand; the separate pass disappears.
  static const bool no_gnf0 = [] { const char* e = getenv("GDDIM_NO_GNF0"); const char* e1 = getenv("GDDIM_NO_GNF");
                                   return (e && e[0] == '1') || (e1 && e1[0] == '1'); }();
  const int groups0 = std::min(Cin / 4, 32);
  bool fuse0 = false, fuse_attn = false;
  if (!no_gnf0 && in2 == nullptr && rs == RS_NONE && !need_sc && in1.prod_op >= 0 && in1.prod_op == (int)ops_.size() - 1 &&
      last_flush_at_ < ops_.size() && gemm_gnf_supported(H, W, Cin, groups0)) {
    const Op& po = ops_[in1.prod_op];
    const GemmOp& pg = po.gemm;
    if (po.kind == OP_ATTN_FUSED) {
      // producer = the fused attention + projection kernel (one image = a cluster of two CTAs)
      const AttnOp& pa = po.attn;
      fuse_attn = pa.w3 != nullptr && pa.out32 == in1.p && pa.gn_gamma == nullptr && pa.C == Cin && pa.T == H * W &&
                  (Cin / groups0 == 4 || Cin / groups0 == 8 || Cin / groups0 == 16);
    }
    int pk = 0;
    bool ptc = true;                         // the producer runs on the tcgen05 kernel (64-wide K blocks)
    if (po.kind == OP_GEMM)
      for (int sgi = 0; sgi < pg.nseg && sgi < 2; ++sgi) { pk += pg.seg[sgi].taps * pg.seg[sgi].c; ptc = ptc && pg.seg[sgi].c % 64 == 0; }
    fuse0 = po.kind == OP_GEMM && ptc && pg.epi == EPI_LINEAR && pg.out32 == in1.p && pg.out16 == nullptr && pg.rowscale == nullptr &&
            pg.n_store == 0 && pg.w_batch_stride == 0 && pg.N == Cin && pg.ldo == Cin && pg.H == H && pg.W == W &&
            !po.out_is_external && gnf_pays(H, W, pk, true);
  }
  hold_pending_ = fuse0 || fuse_attn;    // a1 becomes an output of the producer: it must not reuse what that op still reads
  T16 a1 = new16(Cin, Ho, Wo);
  hold_pending_ = false;
  T16 x16; x16.p = nullptr;
  if (need_sc) x16 = new16(Cin, Ho, Wo);
  if (fuse0) {
    Op& po = ops_[in1.prod_op];
    GemmOp& pg = po.gemm;
    pg.epi = EPI_GNF;
    pg.out16 = a1.p;
    pg.gn_gamma = gn0.first; pg.gn_beta = gn0.second; pg.gn_eps = 1e-6f; pg.gn_groups = groups0; pg.gn_silu = 1;
    po.tag += "+gn0";
  } else if (fuse_attn) {
    Op& po = ops_[in1.prod_op];
    po.attn.gn_gamma = gn0.first; po.attn.gn_beta = gn0.second; po.attn.gn_eps = 1e-6f; po.attn.gn_groups = groups0;
    po.attn.gn_silu = 1; po.attn.gn_out16 = a1.p;
    po.tag += "+gn0";
  } else {
    add_norm(in1, in2, gn0.first, gn0.second, true, rs, &a1, need_sc ? &x16 : nullptr, s.prefix + "gn0");
  }

  // conv1 (Conv_0) + Dense_0(act(temb)).  pf1 / pf2 = 2: pixel-paired (pack_conv_paired) -- half-width grid, twice the
  // channels on both sides; such layers keep their GroupNorms as separate passes and leave the statistics to them
  const int pf1 = needs_pairing(Cin) ? 2 : 1;
  const int pf2 = (needs_pairing(out_ch) || (need_sc && needs_pairing(Cin))) ? 2 : 1;
  if ((pf1 == 2 || pf2 == 2) && Wo % 2 != 0) { err_ = "pixel pairing needs an even width"; return T32{nullptr, 0, 0, 0, 0}; }
  Scope c0 = s.child("Conv");
  const auto* k0 = param(c0, "kernel", {3, 3, Cin, out_ch}, 0, 1.f);
  const auto* b0 = param(c0, "bias", {out_ch}, 1, 1.f);
  const float* bias1 = nullptr;
  int temb_off = -1;
  if (cfg_.conditional) {
    Scope d0 = s.child("Dense");
    const auto* dw = param(d0, "kernel", {temb_dim_, out_ch}, 0, 1.f);
    const auto* db = param(d0, "bias", {out_ch}, 1, 1.f);
    if (!dry_ && (!dw || !db || !b0)) return T32{nullptr, 0, 0, 0, 0};
    temb_off = add_temb_proj(dw, db, b0, out_ch, pf1);
  } else {
    if (dry_) w_alloc(out_ch * pf1 * 4);
    else {
      if (!b0) return T32{nullptr, 0, 0, 0, 0};
      std::vector<float> bb;
      for (int c = 0; c < pf1; ++c) bb.insert(bb.end(), b0->begin(), b0->end());
      bias1 = upload_f32(bb);
    }
  }
  __half* w1 = nullptr;
  {
    const int wmul = precise_ ? 2 : 1;
    const size_t n = (size_t)out_ch * pf1 * 9 * Cin * pf1 * wmul;
    if (dry_) w1 = (__half*)w_alloc(n * 2);
    else {
      if (!k0) return T32{nullptr, 0, 0, 0, 0};
      std::vector<__half> pk(n, __float2half(0.f));
      if (pf1 == 2) pack_conv_paired(*k0, 9, Cin, out_ch, pk, 9 * 2 * Cin * wmul, 0, 1.f, precise_ ? 9 * 2 * Cin : 0);
      else pack_conv(*k0, 9, Cin, out_ch, pk, 9 * Cin * wmul, 0, 1.f, precise_ ? 9 * Cin : 0);
      w1 = upload_f16(pk);
    }
  }
  // GroupNorm_1 + swish applied by conv1's own epilogue (EPI_GNF, conv_gemm.cu): h = conv1(...) + Dense_0(act(temb)) is
  // never written; conv1 emits conv2's fp16 A operand directly.  Geometries the epilogue does not cover (images spanning
  // more than four CTAs) keep the separate pass.  GDDIM_NO_GNF=1: A/B switch.
  static const bool no_gnf = [] { const char* e = getenv("GDDIM_NO_GNF"); return e && e[0] == '1'; }();
  auto gn1 = gn_params(s, out_ch);
  const int groups1 = std::min(out_ch / 4, 32);
  const bool fuse1 = !no_gnf && Cin % 64 == 0 && gemm_gnf_supported(Ho, Wo, out_ch, groups1) && gnf_pays(Ho, Wo, 9 * Cin, false);
  T16 a2 = new16(out_ch, Ho, Wo);
  T32 h2{nullptr, 0, 0, 0, 0};
  if (!fuse1) h2 = new32(out_ch, Ho, Wo);
  {
    Op op; op.kind = OP_GEMM; op.tag = s.prefix + (fuse1 ? "conv1_gn1" : "conv1");
    op.gemm = make_gemm(max_batch_, Ho, Wo / pf1);
    GemmOp& g = op.gemm;
    g.nseg = 1;
    g.seg[0] = {a1.p, Cin * pf1, 0, Cin * pf1, 9};
    g.w = w1; g.N = out_ch * pf1; g.w_ld = 9 * Cin * pf1 * (precise_ ? 2 : 1); g.wsplit = precise_ ? 2 : 1;
    g.bias = bias1;
    g.bias2 = temb_off >= 0 ? temb_cur_ + temb_off : nullptr;
    g.ldo = out_ch * pf1;
    if (fuse1) {
      g.epi = EPI_GNF;
      g.out16 = a2.p;
      g.gn_gamma = gn1.first; g.gn_beta = gn1.second; g.gn_eps = 1e-6f; g.gn_groups = groups1; g.gn_silu = 1;
    } else {
      g.out32 = h2.p;
      if (pf1 == 1) { g.colstats = h2.stats; h2.stats_valid = h2.stats != nullptr; }   // (paired: columns are (pixel, channel))
    }
    ops_.push_back(op);
  }
  rel(a1);
  if (!fuse1) {
    add_norm(h2, nullptr, gn1.first, gn1.second, true, RS_NONE, &a2, nullptr, s.prefix + "gn1");
    rel(h2);
  }

  Scope c1 = s.child("Conv");
  const auto* k1 = param(c1, "kernel", {3, 3, out_ch, out_ch}, 0, scale0(0.f));   // init_scale = config.model.init_scale
  const auto* b1 = param(c1, "bias", {out_ch}, 1, 1.f);
  const std::vector<float>* k2 = nullptr;
  const std::vector<float>* b2 = nullptr;
  if (need_sc) {
    Scope c2 = s.child("Conv");
    k2 = param(c2, "kernel", {1, 1, Cin, out_ch}, 0, 1.f);
    b2 = param(c2, "bias", {out_ch}, 1, 1.f);
  }
  const int ktot = (9 * out_ch + (need_sc ? Cin : 0)) * pf2;
  __half* w2 = nullptr;
  const float* bias2v = nullptr;
  const int wld2 = ktot * (precise_ ? 2 : 1), lo2 = precise_ ? ktot : 0;
  if (dry_) { w2 = (__half*)w_alloc((size_t)out_ch * pf2 * wld2 * 2); w_alloc(out_ch * pf2 * 4); }
  else {
    if (!k1 || !b1 || (need_sc && (!k2 || !b2))) return T32{nullptr, 0, 0, 0, 0};
    std::vector<__half> pk((size_t)out_ch * pf2 * wld2, __float2half(0.f));
    if (pf2 == 2) pack_conv_paired(*k1, 9, out_ch, out_ch, pk, wld2, 0, 1.f, lo2);
    else pack_conv(*k1, 9, out_ch, out_ch, pk, wld2, 0, 1.f, lo2);
    std::vector<float> bsum(*b1);
    if (need_sc) {
      if (pf2 == 2) pack_conv_paired(*k2, 1, Cin, out_ch, pk, wld2, 9 * 2 * out_ch, 1.0f / kRawScale, lo2);
      else pack_conv(*k2, 1, Cin, out_ch, pk, wld2, 9 * out_ch, 1.0f / kRawScale, lo2);
      for (int i = 0; i < out_ch; ++i) bsum[i] += (*b2)[i];
    }
    if (pf2 == 2) bsum.insert(bsum.end(), bsum.begin(), bsum.begin() + out_ch);
    w2 = upload_f16(pk);
    bias2v = upload_f32(bsum);
  }
  T32 out = new32(out_ch, Ho, Wo);
  {
    Op op; op.kind = OP_GEMM; op.tag = s.prefix + "conv2";
    op.gemm = make_gemm(max_batch_, Ho, Wo / pf2);
    GemmOp& g = op.gemm;
    g.nseg = need_sc ? 2 : 1;
    g.seg[0] = {a2.p, out_ch * pf2, 0, out_ch * pf2, 9};
    if (need_sc) g.seg[1] = {x16.p, Cin * pf2, 0, Cin * pf2, 1};
    g.w = w2; g.N = out_ch * pf2; g.w_ld = wld2; g.wsplit = precise_ ? 2 : 1;
    g.bias = bias2v;
    g.residual = need_sc ? nullptr : in1.p;
    g.scale = out_scale;
    g.out32 = out.p; g.ldo = out_ch * pf2;
    if (pf2 == 1) { g.colstats = out.stats; out.stats_valid = out.stats != nullptr; }
    out.prod_op = (int)ops_.size();
    ops_.push_back(op);
  }
  rel(a2);
  if (need_sc) rel(x16);
  return out;
}

// ---- AttnBlockpp (layerspp.py:61-83) ------------------------------------------------------------------------
UNet::T32 UNet::attnblock(Scope& top, const T32& x) {
  Scope s = top.child("AttnBlockpp");
  const int C = x.C, H = x.H, W = x.W, T = H * W;
  const float out_scale = cfg_.skip_rescale ? (float)(1.0 / std::sqrt(2.0)) : 1.f;
  auto gn = gn_params(s, C);
  // GroupNorm apply fused into the q/k/v projection (gn_qkv.cu): only the coefficient table is computed here
  static const bool no_gnqkv = [] { const char* e = getenv("GDDIM_NO_FUSED_GNQKV"); return e && e[0] == '1'; }();
  const bool gnqkv = gn_qkv_supported(T, C, 3 * C) && x.stats_valid && !no_gnqkv;
  T16 h16{nullptr, 0, 0, 0, 0};
  if (gnqkv) {
    add_norm(x, nullptr, gn.first, gn.second, false, RS_NONE, nullptr, nullptr, s.prefix + "gn_coef");
    ops_.back().norm.coef_only = 1;
  } else {
    h16 = new16(C, H, W);
    add_norm(x, nullptr, gn.first, gn.second, false, RS_NONE, &h16, nullptr, s.prefix + "gn");
  }

  const std::vector<float>* nw[4];
  const std::vector<float>* nb[4];
  for (int i = 0; i < 4; ++i) {
    Scope n = s.child("NIN");
    nw[i] = param(n, "W", {C, C}, 0, i == 3 ? scale0(0.f) : 0.1f);
    nb[i] = param(n, "b", {C}, 1, 1.f);
  }
  __half *wqkv = nullptr, *w3 = nullptr;
  const float *bqkv = nullptr, *b3 = nullptr;
  if (dry_) {
    wqkv = (__half*)w_alloc((size_t)3 * C * C * 2); w_alloc(3 * C * 4);
    w3 = (__half*)w_alloc((size_t)C * C * 2); w_alloc(C * 4);
  } else {
    for (int i = 0; i < 4; ++i) if (!nw[i] || !nb[i]) return T32{nullptr, 0, 0, 0, 0};
    std::vector<__half> pk((size_t)3 * C * C);
    std::vector<float> bb(3 * C);
    for (int i = 0; i < 3; ++i) {
      for (int k = 0; k < C; ++k)
        for (int n = 0; n < C; ++n) pk[((size_t)i * C + n) * C + k] = __float2half_rn((*nw[i])[(size_t)k * C + n]);
      for (int n = 0; n < C; ++n) bb[i * C + n] = (*nb[i])[n];
    }
    wqkv = upload_f16(pk);
    bqkv = upload_f32(bb);
    std::vector<__half> p3((size_t)C * C);
    for (int k = 0; k < C; ++k)
      for (int n = 0; n < C; ++n) p3[(size_t)n * C + k] = __float2half_rn((*nw[3])[(size_t)k * C + n]);
    w3 = upload_f16(p3);
    b3 = upload_f32(*nb[3]);
  }
  T16 qkv = new16(3 * C, H, W);
  if (gnqkv) {
    Op op; op.kind = OP_GN_QKV; op.tag = s.prefix + "gn_qkv";
    op.H = H; op.W = W; op.cout = 3 * C; op.T = T;
    memset(&op.gq, 0, sizeof(op.gq));
    op.gq.x = x.p; op.gq.coef = gn_coef_; op.gq.w = wqkv; op.gq.bias = bqkv; op.gq.out16 = qkv.p;
    op.gq.B = max_batch_; op.gq.T = T; op.gq.C = C; op.gq.N = 3 * C;
    ops_.push_back(op);
  } else {
    Op op; op.kind = OP_GEMM; op.tag = s.prefix + "qkv";
    op.gemm = make_gemm(max_batch_, H, W);
    GemmOp& g = op.gemm;
    g.nseg = 1; g.seg[0] = {h16.p, C, 0, C, 1};
    g.w = wqkv; g.N = 3 * C; g.w_ld = C; g.bias = bqkv;
    g.out16 = qkv.p; g.ldo = 3 * C;
    ops_.push_back(op);
  }
  rel(h16);
  T16 o16 = new16(C, H, W);
  const float sm_scale = 1.0f / std::sqrt((float)C);
  static const bool no_fused_attn = [] { const char* e = getenv("GDDIM_NO_FUSED_ATTN"); return e && e[0] == '1'; }();
  static const bool no_fused_proj = [] { const char* e = getenv("GDDIM_NO_FUSED_PROJ"); return e && e[0] == '1'; }();
  bool proj_fused = false;
  T32 out_fused{nullptr, 0, 0, 0, 0};
  if (attn_fused_supported(T, C) && !no_fused_attn && !no_fused_proj) {
    // ... and the output projection + residual as well: one kernel from qkv to the block output
    out_fused = new32(C, H, W);
    Op op; op.kind = OP_ATTN_FUSED; op.tag = s.prefix + "attn_proj_fused";
    op.H = H; op.W = W; op.cout = C; op.T = T;
    memset(&op.attn, 0, sizeof(op.attn));
    op.attn.qkv = qkv.p; op.attn.out16 = nullptr; op.attn.B = max_batch_; op.attn.T = T; op.attn.C = C;
    op.attn.scale = sm_scale;
    op.attn.w3 = w3; op.attn.bias3 = b3; op.attn.residual = x.p; op.attn.out32 = out_fused.p;
    op.attn.colstats = out_fused.stats; op.attn.out_scale = out_scale;
    if (out_fused.stats != nullptr || dry_) {
      out_fused.stats_valid = out_fused.stats != nullptr;
      out_fused.prod_op = (int)ops_.size();
      ops_.push_back(op);
      proj_fused = true;
    } else {
      rel(out_fused);
    }
  }
  if (proj_fused) {
  } else if (attn_fused_supported(T, C) && !no_fused_attn) {
    // QK^T -> softmax -> P.V in one kernel (attn.cu): scores / probabilities stay in TMEM / shared memory
    Op op; op.kind = OP_ATTN_FUSED; op.tag = s.prefix + "attn_fused";
    op.H = H; op.W = W; op.cout = C; op.T = T;
    memset(&op.attn, 0, sizeof(op.attn));
    op.attn.qkv = qkv.p; op.attn.out16 = o16.p; op.attn.B = max_batch_; op.attn.T = T; op.attn.C = C;
    op.attn.scale = sm_scale;
    ops_.push_back(op);
  } else if (T == 256 && C % 64 == 0) {
    T16 p16 = new16(T, H, W);
    T16 vT; vT.C = T; vT.H = C; vT.W = 1; vT.bytes = (size_t)max_batch_ * C * T * 2; vT.p = (__half*)a_alloc(vT.bytes);
    float* rowinv = (float*)a_alloc((size_t)max_batch_ * T * 4);
    {
      Op op; op.kind = OP_GEMM; op.tag = s.prefix + "qk_softmax";
      op.gemm = make_gemm(max_batch_, H, W);
      GemmOp& g = op.gemm;
      g.nseg = 1; g.seg[0] = {qkv.p, 3 * C, 0, C, 1};
      g.w = qkv.p; g.N = T; g.w_ld = 3 * C; g.w_koff = C;
      g.w_batch_stride = (long long)T * 3 * C; g.w_rows_per_batch = T;
      g.scale = sm_scale; g.epi = EPI_SOFTMAX;
      g.out16 = p16.p; g.row_out = rowinv; g.ldo = T;
      ops_.push_back(op);
    }
    {
      Op op; op.kind = OP_TRANSPOSE_V; op.tag = s.prefix + "vT";
      op.h_in = qkv.p; op.h_out = vT.p; op.T = T; op.cin = C; op.ld = 3 * C; op.voff = 2 * C;
      ops_.push_back(op);
    }
    {
      Op op; op.kind = OP_GEMM; op.tag = s.prefix + "pv";
      op.gemm = make_gemm(max_batch_, H, W);
      GemmOp& g = op.gemm;
      g.nseg = 1; g.seg[0] = {p16.p, T, 0, T, 1};
      g.w = vT.p; g.N = C; g.w_ld = T;
      g.w_batch_stride = (long long)C * T; g.w_rows_per_batch = C;
      g.rowscale = rowinv;
      g.out16 = o16.p; g.ldo = C;
      ops_.push_back(op);
    }
    rel(p16); rel(vT);
    a_free(rowinv, (size_t)max_batch_ * T * 4);
  } else if (T > 256 && T % 256 == 0 && C % 64 == 0) {
    // long sequences (32x32 tokens of the 256x256 configuration): scores in fp32 through HBM, row softmax, P.V
    T16 p16 = new16(T, H, W);
    T16 vT; vT.C = T; vT.H = C; vT.W = 1; vT.bytes = (size_t)max_batch_ * C * T * 2; vT.p = (__half*)a_alloc(vT.bytes);
    float* rowinv = (float*)a_alloc((size_t)max_batch_ * T * 4);
    const size_t s_bytes = (size_t)max_batch_ * T * T * 4;
    float* s32 = (float*)a_alloc(s_bytes);
    {
      Op op; op.kind = OP_GEMM; op.tag = s.prefix + "qk";
      op.gemm = make_gemm(max_batch_, H, W);
      GemmOp& g = op.gemm;
      g.nseg = 1; g.seg[0] = {qkv.p, 3 * C, 0, C, 1};
      g.w = qkv.p; g.N = T; g.w_ld = 3 * C; g.w_koff = C;
      g.w_batch_stride = (long long)T * 3 * C; g.w_rows_per_batch = T;
      g.scale = sm_scale;
      g.out32 = s32; g.ldo = T;
      ops_.push_back(op);
    }
    {
      Op op; op.kind = OP_SOFTMAX_ROWS; op.tag = s.prefix + "softmax";
      op.f_in = s32; op.h_out = p16.p; op.f_out = rowinv; op.T = T;
      ops_.push_back(op);
    }
    a_free(s32, s_bytes);
    {
      Op op; op.kind = OP_TRANSPOSE_V; op.tag = s.prefix + "vT";
      op.h_in = qkv.p; op.h_out = vT.p; op.T = T; op.cin = C; op.ld = 3 * C; op.voff = 2 * C;
      ops_.push_back(op);
    }
    {
      Op op; op.kind = OP_GEMM; op.tag = s.prefix + "pv";
      op.gemm = make_gemm(max_batch_, H, W);
      GemmOp& g = op.gemm;
      g.nseg = 1; g.seg[0] = {p16.p, T, 0, T, 1};
      g.w = vT.p; g.N = C; g.w_ld = T;
      g.w_batch_stride = (long long)C * T; g.w_rows_per_batch = C;
      g.rowscale = rowinv;
      g.out16 = o16.p; g.ldo = C;
      ops_.push_back(op);
    }
    rel(p16); rel(vT);
    a_free(rowinv, (size_t)max_batch_ * T * 4);
  } else if (T <= 64) {
    Op op; op.kind = OP_SMALL_ATTN; op.tag = s.prefix + "attn_small";
    op.h_in = qkv.p; op.h_out = o16.p; op.T = T; op.cin = C; op.scale = sm_scale;
    ops_.push_back(op);
  } else {
    err_ = "attention over " + std::to_string(T) + " tokens is not supported (16..64, 256 or a multiple of 256)";
    return T32{nullptr, 0, 0, 0, 0};
  }
  rel(qkv);
  if (proj_fused) {
    rel(o16);
    return out_fused;
  }
  T32 out = new32(C, H, W);
  {
    Op op; op.kind = OP_GEMM; op.tag = s.prefix + "proj";
    op.gemm = make_gemm(max_batch_, H, W);
    GemmOp& g = op.gemm;
    g.nseg = 1; g.seg[0] = {o16.p, C, 0, C, 1};
    g.w = w3; g.N = C; g.w_ld = C; g.bias = b3;
    g.residual = x.p; g.scale = out_scale;
    g.out32 = out.p; g.ldo = C;
    g.colstats = out.stats; out.stats_valid = out.stats != nullptr;
    ops_.push_back(op);
  }
  rel(o16);
  return out;
}

// ---- NCSNpp.__call__ (ncsnpp.py:41-243) ------------------------------------------------------------------------
int UNet::walk() {
  ops_.clear();
  free_.clear();
  pending_free_.clear();
  hold_pending_ = false;
  last_flush_at_ = 0;
  arena_top_ = 0;
  weight_top_ = 0;
  temb_total_ = 0;
  const gddim_model_cfg& m = cfg_;
  if (m.n_levels < 1 || m.n_levels > 8) return fail("n_levels out of range");
  if (!m.centered) return fail("config.data.centered = False is not supported");
  // nf = 32 (simple_cifar10): layers with 32 / 96 input channels are planned pixel-paired (pack_conv_paired)
  if (m.nf % 32 != 0) return fail("nf must be a multiple of 32 (got " + std::to_string(m.nf) + ")");
  const int nf = m.nf, S = m.image_size, Cnet = net_channels();
  Scope top;
  temb_dim_ = nf * 4;

  // time embedding parameters (ncsnpp.py:68-91)
  if (m.embedding_type == 0) {
    Scope f = top.child("GaussianFourierProjection");
    const auto* w = param(f, "W", {nf}, 3, 16.f);
    emb_in_dim_ = 2 * nf;
    if (!dry_) { if (!w) return -1; fourier_w_ = *w; }
  } else {
    emb_in_dim_ = nf;
  }
  if (m.conditional) {
    Scope d0 = top.child("Dense");
    const auto* w0 = param(d0, "kernel", {emb_in_dim_, temb_dim_}, 0, 1.f);
    const auto* b0 = param(d0, "bias", {temb_dim_}, 1, 1.f);
    Scope d1 = top.child("Dense");
    const auto* w1 = param(d1, "kernel", {temb_dim_, temb_dim_}, 0, 1.f);
    const auto* b1 = param(d1, "bias", {temb_dim_}, 1, 1.f);
    if (dry_) {
      w_alloc((size_t)emb_in_dim_ * temb_dim_ * 4); w_alloc(temb_dim_ * 4);
      w_alloc((size_t)temb_dim_ * temb_dim_ * 4); w_alloc(temb_dim_ * 4);
    } else {
      if (!w0 || !b0 || !w1 || !b1) return -1;
      d_dense0_w_ = upload_f32(*w0); d_dense0_b_ = upload_f32(*b0);
      d_dense1_w_ = upload_f32(*w1); d_dense1_b_ = upload_f32(*b1);
    }
  }
  // small persistent buffers live at the bottom of the weight arena
  d_temb0_ = (float*)w_alloc(emb_in_dim_ * 4);
  d_temb1_ = (float*)w_alloc(temb_dim_ * 4);
  d_temb2_ = (float*)w_alloc(temb_dim_ * 4);
  gn_partial_ = (float*)w_alloc((size_t)max_batch_ * 32 * 32 * 2 * 4);
  gn_coef_ = (float*)w_alloc((size_t)max_batch_ * 2 * 2048 * 4);
  gn_ticket_ = (unsigned int*)w_alloc((size_t)max_batch_ * 4);
  // temb_cur_: capacity known from the dry pass (first pass: generous upper bound is unnecessary -- the
  // pointer is only used as an address; its size is fixed in finalize()).
  temb_cur_ = (float*)w_alloc((size_t)(dry_ ? 1 : proj_b_host_.size()) * 4 + 16);

  // stem (ncsnpp.py:146)
  Scope stem = top.child("Conv");
  const auto* sk = param(stem, "kernel", {3, 3, Cnet, nf}, 0, 1.f);
  const auto* sb = param(stem, "bias", {nf}, 1, 1.f);
  std::vector<T32> hs;
  {
    // 9*Cnet (= 54 or 27) input taps padded to a 64-wide k-block: window gather + one tcgen05 GEMM.  The raw
    // state is the one operand that is not normalised, so it is fed as a hi/lo fp16 pair against hi/lo weights
    // (3 k-blocks: hi*hi + lo*hi + hi*lo): the stem stays fp32-accurate at negligible cost.
    T32 h0 = new32(nf, S, S);
    const int kseg = (int)align_up(9 * Cnet, 64);
    const int kpad = 3 * kseg;
    T16 a16 = new16(kpad, S, S);
    {
      Op op; op.kind = OP_IM2COL; op.tag = "stem_im2col";
      op.in_is_external = true; op.h_out = a16.p; op.H = S; op.W = S; op.cin = Cnet; op.kpad = kpad; op.use_fir = 2;
      ops_.push_back(op);
    }
    __half* wp = nullptr; const float* bp = nullptr;
    if (dry_) { wp = (__half*)w_alloc((size_t)nf * kpad * 2); w_alloc(nf * 4); }
    else {
      if (!sk || !sb) return -1;
      std::vector<__half> pk((size_t)nf * kpad, __float2half(0.f));
      pack_conv(*sk, 9, Cnet, nf, pk, kpad, 0, 1.0f / kRawScale);
      for (int co = 0; co < nf; ++co)
        for (int k = 0; k < 9 * Cnet; ++k) {
          const int tap = k / Cnet, ci = k % Cnet;
          const float w = (*sk)[((size_t)tap * Cnet + ci) * nf + co] / kRawScale;
          const __half hi = pk[(size_t)co * kpad + k];
          pk[(size_t)co * kpad + kseg + k] = hi;
          pk[(size_t)co * kpad + 2 * kseg + k] = __float2half_rn(w - __half2float(hi));
        }
      wp = upload_f16(pk); bp = upload_f32(*sb);
    }
    {
      Op op; op.kind = OP_GEMM; op.tag = "stem";
      op.gemm = make_gemm(max_batch_, S, S);
      GemmOp& g = op.gemm;
      g.nseg = 1; g.seg[0] = {a16.p, kpad, 0, kpad, 1};
      g.w = wp; g.N = nf; g.w_ld = kpad; g.bias = bp;
      g.out32 = h0.p; g.ldo = nf;
      g.colstats = h0.stats; h0.stats_valid = h0.stats != nullptr;
      h0.prod_op = (int)ops_.size();
      ops_.push_back(op);
    }
    rel(a16);
    hs.push_back(h0);
  }
  bool has_pyr = m.progressive_input == 1;
  T32 pyr; pyr.p = nullptr; pyr.C = Cnet; pyr.H = S; pyr.W = S;     // p == nullptr & first: external x
  bool pyr_external = true;

  auto in_attn = [&](int res) { for (int i = 0; i < m.n_attn; ++i) if (m.attn_resolutions[i] == res) return true; return false; };

  for (int lvl = 0; lvl < m.n_levels; ++lvl) {
    for (int blk = 0; blk < m.num_res_blocks; ++blk) {
      T32 h = resblock(top, hs.back(), nullptr, nf * m.ch_mult[lvl], false, false);
      if (!err_.empty()) return -1;
      if (in_attn(h.H)) {
        T32 h2 = attnblock(top, h);
        if (!err_.empty()) return -1;
        rel(h);
        h = h2;
      }
      hs.push_back(h);
    }
    if (lvl != m.n_levels - 1) {
      T32 h = resblock(top, hs.back(), nullptr, 0, false, true);
      if (!err_.empty()) return -1;
      if (has_pyr) {
        if (!m.fir) return fail("progressive_input='residual' without FIR is used by no shipped config");
        // Downsample(fir, with_conv) = conv_downsample_2d (layerspp.py:115-143; up_or_down_sampling.py:168-209)
        Scope ds = top.child("Downsample");
        Scope c2d = ds.child("Conv2d");
        const int cin = pyr.C, cout = h.C;
        const auto* pw = param(c2d, "weight", {3, 3, cin, cout}, 0, 1.f);
        const auto* pb = param(c2d, "bias", {cout}, 1, 1.f);
        const int kpad = (int)align_up(9 * cin, 64);
        T16 a16 = new16(kpad, pyr.H / 2, pyr.W / 2);
        {
          Op op; op.kind = OP_IM2COL; op.tag = ds.prefix + "im2col";
          op.in_is_external = pyr_external; op.f_in = pyr.p; op.h_out = a16.p;
          op.H = pyr.H; op.W = pyr.W; op.cin = cin; op.kpad = kpad; op.use_fir = 1;
          ops_.push_back(op);
        }
        __half* wp = nullptr; const float* bp = nullptr;
        if (dry_) { wp = (__half*)w_alloc((size_t)cout * kpad * 2); w_alloc(cout * 4); }
        else {
          if (!pw || !pb) return -1;
          std::vector<__half> pk((size_t)cout * kpad, __float2half(0.f));
          pack_conv(*pw, 9, cin, cout, pk, kpad, 0, 1.0f / kRawScale);
          wp = upload_f16(pk); bp = upload_f32(*pb);
        }
        T32 np = new32(cout, h.H, h.W);
        {
          Op op; op.kind = OP_GEMM; op.tag = ds.prefix + "conv";
          op.gemm = make_gemm(max_batch_, h.H, h.W);
          GemmOp& g = op.gemm;
          g.nseg = 1; g.seg[0] = {a16.p, kpad, 0, kpad, 1};
          g.w = wp; g.N = cout; g.w_ld = kpad; g.bias = bp;
          g.residual = h.p; g.scale = m.skip_rescale ? (float)(1.0 / std::sqrt(2.0)) : 1.f;
          g.out32 = np.p; g.ldo = cout;
          g.colstats = np.stats; np.stats_valid = np.stats != nullptr;
          np.prod_op = (int)ops_.size();
          ops_.push_back(op);
        }
        rel(a16);
        rel(h);
        h = np;
        pyr = np;            // owned by hs from now on
        pyr_external = false;
      }
      hs.push_back(h);
    }
  }

  // middle (ncsnpp.py:175-178)
  T32 h;
  {
    T32 h1 = resblock(top, hs.back(), nullptr, 0, false, false);
    if (!err_.empty()) return -1;
    T32 h2 = attnblock(top, h1);
    if (!err_.empty()) return -1;
    rel(h1);
    h = resblock(top, h2, nullptr, 0, false, false);
    if (!err_.empty()) return -1;
    rel(h2);
  }

  // up path (ncsnpp.py:183-229)
  for (int lvl = m.n_levels - 1; lvl >= 0; --lvl) {
    for (int blk = 0; blk < m.num_res_blocks + 1; ++blk) {
      T32 skip = hs.back();
      hs.pop_back();
      T32 hn = resblock(top, h, &skip, nf * m.ch_mult[lvl], false, false);
      if (!err_.empty()) return -1;
      rel(h);
      rel(skip);
      h = hn;
    }
    if (in_attn(h.H)) {
      T32 h2 = attnblock(top, h);
      if (!err_.empty()) return -1;
      rel(h);
      h = h2;
    }
    if (lvl != 0) {
      T32 hu = resblock(top, h, nullptr, 0, true, false);
      if (!err_.empty()) return -1;
      rel(h);
      h = hu;
    }
  }
  if (!hs.empty()) return fail("internal: skip stack not empty");

  // head (ncsnpp.py:236-237)
  {
    auto gn = gn_params(top, h.C);
    T16 a = new16(h.C, h.H, h.W);
    add_norm(h, nullptr, gn.first, gn.second, true, RS_NONE, &a, nullptr, "head_gn");
    rel(h);
    Scope hc = top.child("Conv");
    const auto* hk = param(hc, "kernel", {3, 3, a.C, Cnet}, 0, scale0(0.f));
    const auto* hb = param(hc, "bias", {Cnet}, 1, 1.f);
    // C_out = 6 (or 3) is padded to one 32-wide N tile; only the real columns are stored (n_store)
    const int npad = 32;
    const int pf = needs_pairing(a.C) ? 2 : 1;          // nf = 32: pixel-paired, 2 x C_out columns per super-pixel
    if (pf == 2 && a.W % 2 != 0) return fail("pixel pairing needs an even width");
    float head_wscale = 1.f;
    __half* wp = nullptr; const float* bp = nullptr;
    if (dry_) { wp = (__half*)w_alloc((size_t)npad * 9 * a.C * pf * 2); w_alloc(npad * 4); }
    else {
      if (!hk || !hb) return -1;
      // The reference initialises this layer with scale 1e-10 (init_scale = 0): bring the weights into fp16's
      // normal range with an exact power-of-two factor and undo it in the epilogue scale.
      float wmax = 0.f;
      for (float v : *hk) wmax = std::max(wmax, std::fabs(v));
      int e = 0;
      if (wmax > 0.f) std::frexp(wmax, &e);
      head_wscale = std::ldexp(1.0f, -e);                         // wmax * head_wscale in [0.5, 1)
      std::vector<__half> pk((size_t)npad * 9 * a.C * pf, __float2half(0.f));
      if (pf == 2) pack_conv_paired(*hk, 9, a.C, Cnet, pk, 9 * a.C * pf, 0, head_wscale);
      else pack_conv(*hk, 9, a.C, Cnet, pk, 9 * a.C, 0, head_wscale);
      std::vector<float> bb(npad, 0.f);
      for (int c = 0; c < pf; ++c)
        for (int co = 0; co < Cnet; ++co) bb[c * Cnet + co] = (*hb)[co] * head_wscale;
      wp = upload_f16(pk); bp = upload_f32(bb);
    }
    Op op; op.kind = OP_GEMM; op.tag = "head";
    op.out_is_external = true;
    op.gemm = make_gemm(max_batch_, a.H, a.W / pf);
    GemmOp& g = op.gemm;
    g.nseg = 1; g.seg[0] = {a.p, a.C * pf, 0, a.C * pf, 9};
    g.w = wp; g.N = npad; g.w_ld = 9 * a.C * pf; g.bias = bp;
    g.out32 = reinterpret_cast<float*>(uintptr_t(16));   // patched with the caller's output at launch
    g.ldo = Cnet * pf; g.n_store = Cnet * pf;
    g.scale = 1.0f / head_wscale;
    ops_.push_back(op);
    rel(a);
  }
  if (m.conditional) {
    if (dry_) { w_alloc((size_t)temb_dim_ * temb_total_ * 4); w_alloc((size_t)temb_total_ * 4); }
    else { d_proj_w_ = upload_f32(proj_w_host_); d_proj_b_ = upload_f32(proj_b_host_); }
  }
  return 0;
}

int UNet::finalize() {
  if (!err_.empty()) return -1;
  if (finalized_) return 0;
  for (const auto& sp : specs_)
    if (!host_params_.count(sp.name)) return fail("gddim_ctx_finalize: parameter not set: " + sp.name);
  const size_t wbytes = weight_top_ + (size_t)temb_total_ * 4 + 4096;
  const size_t abytes = arena_peak_ + 4096;
  const int total = temb_total_;
  wts_ = nullptr; arena_ = nullptr;
  dry_ = false;
  if (cudaMalloc(&wts_, wbytes) != cudaSuccess) { wts_ = nullptr; return fail("cudaMalloc(weights) failed"); }
  if (cudaMalloc(&arena_, abytes) != cudaSuccess) { arena_ = nullptr; return fail("cudaMalloc(workspace) failed"); }
  cudaMemset(wts_, 0, wbytes);
  weight_bytes_ = wbytes;
  arena_bytes_ = abytes;
  proj_w_host_.assign((size_t)temb_dim_ * total, 0.f);
  proj_b_host_.assign(total, 0.f);
  dry_ = false;
  if (walk() != 0) return -1;
  if (arena_top_ > arena_peak_) return fail("internal: workspace plan mismatch");
  host_params_.clear();
  proj_w_host_.clear(); proj_w_host_.shrink_to_fit();
  {
    // Sweep directions: every GEMM / GroupNorm-apply pass starts at the end its main input was written LAST, so
    // the tail of the producer's output is still in L2 when it is read (the 32x32-level tensors exceed the L2;
    // same-direction sweeps get no hits).  GDDIM_ZIGZAG=0 keeps every pass ascending (A/B measurements).
    const char* e = getenv("GDDIM_ZIGZAG");
    const bool zigzag = !(e && e[0] == '0');
    std::map<const void*, int> wdir;                 // buffer -> 1 if its last writer ran in descending order
    auto dir_of = [&](const void* p) { auto it = wdir.find(p); return it == wdir.end() ? 0 : it->second; };
    for (auto& op : ops_) {
      if (op.kind == OP_GEMM) {
        GemmOp& g = op.gemm;
        g.reverse = zigzag ? !dir_of(g.seg[0].ptr) : 0;
        if (g.out32) wdir[g.out32] = g.reverse;
        if (g.out16) wdir[g.out16] = g.reverse;
      } else if (op.kind == OP_NORM) {
        NormOp& n = op.norm;
        n.reverse = zigzag ? !dir_of(n.src1) : 0;
        if (n.dst16) wdir[n.dst16] = n.reverse;
        if (n.raw16) wdir[n.raw16] = n.reverse;
      } else if (op.kind == OP_GN_QKV) {
        op.gq.reverse = zigzag ? !dir_of(op.gq.x) : 0;
        wdir[op.gq.out16] = op.gq.reverse;
      } else if (op.kind == OP_ATTN_FUSED) {
        op.attn.reverse = zigzag ? !dir_of(op.attn.qkv) : 0;
        if (op.attn.out16) wdir[op.attn.out16] = op.attn.reverse;
        if (op.attn.gn_out16) wdir[op.attn.gn_out16] = op.attn.reverse;
        if (op.attn.out32) wdir[op.attn.out32] = op.attn.reverse;
      } else {
        if (op.f_out) wdir[op.f_out] = 0;
        if (op.h_out) wdir[op.h_out] = 0;
      }
    }
  }
  for (auto& op : ops_) {
    if (op.kind == OP_GEMM) {
      if (gemm_prepare(&op.gemm, 0) != 0) return fail(std::string("gemm_prepare(") + op.tag + "): " + gemm_last_error());
    } else if (op.kind == OP_GN_QKV) {
      if (gn_qkv_prepare(&op.gq) != 0) return fail(std::string("gn_qkv_prepare(") + op.tag + "): " + gemm_last_error());
    } else if (op.kind == OP_ATTN_FUSED) {
      if (attn_fused_prepare(&op.attn) != 0) return fail(std::string("attn_fused_prepare(") + op.tag + "): " + gemm_last_error());
    }
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return fail("device error during finalize");
  finalized_ = true;
  return 0;
}

int UNet::time_projections(double t, float* dst_dev, cudaStream_t st) {
  if (!finalized_) return fail("context not finalized");
  if (!cfg_.conditional) return 0;
  const int nf = cfg_.nf;
  std::vector<float> e(emb_in_dim_);
  const float labels = (float)(999.0 * t);     // models/utils.py:172 (fp32 in the reference)
  if (cfg_.embedding_type == 0) {
    // layerspp.py:33-43 on log(labels) (ncsnpp.py:71-77): fp32 argument, then sin/cos
    const float lg = (float)std::log((double)labels);
    for (int i = 0; i < nf; ++i) {
      const float arg = lg * fourier_w_[i] * 2.0f * (float)M_PI;
      e[i] = (float)std::sin((double)arg);
      e[nf + i] = (float)std::cos((double)arg);
    }
  } else {
    // layers.py:450-464 get_timestep_embedding(labels, nf)
    const int half = nf / 2;
    const float k = (float)(std::log(10000.0) / (double)(half - 1));
    for (int i = 0; i < half; ++i) {
      const float f = (float)std::exp((double)((float)i * -k));
      const float arg = labels * f;
      e[i] = (float)std::sin((double)arg);
      e[half + i] = (float)std::cos((double)arg);
    }
  }
  if (cudaMemcpyAsync(d_temb0_, e.data(), e.size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess)
    return fail("temb upload failed");
  if (dense_launch(d_temb0_, d_dense0_w_, d_dense0_b_, d_temb1_, 1, emb_in_dim_, temb_dim_, 0, st)) return fail("dense0");
  if (dense_launch(d_temb1_, d_dense1_w_, d_dense1_b_, d_temb2_, 1, temb_dim_, temb_dim_, 1, st)) return fail("dense1");
  if (dense_launch(d_temb2_, d_proj_w_, d_proj_b_, dst_dev, 1, temb_dim_, temb_total_, 1, st)) return fail("dense proj");
  launches_ += 3;
  return 0;
}

int UNet::forward(const float* x_dev, float* out_dev, int batch, cudaStream_t st, const CldStepArgs* head_update) {
  if (!finalized_) return fail("context not finalized");
  if (batch < 1 || batch > max_batch_) return fail("batch exceeds max_batch");
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cap);
  const bool prof = profile_ && cap == cudaStreamCaptureStatusNone;
  if (prof && prof_ev_.size() != ops_.size() + 1) {
    prof_ev_.resize(ops_.size() + 1);
    for (auto& e : prof_ev_) cudaEventCreate(&e);
    prof_op_ms_.assign(ops_.size(), 0.0);
    prof_op_flops_.assign(ops_.size(), 0.0);
    prof_op_bytes_.assign(ops_.size(), 0.0);
  }
  size_t op_idx = 0;
  if (prof) cudaEventRecord(prof_ev_[0], st);
  for (auto& op : ops_) {
    int rc = 0;
    switch (op.kind) {
      case OP_NORM: {
        NormOp n = op.norm;
        n.B = batch;
        rc = norm_launch(&n, st);
        launches_ += norm_num_launches(&n);
        break;
      }
      case OP_GEMM: {
        GemmOp g = op.gemm;
        g.B = batch;
        if (op.out_is_external) g.out32 = out_dev;
        g.m_tiles = (int)(((long long)batch * g.H * g.W + 128 * g.m_sub - 1) / (128 * g.m_sub));
        bool update_behind = false;
        if (op.out_is_external && head_update != nullptr) {
          static const bool no_fuse = [] { const char* e = getenv("GDDIM_NO_HEAD_UPDATE"); return e && e[0] == '1'; }();   // A/B switch
          if (!no_fuse && gemm_impl == 0 && gemm_head_update_supported(&g, head_update)) g.upd = head_update;
          else update_behind = true;
        }
        rc = gemm_launch(&g, gemm_impl, st);
        if (rc) return fail(std::string("gemm_launch(") + op.tag + "): " + gemm_last_error());
        if (update_behind) {
          if (cld_step_launch(head_update, st)) return fail("cld_step launch failed");
          launches_ += 1;
        }
        if (gemm_impl == 1 || g.cuda_core)     // CUDA-core path: column statistics / softmax are separate kernels
          launches_ += 1 + ((g.colstats && g.out32) ? 1 : 0) + (g.epi == EPI_SOFTMAX ? 1 : 0) + (g.epi == EPI_GNF ? 1 : 0);
        else
          launches_ += 1;
        break;
      }
      case OP_IM2COL:
        if (op.use_fir == 2)
          rc = im2col_same3x3_launch(x_dev, op.h_out, batch, op.H, op.W, op.cin, op.kpad, kRawScale, 1, st);
        else
          rc = im2col_fir_down_launch(op.in_is_external ? x_dev : op.f_in, op.h_out, batch, op.H, op.W, op.cin, op.kpad,
                                      op.use_fir, kRawScale, st);
        launches_ += 1;
        break;
      case OP_TRANSPOSE_V:
        rc = transpose_v_launch(op.h_in, op.h_out, batch, op.T, op.cin, op.ld, op.voff, st);
        launches_ += 1;
        break;
      case OP_SOFTMAX_ROWS:
        rc = softmax_rows_launch(op.f_in, op.h_out, op.f_out, (long long)batch * op.T, op.T, st);
        launches_ += 1;
        break;
      case OP_GN_QKV:
        rc = gn_qkv_launch(&op.gq, batch, st);
        launches_ += 1;
        break;
      case OP_ATTN_FUSED:
        rc = attn_fused_launch(&op.attn, batch, st);
        launches_ += 1;
        break;
      case OP_SMALL_ATTN:
        rc = small_attn_launch(op.h_in, op.h_out, batch, op.T, op.cin, op.scale, st);
        launches_ += 1;
        break;
      case OP_STEM: case OP_HEAD:       // never planned (unet.h): stem and head are OP_GEMM
        return fail("internal: unplanned op kind in " + op.tag);
    }
    if (rc) return fail("launch failed in op " + op.tag + " (rc=" + std::to_string(rc) + ")");
    ++op_idx;
    if (prof) cudaEventRecord(prof_ev_[op_idx], st);
  }
  if (prof) {
    if (cudaStreamSynchronize(st) != cudaSuccess) return fail("device error during profiled forward");
    for (size_t i = 0; i < ops_.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, prof_ev_[i], prof_ev_[i + 1]);
      prof_op_ms_[i] += ms;
      if (ops_[i].kind == OP_GN_QKV)
        prof_op_flops_[i] = 2.0 * batch * (double)ops_[i].gq.T * ops_[i].gq.C * ops_[i].gq.N;
      if (ops_[i].kind == OP_ATTN_FUSED)
        prof_op_flops_[i] = 4.0 * batch * (double)ops_[i].attn.T * ops_[i].attn.T * ops_[i].attn.C +   // QK^T + P.V
                            (ops_[i].attn.w3 ? 2.0 * batch * (double)ops_[i].attn.T * ops_[i].attn.C * ops_[i].attn.C : 0.0);
      if (ops_[i].kind == OP_GEMM) {
        const GemmOp& g = ops_[i].gemm;
        double k = 0;
        for (int s2 = 0; s2 < g.nseg; ++s2) k += (double)g.seg[s2].taps * g.seg[s2].c;
        prof_op_flops_[i] = 2.0 * batch * g.H * g.W * (double)g.N * k;
      }
      if (ops_[i].kind == OP_NORM) {
        // algorithmic HBM bytes of the GroupNorm(+swish, +resample) pass: fp32 source read once (twice when the
        // statistics do not come from the producer's epilogue), fp16 outputs written once
        const NormOp& n = ops_[i].norm;
        const double in_el = (double)batch * n.H * n.W * (n.c1 + n.c2);
        const double f = (n.resample == RS_FIR_DOWN || n.resample == RS_NAIVE_DOWN) ? 0.25
                       : ((n.resample == RS_FIR_UP || n.resample == RS_NAIVE_UP) ? 4.0 : 1.0);
        const bool stats_pass = n.colstats1 == nullptr || (n.src2 != nullptr && n.colstats2 == nullptr);
        double b = 0;
        if (!n.coef_only || stats_pass) b += 4.0 * in_el * ((stats_pass && !n.coef_only) ? 2.0 : 1.0);
        if (n.dst16) b += 2.0 * in_el * f;
        if (n.raw16) b += 2.0 * in_el * f;
        prof_op_bytes_[i] = b;
      }
    }
    ++prof_forwards_;
  }
  return 0;
}

void UNet::set_profile(bool on) {
  profile_ = on;
  std::fill(prof_op_ms_.begin(), prof_op_ms_.end(), 0.0);
  prof_forwards_ = 0;
}

void UNet::get_profile(double ms_by_kind[8], double* gemm_flops, long long* gemm_launches) const {
  for (int i = 0; i < 8; ++i) ms_by_kind[i] = 0;
  double fl = 0;
  long long nl = 0;
  for (size_t i = 0; i < prof_op_ms_.size(); ++i) {
    const OpKind k = ops_[i].kind;       // one attention family; the GroupNorm-fused projection counts as a GEMM
    ms_by_kind[k == OP_ATTN_FUSED ? (int)OP_SMALL_ATTN : (k == OP_GN_QKV ? (int)OP_GEMM : (int)k)] += prof_op_ms_[i];
    if (k == OP_GN_QKV) { fl += prof_op_flops_[i] * prof_forwards_; nl += prof_forwards_; }
    if (ops_[i].kind == OP_GEMM) { fl += prof_op_flops_[i] * prof_forwards_; nl += prof_forwards_; }
  }
  if (gemm_flops) *gemm_flops = fl;
  if (gemm_launches) *gemm_launches = nl;
}

double UNet::profile_norm_bytes() const {
  double b = 0;
  for (size_t i = 0; i < prof_op_bytes_.size(); ++i)
    if (ops_[i].kind == OP_NORM) b += prof_op_bytes_[i] * prof_forwards_;
  return b;
}

int UNet::dump_profile(const char* path) const {
  FILE* f = fopen(path, "w");
  if (!f) return -1;
  fprintf(f, "op,kind,H,W,N,K,block_n,ms_per_forward,gflop,tflops,mbytes,gbs\n");
  for (size_t i = 0; i < prof_op_ms_.size(); ++i) {
    const Op& op = ops_[i];
    const double ms = prof_forwards_ ? prof_op_ms_[i] / prof_forwards_ : 0.0;
    double k = 0;
    int H = op.H, W = op.W, N = op.cout, bn = 0;
    if (op.kind == OP_GEMM) {
      for (int s2 = 0; s2 < op.gemm.nseg; ++s2) k += (double)op.gemm.seg[s2].taps * op.gemm.seg[s2].c;
      H = op.gemm.H; W = op.gemm.W; N = op.gemm.N; bn = op.gemm.block_n;
    } else if (op.kind == OP_NORM) {
      H = op.norm.H; W = op.norm.W; N = op.norm.c1 + op.norm.c2;
    }
    const double by = i < prof_op_bytes_.size() ? prof_op_bytes_[i] : 0.0;
    fprintf(f, "%s,%d,%d,%d,%d,%.0f,%d,%.5f,%.4f,%.2f,%.3f,%.1f\n", op.tag.c_str(), (int)op.kind, H, W, N, k, bn, ms,
            prof_op_flops_[i] * 1e-9, ms > 0 ? prof_op_flops_[i] / (ms * 1e-3) * 1e-12 : 0.0, by * 1e-6,
            ms > 0 ? by / (ms * 1e-3) * 1e-9 : 0.0);
  }
  fclose(f);
  return 0;
}

}  // namespace gddim
