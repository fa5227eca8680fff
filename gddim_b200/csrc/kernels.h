// Host-side declarations of every device op of the sampling hot path (sm_100a).
// Tensors are NHWC; "T32" = fp32 trunk / conv outputs, "T16" = fp16 GEMM operands.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace gddim {

struct CldStepArgs;

// ---- implicit-GEMM convolution / GEMM (conv_gemm.cu) -------------------------------------------
// out[m, n] = epilogue( sum_seg sum_tap sum_c A_seg[pixel(m)+tap, c] * Wt[n, k(seg,tap,c)] )
//   m enumerates pixels of a [B,H,W] grid in NHW order; taps = 9 -> 3x3 SAME window, 1 -> pointwise.
//   Wt is K-major: [N, Ktot] fp16 (optionally one matrix per batch item: attention QK^T / PV).
// Linear epilogue: out = (acc * rowscale[m] + bias[n] + bias2[n] + residual[m,n]) * scale
// Softmax epilogue (N == block_n == 256): out16 = exp(acc*scale - rowmax), row_out[m] = 1/rowsum.
struct GemmSeg {
  const __half* ptr;   // [B,H,W,c_total]
  int c_total;         // channel count of the tensor
  int c_off;           // first channel used
  int c;               // channels used (multiple of 64 for the tcgen05 kernel; multiples of 16 run on CUDA cores)
  int taps;            // 1 or 9
};

enum { EPI_LINEAR = 0, EPI_SOFTMAX = 1, EPI_GNF = 2 };

struct GemmOp {
  GemmSeg seg[2];
  int nseg;
  int B, H, W;
  const __half* w;          // weights / B operand, K-major
  int N;                    // output columns
  int w_ld;                 // row stride of w in elements (>= Ktot)
  int w_koff;               // first K column used in w
  int wsplit;               // 0 / 1: one pass; 2: W = W_hi + W_lo, columns [w_koff, +K) and [w_koff + K, +2K) (precise mode)
  long long w_batch_stride; // elements between per-image B matrices; 0 = shared weights
  int w_rows_per_batch;     // rows (N) per batch matrix when batched
  const float* bias;
  const float* bias2;
  const float* residual;    // [M, ldo] fp32
  const float* rowscale;    // [M]
  float scale;
  float* out32;             // [M, ldo]
  __half* out16;            // [M, ldo]
  float* row_out;           // [M] (softmax)
  int ldo;
  int n_store;              // 0 = all N columns; else only columns < n_store are written (fp32, scalar stores,
                            // ldo may then be any value): few-channel outputs such as the 6-channel head conv
  int epi;
  // optional GroupNorm statistics of the output, fused into the epilogue: per 32-row slab and column,
  // colstats[slab][0][n] = sum, colstats[slab][1][n] = sum of squares (slab = m / 32; rows >= M excluded)
  float* colstats;
  // tiles in descending order.  A consumer that starts with what its producer wrote last finds it still in L2 (the
  // 32x32-level tensors are larger than the 126 MB L2, so same-direction sweeps get no hits at all); unet.cpp
  // alternates the direction along every producer -> consumer chain.  Results do not depend on it.
  int reverse;
  // EPI_GNF: out16 = act(GroupNorm(linear epilogue value)) -- the GroupNorm that follows the convolution is applied by
  // its own epilogue (two passes over the TMEM accumulator, statistics exchanged inside the CTA / across the cluster);
  // out32, colstats, residual and rowscale must be null.  gemm_gnf_supported() tells which geometries qualify.
  const float* gn_gamma;    // [N]
  const float* gn_beta;     // [N]
  float gn_eps;
  int gn_groups;            // groups over the N output channels (channels per group must be 4, 8 or 16)
  int gn_silu;
  // ---- filled by gemm_prepare ----
  CUtensorMap tmA[2];
  CUtensorMap tmB;
  int block_n;
  int m_sub;                // 128-row sub-tiles per CTA tile (1 or 2)
  int cg;                   // CTAs per MMA (cta_group): 2 = cluster of two CTAs, each staging half the weight tile
  int halo;                 // 3x3 conv with halo tiles: one (rows + 2)-row box per x-shift, y-shifts by descriptor offset
  int stages, stage_bytes, a_bytes;   // halo: smem ring geometry
  CUtensorMap tmH;          // halo box of segment 0
  int m_tiles, n_tiles, tiles_per_batch;
  int gn_xc;                // EPI_GNF: CTAs per image = cluster size (1, 2, 4)
  // optional: the CLD sampler's update applied by this launch's epilogue (head convolution only; see
  // gemm_head_update_supported).  Host pointer, read at launch.
  const CldStepArgs* upd;
  int cuda_core;            // a segment's channel count is not a multiple of 64: the op runs on the CUDA-core kernel
                            // whatever `impl` says; linear epilogue only (operator ABI -- the network planner pairs pixels
                            // instead, unet.cpp pack_conv_paired)
  int prepared;
};

// EPI_GNF is available for H*W in {16, 64, 256} (any N that is a multiple of 32) and for H*W = 1024 with N in {64, 128};
// channels per group 4, 8 or 16
int gemm_gnf_supported(int H, int W, int N, int groups);

// Encodes the TMA descriptors (needs the final device addresses). Returns 0 or a negative error.
int gemm_prepare(GemmOp* op, int force_block_n, int force_m_sub = 0, int force_cg = 0);
int gemm_head_update_supported(const GemmOp* op, const CldStepArgs* u);
// impl: 0 = tcgen05/TMA kernel, 1 = CUDA-core reference kernel (validation only)
int gemm_launch(const GemmOp* op, int impl, cudaStream_t st);
const char* gemm_last_error();

// 128B-swizzled fp16 tensor map of rank 2..4 over a dense tensor (dims innermost first, box in elements)
int tmap_encode_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint32_t* box);

// ---- fused attention for T = 256 tokens, C = 256 channels (attn.cu) ------------------------------------------------
// out16[b, t, :] = softmax_s( q[b,t,:] . k[b,s,:] * scale ) v[b,s,:]   with q, k, v = channel thirds of qkv16 [B, T, 3C]
struct AttnOp {
  const __half* qkv;   // [B, T, 3C]
  __half* out16;       // [B, T, C]
  int B, T, C;
  float scale;
  int reverse;         // CTAs walk the batch in descending order (see GemmOp::reverse)
  // optional fused output projection + residual (AttnBlockpp NIN_3, layerspp.py:79-83):
  //   out32 = (out @ w3^T + residual) * out_scale + bias3 * out_scale, plus the column statistics of out32
  const __half* w3;    // [C, C] fp16 K-major ([C_out][C_in]); null = attention only (out16)
  const float* bias3;  // [C]
  const float* residual;   // [B, T, C] fp32
  float* out32;        // [B, T, C]
  float* colstats;     // [B*T/32][2][C]
  float out_scale;
  // optional (with w3): out16n = act(GroupNorm(out32)) -- the GroupNorm_0 of the ResBlock that consumes the attention
  // block's output (layerspp.py:196), applied by this kernel's epilogue (dual GroupNorm epilogue, gemm_epilogue.cuh)
  const float* gn_gamma; const float* gn_beta; float gn_eps; int gn_groups; int gn_silu;
  __half* gn_out16;    // [B, T, C]
  CUtensorMap tm_qkv, tm_w3;  // filled by attn_fused_prepare
  int prepared;
};
int attn_fused_supported(int T, int C);
int attn_fused_prepare(AttnOp* op);
int attn_fused_launch(const AttnOp* op, int batch, cudaStream_t st);

// ---- GroupNorm apply fused into the q/k/v projection (gn_qkv.cu; C = 256, N = 768, T % 128 == 0) ----------------------
// out16[m, :] = fp16(x[m, :] * coef_a[b(m), :] + coef_b[b(m), :]) @ w^T + bias, with coef from a coef_only NormOp
struct GnQkvOp {
  const float* x;      // [B, T, C] fp32
  const float* coef;   // [B, 2, C]
  const __half* w;     // [N, C] fp16 K-major
  const float* bias;   // [N]
  __half* out16;       // [B, T, N]
  int B, T, C, N;
  int reverse;
  CUtensorMap tm_w;    // filled by gn_qkv_prepare
  int prepared;
};
int gn_qkv_supported(int T, int C, int N);
int gn_qkv_prepare(GnQkvOp* op);
int gn_qkv_launch(const GnQkvOp* op, int batch, cudaStream_t st);

// ---- GroupNorm (+SiLU) (+FIR / naive resampling) (norm.cu) ---------------------------------------
enum { RS_NONE = 0, RS_FIR_DOWN = 1, RS_FIR_UP = 2, RS_NAIVE_DOWN = 3, RS_NAIVE_UP = 4 };

struct NormOp {
  const float* src1; int c1;     // [B,H,W,c1]
  const float* src2; int c2;     // optional second source, channel-concatenated after src1
  int B, H, W;                   // input grid
  int groups;
  const float* gamma; const float* beta;   // [c1+c2]
  float eps;
  int silu;                      // apply x*sigmoid(x) after the affine
  int resample;                  // RS_*
  const float* colstats1;        // producer-written column statistics of src1 / src2 (see GemmOp); when both
  const float* colstats2;        // present (src2 optional) the stats pass over the tensor is skipped
  float* partial;                // [B, splits, groups, 2] scratch
  float* coef;                   // [B, 2, C] scratch: per-channel scale / shift
  unsigned int* ticket;          // [B] zero-initialised slab counters (self-resetting)
  int splits;
  __half* dst16;                 // normalised (+act, +resample) output  [B,H',W',C]; may be null
  __half* raw16;                 // raw (resampled) copy of the input in fp16; may be null
  float raw_scale;               // raw16 = x * raw_scale (power of two: headroom against fp16 overflow)
  int reverse;                   // apply pass walks images / pixel chunks in descending order (see GemmOp::reverse)
  int coef_only;                 // compute the scale / shift table `coef` only; the consumer applies it (gn_qkv.cu)
};
int norm_launch(const NormOp* op, cudaStream_t st);
int norm_num_launches(const NormOp* op);   // kernels norm_launch issues for this op (1 or 2)
int norm_splits(int B, int H, int W);

// ---- gathers and data movement (small.cu) ---------------------------------------------------------
// FIR (pad 2, [1,3,3,1]x[1,3,3,1]/64) followed by the 3x3 stride-2 VALID window gather:
// in fp32 [B,H,W,c] -> A16 [B,H/2,W/2,kpad] with k = tap*c + ch (zero padded to kpad)
int im2col_fir_down_launch(const float* in, __half* a16, int B, int H, int W, int c, int kpad, int use_fir,
                           float out_scale, cudaStream_t st);
// 3x3 SAME window gather of a few-channel fp32 image (the stem): in [B,H,W,c] -> A16 [B,H,W,kpad], k = tap*c + ch
// split = 1: kpad = 3 segments (hi(x), lo(x), hi(x)) for an fp32-accurate product with (hi(w), hi(w), lo(w)) rows
int im2col_same3x3_launch(const float* in, __half* a16, int B, int H, int W, int c, int kpad, float out_scale,
                          int split, cudaStream_t st);
// V^T per image for the PV GEMM: qkv16 [B,T,ld] (V at channel voff) -> vT [B,C,T]
int transpose_v_launch(const __half* qkv, __half* vT, int B, int T, int C, int ld, int voff, cudaStream_t st);
// Full attention for very short sequences (T <= 64): qkv16 [B,T,3C] -> o16 [B,T,C]
// row softmax over T keys (long-sequence attention): p16 = exp(s - rowmax), rowinv = 1 / sum
int softmax_rows_launch(const float* s32, __half* p16, float* rowinv, long long rows, int T, cudaStream_t st);
int small_attn_launch(const __half* qkv, __half* o16, int B, int T, int C, float scale, cudaStream_t st);
// y[r, n] = bias[n] + sum_k act(x[r,k]) * w[k,n]  (time-embedding MLP and per-block projections; fp32)
int dense_launch(const float* x, const float* w, const float* bias, float* y, int rows, int K, int N, int silu_in,
                 cudaStream_t st);
int add_vec_launch(const float* a, const float* b, float* y, int n, cudaStream_t st);

// ---- sampler updates (update.cu) --------------------------------------------------------------------
// State in "net layout" [B,H,W,2C]: channel d (< C) = x_d, channel C+d = v_d.
// u' = A u + sum_j Cj e_j   with 2x2 matrices acting on (x_d, v_d); e_0 = eps_new (+ M u if mixed).
struct CldStepArgs {
  const float* u; float* u_out;
  const float* eps[6];         // eps[0] newest
  int n_eps;                   // number of eps terms with (possibly) non-zero coefficient
  float coef[7][4];            // coef[0] = A, coef[1+j] = C_j (row-major 2x2)
  int mixed; float mixm[4];    // eps_0 += M u, written back to eps_store
  float* eps_store;
  long long n_pix; int C;
  // stochastic gDDIM: u' += F z, z ~ N(0, I_2) per (pixel, channel) pair
  int noise_mode;              // 0 none, 1 from `noise` (reference layout [n_pix, C, 2]), 2 Philox4x32-10
  const float* noise;
  float nfac[4];               // F (row-major 2x2)
  unsigned long long seed, stream_id;   // Philox key / per-step counter word
};
int cld_step_launch(const CldStepArgs* a, cudaStream_t st);
// x = u_x * mul + add, v = u_v : [B,H,W,2C] -> x[B,H,W,C], v[B,H,W,C]
int cld_split_launch(const float* u, float* x, float* v, long long n_pix, int C, float mul, float add, cudaStream_t st);
// reference layout [..., C, 2] <-> net layout [..., 2C]
int relayout_launch(const float* src, float* dst, long long n_pix, int C, int to_net, cudaStream_t st);
// Reference-API update on reference-layout arrays (deis.multistep_ab_step): x,new_eps [n,2]; hist [order+1,n,2]
int ab_step_ref_layout_launch(const float* x, const float* coef /*[(order+3)*4]*/, const float* new_eps,
                              const float* hist, float* x_out, float* hist_out, int order, long long n_pairs,
                              cudaStream_t st);

// 2-D orthonormal DCT-II (fwd=1) / DCT-III (fwd=0) on 32x32 NHWC planes
int dct32_launch(const float* in, float* out, int B, int C, int fwd, cudaStream_t st);
// Blur DDIM step: y' = a .* y + b .* DCT(eps_x);  x_next = IDCT(y')   (a, b: [32,32])
int blur_step_launch(const float* y, const float* eps_x, const float* a, const float* b, float* y_out,
                     float* x_next, int B, int C, cudaStream_t st);
int scale_shift_launch(const float* in, float* out, long long n, float mul, float add, cudaStream_t st);
// Scalar-coefficient AB step (blur multistep.ab_step): x' = c0 x + sum_j c_{1+j} e_j
int scalar_ab_step_launch(const float* x, const float* coef, const float* new_eps, const float* hist, float* x_out,
                          float* hist_out, int n_hist, long long n, cudaStream_t st);

}  // namespace gddim
