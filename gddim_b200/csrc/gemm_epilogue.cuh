// Shared pieces of the tcgen05 GEMM kernels: tile constants, kernel arguments, shared-memory layout and the linear
// epilogue (TMEM -> swizzled shared memory -> full-line global accesses with bias / residual / scale / column
// statistics).  Used by conv_gemm.cu and by the output projection fused into attn.cu.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>

#include "ptx.cuh"
#include "cld_update.cuh"

#ifdef GDDIM_ABLATE
#define GDDIM_DBG_STORE(p) ((p).dbg != 2)
#define GDDIM_DBG_IS(p, k) ((p).dbg == (k))
// clock64 stamp `slot` of tile `t` (CTA 0, one lane): timeline experiments only
#define GDDIM_STAMP(p, cond, t, slot) do { if ((p).dbg_clk && blockIdx.x == 0 && (cond) && (t) < 16) (p).dbg_clk[(t) * 16 + (slot)] = clock64(); } while (0)
#else
#define GDDIM_DBG_STORE(p) true
#define GDDIM_DBG_IS(p, k) false
#define GDDIM_STAMP(p, cond, t, slot) do { } while (0)
#endif

namespace gddim {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                       // 64 fp16 = 128 bytes = one swizzle row
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int NUM_THREADS = 320;
// Warp roles.  The scheduler prefers the highest warp id among eligible warps of a sub-partition, so the two
// single-thread roles that must never wait for an issue slot get the highest ids: warps 0-7 epilogue (TMEM lane
// quadrant = warp id & 3; the two groups of four take alternate 32-column chunks of every tile, so each scheduler
// has two epilogue warps to hide TMEM / smem / store latency behind each other and a tile drains in half the
// time), warp 8 TMA producer, warp 9 MMA issuer + TMEM owner.
constexpr int EPI_WARPS = 8;
constexpr int EPI_GROUPS = 2;
constexpr int PRODUCER_THREAD = 256;
constexpr int MMA_WARP = 9;
constexpr int MMA_THREAD = 288;

struct GemmArgs {
  int taps[2], kch[2], coff[2];
  int nseg;
  int H, W;
  int M, N;
  int m_tiles, n_tiles, tiles_per_batch;
  int w_koff;
  int wsplit;        // 1, or 2: the K loop runs twice, the second time against the weight columns shifted by ktot (W = W_hi + W_lo)
  int ktot;          // K of one pass in elements
  const float* bias;
  const float* bias2;
  const float* residual;
  const float* rowscale;
  float scale;
  float* out32;
  __half* out16;
  float* row_out;
  int ldo;
  int n_store;
  float* colstats;
  int stages, stage_bytes, a_bytes;   // HALO kernels: smem ring geometry (depends on W)
  int reverse;   // tiles in descending order (the consumer of a tensor starts with what its producer wrote last: L2 hits)
  // EPI_GNF: GroupNorm (+ swish) of the OUTPUT fused into the epilogue (see epi_tile_gnf): out16 = act(GN(v))
  const float* gn_gamma;   // [N]
  const float* gn_beta;    // [N]
  float gn_eps;
  int gn_cpg;              // channels per group (4, 8 or 16)
  int gn_silu;
  int gn_rpi;              // rows (pixels) per image = H * W
  int gn_xc;               // CTAs an image spans (1, 2 or 4)
  // gn_xg != null: the gn_xc CTAs of an image are `gn_xc` CONSECUTIVE CTAs of the grid (not one cluster) and exchange
  // their statistics through this global buffer of tagged 64-bit words (see gnf_fold); null: one cluster, DSMEM exchange
  unsigned long long* gn_xg;
  // weights of this launch (contiguous [N, w_ld] fp16): every CTA asks L2 for its slice at kernel entry, so that the whole
  // matrix streams in from HBM as ONE parallel burst instead of k-block by k-block behind the ring's HBM round trips
  const void* pf_ptr;
  long long pf_bytes;
  // head convolution of a CLD network inside a deterministic sampler: the gDDIM / DEIS update of the state is applied by
  // this launch's epilogue (epi_head_update) to the eps values it holds in registers -- the update kernel disappears
  int upd_on;
  CldUpd upd;
  long long* dbg_clk;   // GDDIM_ABLATE builds: clock64 timeline of CTA 0 ([tile < 16][16 stamps]), else null
  int dbg;   // GDDIM_GEMM_DBG (timing experiments only): 1 = epilogue drains TMEM only, 2 = no global stores, 3 = no TMEM reads
};

// MT = number of 128-row M sub-tiles a CTA tile covers (2 for narrow N: the weight tile is then shared by 256
// output rows, which halves the L2->smem operand traffic per FLOP of the N <= 128 layers)
// CG = CTAs cooperating on one MMA (cta_group): with CG = 2 a cluster of two CTAs computes 256 x BLOCK_N per MMA, each
// CTA staging its own 128 A rows and HALF of the weight tile -- a third less L2->smem traffic on the N = 256 layers
// GNF epilogue scratch (see epi_tile_gnf): per-warp group partials, per-(image, group) mean / rstd, double-buffered
// gamma / beta rows, the cluster exchange buffer
constexpr int GNF_GMAX = 32;                                   // groups per N tile (BLOCK_N / cpg <= 32 ... 64 for cpg 4, N 256: rejected)
constexpr int GNF_IMGS = 16;                                   // images per CTA tile (256 rows of 4x4 images)
constexpr int GNF_PSTAT_BYTES = 2 * 4 * 2 * GNF_GMAX * 8;      // [MT<=2][quad][seg][group] float2
constexpr int GNF_GSTAT_BYTES = GNF_IMGS * GNF_GMAX * 8;       // [image][group] (mean, rstd)
constexpr int GNF_XCHG_BYTES = 2 * 4 * GNF_GMAX * 8;           // [tile parity][cluster rank][group] (sum, sumsq)
template <int BLOCK_N>
struct GnfSmem {
  static constexpr int GB_BYTES = 2 * 2 * BLOCK_N * 4;         // [buf][gamma | beta][BLOCK_N]
  static constexpr int OFF_GSTAT = GNF_PSTAT_BYTES;
  static constexpr int OFF_XCHG = OFF_GSTAT + GNF_GSTAT_BYTES;
  static constexpr int OFF_GB = OFF_XCHG + GNF_XCHG_BYTES;
  static constexpr int OFF_BAR = OFF_GB + GB_BYTES;
  static constexpr int BYTES = OFF_BAR + 16;                    // two mbarriers
};

template <int BLOCK_N, int MT, int CG = 1, int EXTRA = 0>
struct SmemLayout {
  static constexpr int B_TILE_BYTES = (BLOCK_N / CG) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = MT * A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int BAR_BYTES = 256;
  // epilogue staging: 8 warps x 32 rows x 32 fp32, XOR-swizzled in 16-byte units -- conflict-free 128-bit
  // transposition without padding (the BLOCK_N = 256 layout has < 1 KB to spare next to four 48 KB stages)
  static constexpr int EPI_ROW_FLOATS = 32;
  static constexpr int EPI_STAGE_BYTES = EPI_WARPS * 32 * EPI_ROW_FLOATS * 4;
  static constexpr int EPI_BYTES = EPI_STAGE_BYTES + 2 * BLOCK_N * 4 + EXTRA;   // + double-buffered (bias + bias2) row (+ GNF scratch)
  static constexpr int AVAIL = SMEM_BUDGET - BAR_BYTES - EPI_BYTES;
  static constexpr int STAGES = (AVAIL / STAGE_BYTES) > 8 ? 8 : (AVAIL / STAGE_BYTES);
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + EPI_BYTES;   // dynamic smem is declared 1024-aligned
};


// ---- linear epilogue --------------------------------------------------------------------------------------------
// TMEM -> registers (thread = row) -> swizzled smem -> registers (8 lanes = one 32-column row segment), so that
// every global access is a full 128-byte line.  Two warps per scheduler run this, so it is written for a low
// instruction count: the variant (residual / fp32 out / fp16 out / column statistics / row scale / full tile) is a
// template parameter, and everything that does not depend on the accumulator (bias row, first residual chunk) is
// fetched before waiting for the MMAs; the residual of chunk q+1 is in flight while chunk q is processed.
template <int BLOCK_N, int MT>
struct EpiCtx {
  const GemmArgs& p;
  float* stg;
  float* bias_s;
  uint64_t* tfull;
  uint32_t tfull_phase;
  uint32_t taddr;          // TMEM address of this warp's lane quadrant, first column of the accumulator stage
  long long m0;            // first row of this warp in sub-tile 0
  int n_tile0;             // first output column of the tile
  int lane;
  int group;               // epilogue group: takes the 32-column chunks q = group, group + 2, ...
  int tile_seq = 0;        // tiles this CTA has processed before (timeline stamps)
  int warp_id = -1;
};

__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// GNP: additionally reduce the column sums to GroupNorm (segment, group) partials in shared memory (gnp_pstat, layout of
// GnfCtx::pstat) -- pass 1 of the dual GroupNorm epilogue, see epi_tile_gnf_dual
template <int BLOCK_N, int MT, bool RES_, bool O32_, bool O16_, bool STATS_, bool RSCALE_, bool FULL, bool GENERIC,
          bool NARROW = false, bool GNP = false>
__device__ __forceinline__ void epi_tile(const EpiCtx<BLOCK_N, MT>& cx, float2* gnp_pstat = nullptr, int gnp_quad = 0) {
  constexpr int RS = SmemLayout<BLOCK_N, MT>::EPI_ROW_FLOATS;
  constexpr int NCH = BLOCK_N / 32;
  constexpr int NQ = MT * NCH;
  const GemmArgs& p = cx.p;
  // specialised variants know their features at compile time; the generic one tests the pointers
  const bool RES = GENERIC ? (p.residual != nullptr) : RES_;
  const bool O32 = GENERIC ? (p.out32 != nullptr) : O32_;
  const bool O16 = GNP ? false : (GENERIC ? (p.out16 != nullptr) : O16_);     // (GNP: out16 is the normalised output of pass 2)
  const bool STATS = GNP || (GENERIC ? (p.colstats != nullptr) : STATS_);
  const bool gnp_split = GNP && p.gn_rpi < 32;
  const bool RSCALE = GENERIC ? (p.rowscale != nullptr) : RSCALE_;
  const int lane = cx.lane;
  const int rsub = lane >> 3;
  const int c4 = (lane & 7) * 4;
  // staging rows are 128 bytes; the 16-byte unit u of row r lives at unit (u ^ (r & 7))
  const uint32_t stg_w = ptx::smem_u32(cx.stg) + lane * RS * 4;                 // this thread's row (write side)
  const uint32_t wx = lane & 7;
  // read side: rows rsub + 4 i, unit lane & 7; (row & 7) alternates between rsub and rsub + 4 with the parity of i
  const uint32_t stg_r0 = ptx::smem_u32(cx.stg) + rsub * RS * 4 + (((lane & 7) ^ rsub) << 4);
  const uint32_t stg_r1 = ptx::smem_u32(cx.stg) + (rsub + 4) * RS * 4 + (((lane & 7) ^ (rsub + 4)) << 4);
  const uint32_t bias_a = ptx::smem_u32(cx.bias_s) + c4 * 4;
  const long long ldo = p.ldo;
  const float scale = p.scale;

  auto load_res = [&](int q, float4 (&res)[8]) {
    const long long mb = cx.m0 + (long long)(q / NCH) * BLOCK_M + rsub;
    const float* base = p.residual + mb * ldo + cx.n_tile0 + (q % NCH) * 32 + c4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (FULL || mb + i * 4 < p.M) res[i] = __ldg(reinterpret_cast<const float4*>(base + (long long)(i * 4) * ldo));
      else res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  float4 res[8];
  if (RES && cx.group < NQ) load_res(cx.group, res);
  __syncwarp();
  GDDIM_STAMP(p, cx.warp_id == 0 && lane == 0, cx.tile_seq, 0);
  ptx::mbar_wait(cx.tfull, cx.tfull_phase);
  ptx::tc_fence_after();
  GDDIM_STAMP(p, cx.warp_id == 0 && lane == 0, cx.tile_seq, 1);
  uint32_t r[32];
#ifdef GDDIM_ABLATE
  if (p.dbg == 3) return;
#endif
#pragma unroll 1
  for (int q = cx.group; q < NQ; q += EPI_GROUPS) {
    const int mi = q / NCH, c0 = (q % NCH) * 32;
    ptx::tmem_ld_32x32b_x32(cx.taddr + mi * BLOCK_N + c0, r);
    // residual of this group's next chunk: each float4 is re-loaded in place right after it has been consumed
    const bool res_more = RES && (q + EPI_GROUPS < NQ);
    const float* res_nbase = nullptr;
    long long res_nmb = 0;
    if (res_more) {
      const int qn = q + EPI_GROUPS;
      res_nmb = cx.m0 + (long long)(qn / NCH) * BLOCK_M + rsub;
      res_nbase = p.residual + res_nmb * ldo + cx.n_tile0 + (qn % NCH) * 32 + c4;
    }
    ptx::tmem_ld_wait();
#ifdef GDDIM_ABLATE
    if (p.dbg == 1) continue;
#endif
#pragma unroll
    for (int j = 0; j < 8; ++j) sts128(stg_w + ((j ^ wx) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    __syncwarp();
    const long long mb = cx.m0 + (long long)mi * BLOCK_M + rsub;                 // this lane's first row
    const int n0 = cx.n_tile0 + c0 + c4;
    float4 bsum = lds128(bias_a + c0 * 4);
    bsum.x *= scale; bsum.y *= scale; bsum.z *= scale; bsum.w *= scale;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f), cq = make_float4(0.f, 0.f, 0.f, 0.f);
    float gs0 = 0.f, gq0 = 0.f, gs1 = 0.f, gq1 = 0.f;            // GNP, 16-row images: rows 0..15 / 16..31 of the slab
    float* o32 = O32 ? p.out32 + mb * ldo + n0 : nullptr;
    __half* o16 = O16 ? p.out16 + mb * ldo + n0 : nullptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 v = lds128(((i & 1) ? stg_r1 : stg_r0) + (i >> 1) * 8 * RS * 4);
      const bool ok = FULL || (mb + i * 4 < p.M);
      if (RSCALE) {
        const float rsc = ok ? __ldg(p.rowscale + mb + i * 4) : 1.0f;
        v.x *= rsc; v.y *= rsc; v.z *= rsc; v.w *= rsc;
      }
      if (RES) {
        v.x += res[i].x; v.y += res[i].y; v.z += res[i].z; v.w += res[i].w;
        if (res_more) {
          if (FULL || res_nmb + i * 4 < p.M) res[i] = __ldg(reinterpret_cast<const float4*>(res_nbase + (long long)(i * 4) * ldo));
          else res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      v.x = fmaf(v.x, scale, bsum.x); v.y = fmaf(v.y, scale, bsum.y);
      v.z = fmaf(v.z, scale, bsum.z); v.w = fmaf(v.w, scale, bsum.w);
      if (ok && GDDIM_DBG_STORE(p)) {
        if (STATS) {
          cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
          cq.x = fmaf(v.x, v.x, cq.x); cq.y = fmaf(v.y, v.y, cq.y); cq.z = fmaf(v.z, v.z, cq.z); cq.w = fmaf(v.w, v.w, cq.w);
          if (GNP && gnp_split) {
            const float s4 = (v.x + v.y) + (v.z + v.w);
            if (i >= 4) { gs1 += s4; gq1 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, gq1)))); }
            else { gs0 += s4; gq0 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, gq0)))); }
          }
        }
        if (NARROW) {
          float* q = o32 + (long long)(i * 4) * ldo;
          if (n0 + 0 < p.n_store) q[0] = v.x;
          if (n0 + 1 < p.n_store) q[1] = v.y;
          if (n0 + 2 < p.n_store) q[2] = v.z;
          if (n0 + 3 < p.n_store) q[3] = v.w;
        } else if (O32) *reinterpret_cast<float4*>(o32 + (long long)(i * 4) * ldo) = v;
        if (O16) {
          __half2 h0 = __floats2half2_rn(v.x, v.y);
          __half2 h1 = __floats2half2_rn(v.z, v.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0);
          pk.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(o16 + (long long)(i * 4) * ldo) = pk;
        }
      }
    }
    if (STATS) {
      // fold the 4 row sub-groups (lanes l, l+8, l+16, l+24): lanes 0..7 then own 32 rows x 4 columns
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
        cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
        cq.x += __shfl_xor_sync(0xffffffffu, cq.x, o); cq.y += __shfl_xor_sync(0xffffffffu, cq.y, o);
        cq.z += __shfl_xor_sync(0xffffffffu, cq.z, o); cq.w += __shfl_xor_sync(0xffffffffu, cq.w, o);
      }
      const long long mrow0 = cx.m0 + (long long)mi * BLOCK_M;
      if (lane < 8 && (FULL || mrow0 < p.M) && (!GNP || p.colstats != nullptr)) {
        float* cp = p.colstats + ((mrow0 >> 5) * 2) * ldo + n0;
        *reinterpret_cast<float4*>(cp) = cs;
        *reinterpret_cast<float4*>(cp + ldo) = cq;
      }
      if (GNP) {
        // (segment, group) partials of this warp's 32-row slab: 4 columns per lane -> one group (cpg 4) or half of one
        float s0 = (cs.x + cs.y) + (cs.z + cs.w), q0 = (cq.x + cq.y) + (cq.z + cq.w);
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
          gs0 += __shfl_xor_sync(0xffffffffu, gs0, o); gq0 += __shfl_xor_sync(0xffffffffu, gq0, o);
          gs1 += __shfl_xor_sync(0xffffffffu, gs1, o); gq1 += __shfl_xor_sync(0xffffffffu, gq1, o);
        }
        for (int o = 1; o * 4 < p.gn_cpg; o <<= 1) {          // lanes of one group: 2 (cpg 8) or 4 (cpg 16)
          s0 += __shfl_xor_sync(0xffffffffu, s0, o); q0 += __shfl_xor_sync(0xffffffffu, q0, o);
          gs0 += __shfl_xor_sync(0xffffffffu, gs0, o); gq0 += __shfl_xor_sync(0xffffffffu, gq0, o);
          gs1 += __shfl_xor_sync(0xffffffffu, gs1, o); gq1 += __shfl_xor_sync(0xffffffffu, gq1, o);
        }
        if (lane < 8 && ((lane * 4) % p.gn_cpg) == 0) {
          float2* dst = gnp_pstat + ((mi * 4 + gnp_quad) * 2) * GNF_GMAX + (c0 + c4) / p.gn_cpg;
          dst[0] = gnp_split ? make_float2(gs0, gq0) : make_float2(s0, q0);
          dst[GNF_GMAX] = make_float2(gs1, gq1);
        }
      }
    }
    __syncwarp();
  }
  GDDIM_STAMP(p, cx.warp_id == 0 && lane == 0, cx.tile_seq, 6);
}

// ---- head convolution + gDDIM update ---------------------------------------------------------------------------------
// The head convolution (ncsnpp.py:237) produces eps = [eps_x | eps_v], six columns of one 32-wide N tile.  Thread = row =
// pixel after the TMEM load, so a thread holds the complete eps of its pixel and can apply the sampler's update
//   eps_0 (+= M u, mixed score) -> ring slot;   u' = A u + sum_j C_j eps_j      (cld_jax/deis.py:141-151, sampling.py:30-39)
// right there: u (requested before the accumulator is waited for) and the earlier evaluations eps_1.. (24-byte rows) are read,
// u' and eps_0 written, all as 8-byte accesses of consecutive rows by consecutive lanes.  Same arithmetic as
// cld_step_c3_kernel (cld_update.cuh): the fused and the standalone update give bit-identical states.
template <int BLOCK_N, int MT>
__device__ __forceinline__ void epi_head_update(const EpiCtx<BLOCK_N, MT>& cx) {
  static_assert(BLOCK_N == 32 && MT == 1, "head tile: one 32-column chunk");
  const GemmArgs& p = cx.p;
  const CldUpd& up = p.upd;
  const long long m = cx.m0 + cx.lane;
  const bool ok = m < p.M;                              // ragged last tile: such lanes compute on row 0 and store nothing
  const long long mr = ok ? m : 0;
  float u[6];
  if (cx.group == 0) {
    const float2* q = reinterpret_cast<const float2*>(up.u + mr * 6);
    const float2 a = q[0], b = q[1], c = q[2];          // plain loads: u_out may be u
    u[0] = a.x; u[1] = a.y; u[2] = b.x; u[3] = b.y; u[4] = c.x; u[5] = c.y;
  }
  ptx::mbar_wait(cx.tfull, cx.tfull_phase);
  ptx::tc_fence_after();
  if (cx.group != 0) return;                            // one 32-column chunk per tile: the second epilogue group only waits
  uint32_t r[8];
  ptx::tmem_ld_32x32b_x8(cx.taddr, r);                  // (warp-wide) the six real columns + 2 of the padding
  float eh[3][6];
#pragma unroll
  for (int j = 1; j < 4; ++j) {         // (unused slots point at a valid row: unconditional loads, conditional sums)
    const float2* h = reinterpret_cast<const float2*>(up.eps[j] + mr * 6);
    const float2 ha = __ldg(h), hb = __ldg(h + 1), hc = __ldg(h + 2);
    eh[j - 1][0] = ha.x; eh[j - 1][1] = ha.y; eh[j - 1][2] = hb.x; eh[j - 1][3] = hb.y; eh[j - 1][4] = hc.x; eh[j - 1][5] = hc.y;
  }
  ptx::tmem_ld_wait();
  float e[6], acc[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) e[j] = fmaf(__uint_as_float(r[j]), p.scale, cx.bias_s[j] * p.scale);     // as epi_tile
  if (up.mixed) cld_px_mix(e, u, up.mixm[0], up.mixm[1], up.mixm[2], up.mixm[3]);
  if (ok) {
    float2* o = reinterpret_cast<float2*>(p.out32 + m * 6);
    o[0] = make_float2(e[0], e[1]); o[1] = make_float2(e[2], e[3]); o[2] = make_float2(e[4], e[5]);
  }
  cld_px_apply(acc, u, up.coef[0][0], up.coef[0][1], up.coef[0][2], up.coef[0][3]);
  cld_px_acc(acc, e, up.coef[1][0], up.coef[1][1], up.coef[1][2], up.coef[1][3]);
#pragma unroll
  for (int j = 1; j < 4; ++j)
    if (j < up.n_eps) cld_px_acc(acc, eh[j - 1], up.coef[1 + j][0], up.coef[1 + j][1], up.coef[1 + j][2], up.coef[1 + j][3]);
  if (ok) {
    float2* o = reinterpret_cast<float2*>(up.u_out + m * 6);
    o[0] = make_float2(acc[0], acc[1]); o[1] = make_float2(acc[2], acc[3]); o[2] = make_float2(acc[4], acc[5]);
  }
}

// ---- GroupNorm-fused epilogue (EPI_GNF) ----------------------------------------------------------------------------
// out16 = act(GroupNorm(v)), v = acc * scale + (bias + bias2) * scale: the normalisation that FOLLOWS a convolution
// (layerspp.py:218: h = act(GroupNorm_1(conv1(...) + Dense_0(act(temb))))) is applied by the convolution's own epilogue,
// so v never exists in HBM and the separate GroupNorm pass (one fp32 read + one fp16 write per element, the HBM-bound
// third of an evaluation) disappears for these layers.  GroupNorm statistics are per (image, group of cpg channels):
//   * pass 1 reads the accumulator from TMEM, forms v and reduces it to per-warp (32-row, group) partial sums;
//   * the eight epilogue warps meet on a named barrier and fold the partials of every image in the tile in a fixed order;
//     when an image spans several CTAs (gn_xc = 2 or 4: the CTAs of one cluster own the tiles of one image) the per-CTA
//     sums are exchanged through distributed shared memory (st.shared::cluster + mbarrier release / acquire at cluster
//     scope) and added in rank order -- deterministic, no atomics;
//   * pass 2 reads the accumulator AGAIN (it stays resident in its TMEM stage), normalises, applies swish and writes
//     fp16 rows through the usual swizzled staging buffer as full 64-byte row segments.
// Tiles may hold several whole images (8x8: two per 128-row tile, 4x4: eight; then a warp's 32 rows split into two
// 16-row segments).  Rows >= M (ragged last tile of a small batch) are excluded from the sums and not stored.
struct GnfCtx {
  float2* pstat;          // [MT][4][2][GNF_GMAX]
  float2* gstat;          // [GNF_IMGS][GNF_GMAX]
  float2* xchg;           // [2][4][GNF_GMAX]
  const float* gam;       // this tile's gamma / beta rows in shared memory
  const float* bet;
  uint64_t* xbar;         // two cluster exchange barriers (count 1 + transaction bytes), alternating by tile parity
  uint32_t xparity;       // which barrier / exchange buffer this tile uses (tile_seq & 1)
  uint32_t xphase;        // phase parity of that barrier ((tile_seq >> 1) & 1)
  uint32_t xrank;         // rank of this CTA in the cluster
  int warp;               // epilogue warp 0..7 (quad = warp & 3, group = warp >> 2)
  int tile_seq;           // how many tiles this CTA has processed before (timeline stamps)
};

__device__ __forceinline__ void st_cluster_f32x2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "GNF_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra GNF_DONE;\n\t"
      "bra GNF_WAIT_LOOP;\n\t"
      "GNF_DONE:\n\t"
      "}\n" ::"r"(ptx::smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// y * sigmoid(y) from y and t = -log2(e) * y (both formed by independent FMAs off the accumulator): ex2, add, rcp, mul --
// four issue slots per value, two of them on the special-function pipe.  (A variant sharing one reciprocal between four
// values -- 1.25 special-function results but 7.5 issue slots per value -- measured 2 - 3 % slower: the epilogue is bound
// by instruction issue, not by the special-function pipe.)  No clamp is
// needed: t -> +inf gives ex2 = inf, rcp = 0, y * 0 = 0.
__device__ __forceinline__ float gnf_silu_t(float y, float t) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return y * r;
}
// asynchronous remote store that completes `8` bytes on the destination CTA's mbarrier (no fence on the sender side:
// a release at cluster scope would first drain this thread's outstanding global stores, measured ~5 k cycles per tile)
__device__ __forceinline__ void st_async_f32x2(uint32_t cluster_addr, float a, float b, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(cluster_addr),
               "f"(a), "f"(b), "r"(cluster_bar)
               : "memory");
}
// Recursive-halving all-to-one reduction over the lanes of a warp (or of each 16-lane half when HALF): every lane enters
// with NV partial values; after log2(NV) exchange steps lane L holds the total of value gnf_holder_index(L) over all
// participating lanes -- NV shuffles in total instead of NV * log2(lanes).
template <int N>
__device__ __forceinline__ void gnf_halve(float (&a)[16], int mask, int lane) {
  const bool upper = (lane & mask) != 0;
#pragma unroll
  for (int k = 0; k < N / 2; ++k) {
    const float send = upper ? a[k] : a[k + N / 2];
    const float keep = upper ? a[k + N / 2] : a[k];
    a[k] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
  }
}
template <int NV, bool HALF>
__device__ __forceinline__ float gnf_reduce(float (&a)[16], int lane) {
  // halving steps use the highest lane bits; the remaining low bits are folded with plain butterfly adds
  if (NV == 4) {                                  // 16 channels per group: two groups per 32-column chunk
    if (HALF) { gnf_halve<4>(a, 8, lane); gnf_halve<2>(a, 4, lane); }
    else { gnf_halve<4>(a, 16, lane); gnf_halve<2>(a, 8, lane); a[0] += __shfl_xor_sync(0xffffffffu, a[0], 4); }
    a[0] += __shfl_xor_sync(0xffffffffu, a[0], 2);
    a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
    return a[0];
  }
  if (HALF) {
    if (NV == 16) { gnf_halve<16>(a, 8, lane); gnf_halve<8>(a, 4, lane); gnf_halve<4>(a, 2, lane); gnf_halve<2>(a, 1, lane); }
    else { gnf_halve<8>(a, 8, lane); gnf_halve<4>(a, 4, lane); gnf_halve<2>(a, 2, lane); a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1); }
  } else {
    if (NV == 16) { gnf_halve<16>(a, 16, lane); gnf_halve<8>(a, 8, lane); gnf_halve<4>(a, 4, lane); gnf_halve<2>(a, 2, lane);
                    a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1); }
    else { gnf_halve<8>(a, 16, lane); gnf_halve<4>(a, 8, lane); gnf_halve<2>(a, 4, lane);
           a[0] += __shfl_xor_sync(0xffffffffu, a[0], 2); a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1); }
  }
  return a[0];
}
// which value a lane holds after gnf_reduce, and whether it is the lane that publishes it
template <int NV, bool HALF>
__device__ __forceinline__ int gnf_holder_index(int lane, bool* writer) {
  if (NV == 4) {
    if (HALF) { *writer = (lane & 3) == 0; return (lane >> 2) & 3; }
    *writer = (lane & 7) == 0; return (lane >> 3) & 3;
  }
  if (HALF) {
    if (NV == 16) { *writer = true; return lane & 15; }
    *writer = (lane & 1) == 0; return (lane >> 1) & 7;
  }
  if (NV == 16) { *writer = (lane & 1) == 0; return (lane >> 1) & 15; }
  *writer = (lane & 3) == 0; return (lane >> 2) & 7;
}

// pass 1 of one 32 x 32 chunk held as thread = row: v = acc * scale + bias, per-(group, stat) partials (gnf_chunk_partials),
// then the lane reduction (gnf_chunk_reduce) -- split in two so that the TMEM load of the next chunk can be issued in between
// (the bias row in shared memory is already multiplied by `scale`; FULL: every row of the tile exists)
template <int NV, bool FULL>
__device__ __forceinline__ void gnf_chunk_partials(const uint32_t (&r)[32], uint32_t bias_base, float scale, bool row_ok,
                                                   float (&a)[16]) {
  constexpr int CPG = 64 / NV;                        // NV = 2 * groups per 32-column chunk
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = 0.f;
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    float4 b = lds128(bias_base + j4 * 16);
    const float v0 = (FULL || row_ok) ? fmaf(__uint_as_float(r[4 * j4 + 0]), scale, b.x) : 0.f;
    const float v1 = (FULL || row_ok) ? fmaf(__uint_as_float(r[4 * j4 + 1]), scale, b.y) : 0.f;
    const float v2 = (FULL || row_ok) ? fmaf(__uint_as_float(r[4 * j4 + 2]), scale, b.z) : 0.f;
    const float v3 = (FULL || row_ok) ? fmaf(__uint_as_float(r[4 * j4 + 3]), scale, b.w) : 0.f;
    const int g = (4 * j4) / CPG;                     // compile-time after unrolling
    a[2 * g] += (v0 + v1) + (v2 + v3);
    a[2 * g + 1] = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, a[2 * g + 1]))));
  }
}
template <int NV, bool HALF>
__device__ __forceinline__ void gnf_chunk_reduce(float (&a)[16], int lane, float2* dst_seg0, int gi0) {
  const float tot = gnf_reduce<NV, HALF>(a, lane);
  bool writer;
  const int idx = gnf_holder_index<NV, HALF>(lane, &writer);
  if (writer) {
    float* d = reinterpret_cast<float*>(dst_seg0 + (HALF && lane >= 16 ? GNF_GMAX : 0) + gi0 + (idx >> 1)) + (idx & 1);
    *d = tot;
  }
}

// fold of the per-warp partials into per-(image, group) mean / rstd (with the cluster exchange when an image spans
// several CTAs), bracketed by the two named barriers of the eight epilogue warps
template <int BLOCK_N, int MT>
__device__ __forceinline__ void gnf_fold(const GemmArgs& p, const GnfCtx& gx, int lane) {
  constexpr int TR = MT * BLOCK_M;
  const int cpg = p.gn_cpg, rpi = p.gn_rpi;
  const bool split = rpi < 32;
  const int G = BLOCK_N / cpg;
  const bool stamp = gx.warp == 0 && lane == 0;
  (void)stamp;
  asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
  GDDIM_STAMP(p, stamp, gx.tile_seq, 3);
  // ---------------- fold: (image, group) -> mean, rstd ----------------
  const int tid = gx.warp * 32 + lane;
  const float inv_n = 1.0f / ((float)rpi * (float)cpg);
  if (p.gn_xc > 1 && p.gn_xg != nullptr) {
    // one image spans gn_xc consecutive CTAs of the grid that are NOT one cluster (32x32 images on cta_group::2 pairs: two
    // pairs per image, all 148 SMs busy -- clusters of four only fit 132).  Exchange through global memory without any
    // fence: every (sum, sumsq) travels in two 64-bit words {value, sequence tag} written with relaxed gpu-scope stores
    // (64-bit accesses are single-copy atomic), the readers poll until the tag of this round shows up.  Two buffers
    // alternate by round; nobody can be two rounds ahead (round r + 1 needs every peer's data of round r + 1, which a
    // peer only writes after it has read round r).  All CTAs of the persistent grid are co-resident (one per SM).
    if (tid < G) {
      float S = 0.f, Q = 0.f;
      for (int sl = 0; sl < MT * 4; ++sl) {
        const float2 a = gx.pstat[(sl * 2) * GNF_GMAX + tid];
        S += a.x; Q += a.y;
      }
      const unsigned int seq = (unsigned int)gx.tile_seq + 1u;
      const int xc = p.gn_xc;
      const int grp = blockIdx.x / xc, me = blockIdx.x % xc;
      unsigned long long* base = p.gn_xg + ((size_t)(grp * 2 + (int)gx.xparity) * 4) * (GNF_GMAX * 2);
      unsigned long long* mine = base + (size_t)me * (GNF_GMAX * 2) + tid * 2;
      const unsigned long long w0 = ((unsigned long long)seq << 32) | __float_as_uint(S);
      const unsigned long long w1 = ((unsigned long long)seq << 32) | __float_as_uint(Q);
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mine), "l"(w0) : "memory");
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mine + 1), "l"(w1) : "memory");
      float St = 0.f, Qt = 0.f;
      for (int rk = 0; rk < xc; ++rk) {
        float s_r = S, q_r = Q;
        if (rk != me) {
          const unsigned long long* src = base + (size_t)rk * (GNF_GMAX * 2) + tid * 2;
          unsigned long long a0, a1;
          unsigned int spins = 0;
          do {
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a0) : "l"(src) : "memory");
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a1) : "l"(src + 1) : "memory");
            if (++spins > (1u << 26)) __trap();          // a peer CTA never showed up: fail loudly instead of hanging
          } while ((unsigned int)(a0 >> 32) != seq || (unsigned int)(a1 >> 32) != seq);
          s_r = __uint_as_float((unsigned int)a0);
          q_r = __uint_as_float((unsigned int)a1);
        }
        St += s_r; Qt += q_r;
      }
      const float mean = St * inv_n;
      const float var = fmaxf(Qt * inv_n - mean * mean, 0.f);
      gx.gstat[tid] = make_float2(mean, rsqrtf(var + p.gn_eps));
    }
  } else if (p.gn_xc > 1) {
    // one image spans the gn_xc CTAs of this cluster: every CTA sends its G (sum, sumsq) pairs to all CTAs of the cluster
    // with asynchronous remote stores that complete bytes on the receiver's mbarrier (armed with the expected byte count);
    // two barriers / buffers alternate by tile parity, so data of tile i + 1 can never be counted into the phase of tile i
    uint64_t* xb = gx.xbar + gx.xparity;
    if (tid == 0) ptx::mbar_arrive_expect_tx(xb, (uint32_t)(p.gn_xc * G * 8));
    if (tid < G) {
      float S = 0.f, Q = 0.f;
      for (int sl = 0; sl < MT * 4; ++sl) {
        const float2 a = gx.pstat[(sl * 2) * GNF_GMAX + tid];
        S += a.x; Q += a.y;
      }
      float2* mine = gx.xchg + (gx.xparity * 4 + gx.xrank) * GNF_GMAX + tid;
      const uint32_t my_addr = ptx::smem_u32(mine), bar_addr = ptx::smem_u32(xb);
      for (int rk = 0; rk < p.gn_xc; ++rk) st_async_f32x2(ptx::mapa(my_addr, rk), S, Q, ptx::mapa(bar_addr, rk));
      ptx::mbar_wait(xb, gx.xphase);
      S = 0.f; Q = 0.f;
      for (int rk = 0; rk < p.gn_xc; ++rk) {
        const float2 a = gx.xchg[(gx.xparity * 4 + rk) * GNF_GMAX + tid];
        S += a.x; Q += a.y;
      }
      const float mean = S * inv_n;
      const float var = fmaxf(Q * inv_n - mean * mean, 0.f);
      gx.gstat[tid] = make_float2(mean, rsqrtf(var + p.gn_eps));
    }
  } else {
    const int n_img = rpi >= TR ? 1 : TR / rpi;
    for (int t = tid; t < n_img * G; t += EPI_WARPS * 32) {
      const int img = t / G, g = t - img * G;
      float S = 0.f, Q = 0.f;
      if (split) {
        const float2 a = gx.pstat[((img >> 1) * 2 + (img & 1)) * GNF_GMAX + g];
        S = a.x; Q = a.y;
      } else {
        const int per = rpi >= TR ? MT * 4 : rpi / 32;      // 32-row slabs per image
        for (int k = 0; k < per; ++k) {
          const float2 a = gx.pstat[((img * per + k) * 2) * GNF_GMAX + g];
          S += a.x; Q += a.y;
        }
      }
      const float mean = S * inv_n;
      const float var = fmaxf(Q * inv_n - mean * mean, 0.f);
      gx.gstat[img * GNF_GMAX + g] = make_float2(mean, rsqrtf(var + p.gn_eps));
    }
  }
  GDDIM_STAMP(p, stamp, gx.tile_seq, 4);
  asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
  GDDIM_STAMP(p, stamp, gx.tile_seq, 5);
}

template <int BLOCK_N, int MT, bool FULL>
__device__ __forceinline__ void epi_tile_gnf(const EpiCtx<BLOCK_N, MT>& cx, const GnfCtx& gx) {
  constexpr int RS = 32;
  constexpr int NCH = BLOCK_N / 32;
  constexpr int NQ = MT * NCH;
  constexpr int TR = MT * BLOCK_M;
  const GemmArgs& p = cx.p;
  const int lane = cx.lane;
  const int rsub = lane >> 3;
  const int c4 = (lane & 7) * 4;
  const int quad = gx.warp & 3;
  const uint32_t stg_w = ptx::smem_u32(cx.stg) + lane * RS * 4;
  const uint32_t wx = lane & 7;
  const uint32_t stg_r0 = ptx::smem_u32(cx.stg) + rsub * RS * 4 + (((lane & 7) ^ rsub) << 4);
  const uint32_t stg_r1 = ptx::smem_u32(cx.stg) + (rsub + 4) * RS * 4 + (((lane & 7) ^ (rsub + 4)) << 4);
  const uint32_t bias_a = ptx::smem_u32(cx.bias_s) + c4 * 4;
  const float scale = p.scale;
  const int cpg = p.gn_cpg;
  const int rpi = p.gn_rpi;
  const bool split = rpi < 32;                       // 4x4 images: rows 0-15 and 16-31 of a warp are different images
  const long long mwarp = cx.m0;                     // first row of this warp in sub-tile 0 (tile row0 + quad * 32)
  const long long mtile = cx.m0 - quad * 32;

  [[maybe_unused]] const bool stamp = gx.warp == 0 && lane == 0;
  GDDIM_STAMP(p, stamp, gx.tile_seq, 0);
  ptx::mbar_wait(cx.tfull, cx.tfull_phase);
  ptx::tc_fence_after();
  GDDIM_STAMP(p, stamp, gx.tile_seq, 1);
  uint32_t r[32];
  // ---------------- pass 1: per-warp (segment, group) sums of v and v^2 ----------------
  // thread = accumulator row: no staging through shared memory, the lanes are reduced with recursive-halving shuffles;
  // the TMEM load of chunk q + 1 is in flight while the partials of chunk q are reduced
  if (!GDDIM_DBG_IS(p, 6)) {
    const long long mrow = mwarp + lane;                         // this thread's row in sub-tile 0
    const uint32_t bias_row = ptx::smem_u32(cx.bias_s);
    if (cx.group < NQ) ptx::tmem_ld_32x32b_x32(cx.taddr + (cx.group / NCH) * BLOCK_N + (cx.group % NCH) * 32, r);
#pragma unroll 1
    for (int q = cx.group; q < NQ; q += EPI_GROUPS) {
      const int mi = q / NCH, c0 = (q % NCH) * 32;
      const bool row_ok = mrow + (long long)mi * BLOCK_M < p.M;
      float2* dst = gx.pstat + ((mi * 4 + quad) * 2) * GNF_GMAX;
      const int gi0 = c0 / cpg;
      float a[16];
      ptx::tmem_ld_wait();
      if (cpg == 4) gnf_chunk_partials<16, FULL>(r, bias_row + c0 * 4, scale, row_ok, a);
      else if (cpg == 8) gnf_chunk_partials<8, FULL>(r, bias_row + c0 * 4, scale, row_ok, a);
      else gnf_chunk_partials<4, FULL>(r, bias_row + c0 * 4, scale, row_ok, a);
      const int qn = q + EPI_GROUPS;
      if (qn < NQ) ptx::tmem_ld_32x32b_x32(cx.taddr + (qn / NCH) * BLOCK_N + (qn % NCH) * 32, r);
      if (cpg == 4) {
        if (split) gnf_chunk_reduce<16, true>(a, lane, dst, gi0); else gnf_chunk_reduce<16, false>(a, lane, dst, gi0);
      } else if (cpg == 8) {
        if (split) gnf_chunk_reduce<8, true>(a, lane, dst, gi0); else gnf_chunk_reduce<8, false>(a, lane, dst, gi0);
      } else {
        if (split) gnf_chunk_reduce<4, true>(a, lane, dst, gi0); else gnf_chunk_reduce<4, false>(a, lane, dst, gi0);
      }
    }
  }
  GDDIM_STAMP(p, stamp, gx.tile_seq, 2);
  gnf_fold<BLOCK_N, MT>(p, gx, lane);
  // ---------------- pass 2: normalise, activate, store fp16 ----------------
  const long long ldo = p.ldo;
  if (GDDIM_DBG_IS(p, 8)) return;
  if (cx.group < NQ && !GDDIM_DBG_IS(p, 4)) ptx::tmem_ld_32x32b_x32(cx.taddr + (cx.group / NCH) * BLOCK_N + (cx.group % NCH) * 32, r);
#pragma unroll 1
  for (int q = cx.group; q < NQ; q += EPI_GROUPS) {
    const int mi = q / NCH, c0 = (q % NCH) * 32;
    const float4 g4 = *reinterpret_cast<const float4*>(gx.gam + c0 + c4);
    const float4 b4 = *reinterpret_cast<const float4*>(gx.bet + c0 + c4);
    const int gi = (c0 + c4) / cpg;
    // image of this warp's rows inside the tile: row offset (mi * 128 + quad * 32 [+ 16]) / rpi
    const int row_off = mi * BLOCK_M + quad * 32;
    const int img0 = (p.gn_xc > 1 || rpi >= TR) ? 0 : row_off / rpi;
    const float2 st0 = gx.gstat[img0 * GNF_GMAX + gi];
    const float2 st1 = split ? gx.gstat[(img0 + 1) * GNF_GMAX + gi] : st0;
    // y = (acc * scale + bias * scale) * A + O = acc * (scale * A) + (bias_s * A + O) with A = rstd * gamma, O = beta - mean * A
    // (bias_s holds bias * scale); t = -log2(e) * y comes from a second FMA off the accumulator
    constexpr float NL2E = -1.4426950408889634f;
    float4 bsum = lds128(bias_a + c0 * 4);
    float4 a0, o0, a1, o1;
    a0.x = st0.y * g4.x; a0.y = st0.y * g4.y; a0.z = st0.y * g4.z; a0.w = st0.y * g4.w;
    o0.x = fmaf(bsum.x - st0.x, a0.x, b4.x); o0.y = fmaf(bsum.y - st0.x, a0.y, b4.y);
    o0.z = fmaf(bsum.z - st0.x, a0.z, b4.z); o0.w = fmaf(bsum.w - st0.x, a0.w, b4.w);
    a0.x *= scale; a0.y *= scale; a0.z *= scale; a0.w *= scale;
    a1 = a0; o1 = o0;
    if (split) {
      a1.x = st1.y * g4.x; a1.y = st1.y * g4.y; a1.z = st1.y * g4.z; a1.w = st1.y * g4.w;
      o1.x = fmaf(bsum.x - st1.x, a1.x, b4.x); o1.y = fmaf(bsum.y - st1.x, a1.y, b4.y);
      o1.z = fmaf(bsum.z - st1.x, a1.z, b4.z); o1.w = fmaf(bsum.w - st1.x, a1.w, b4.w);
      a1.x *= scale; a1.y *= scale; a1.z *= scale; a1.w *= scale;
    }
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) sts128(stg_w + ((j ^ wx) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    // the registers are free again: the TMEM load of the next chunk overlaps the normalise / swish / store work below
    {
      const int qn = q + EPI_GROUPS;
      if (qn < NQ && !GDDIM_DBG_IS(p, 4)) ptx::tmem_ld_32x32b_x32(cx.taddr + (qn / NCH) * BLOCK_N + (qn % NCH) * 32, r);
    }
    __syncwarp();
    const long long mb = mwarp + (long long)mi * BLOCK_M + rsub;
    const int n0 = cx.n_tile0 + c0 + c4;
    __half* o16 = p.out16 + mb * ldo + n0;
    const bool silu = p.gn_silu != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 acc = lds128(((i & 1) ? stg_r1 : stg_r0) + (i >> 1) * 8 * RS * 4);
      const float4 aa = (i >= 4) ? a1 : a0, oo = (i >= 4) ? o1 : o0;
      float4 v;
      v.x = fmaf(acc.x, aa.x, oo.x); v.y = fmaf(acc.y, aa.y, oo.y); v.z = fmaf(acc.z, aa.z, oo.z); v.w = fmaf(acc.w, aa.w, oo.w);
      if (silu && !GDDIM_DBG_IS(p, 5)) {
        v.x = gnf_silu_t(v.x, fmaf(acc.x, aa.x * NL2E, oo.x * NL2E)); v.y = gnf_silu_t(v.y, fmaf(acc.y, aa.y * NL2E, oo.y * NL2E));
        v.z = gnf_silu_t(v.z, fmaf(acc.z, aa.z * NL2E, oo.z * NL2E)); v.w = gnf_silu_t(v.w, fmaf(acc.w, aa.w * NL2E, oo.w * NL2E));
      }
      if ((FULL || mb + i * 4 < p.M) && !GDDIM_DBG_IS(p, 7)) {
        __half2 h0 = __floats2half2_rn(v.x, v.y);
        __half2 h1 = __floats2half2_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(o16 + (long long)(i * 4) * ldo) = pk;
      }
    }
    __syncwarp();
  }
  GDDIM_STAMP(p, stamp, gx.tile_seq, 6);
  (void)mtile;
}

// Dual GroupNorm epilogue: the convolution ALSO keeps its plain result.  out32 (+ colstats) = the linear epilogue value v
// (bias, residual, scale: the (x + h)/sqrt(2) trunk of layerspp.py:224-227) and out16 = act(GroupNorm(v)) with the
// parameters of the NEXT block's GroupNorm_0 (layerspp.py:196) -- the consumer's normalisation pass disappears.
// Pass 1 is the ordinary linear epilogue (epi_tile with the GNP hook); the accumulator is not needed afterwards, pass 2
// re-reads v from out32: every lane reads back exactly the elements it stored itself (program order, L2 hits).
template <int BLOCK_N, int MT, bool FULL>
__device__ __forceinline__ void epi_tile_gnf_dual_pass2(const EpiCtx<BLOCK_N, MT>& cx, const GnfCtx& gx) {
  constexpr int NCH = BLOCK_N / 32;
  constexpr int NQ = MT * NCH;
  constexpr int TR = MT * BLOCK_M;
  constexpr float NL2E = -1.4426950408889634f;
  const GemmArgs& p = cx.p;
  const int lane = cx.lane;
  const int rsub = lane >> 3;
  const int c4 = (lane & 7) * 4;
  const int quad = gx.warp & 3;
  const int cpg = p.gn_cpg, rpi = p.gn_rpi;
  const bool split = rpi < 32;
  const bool silu = p.gn_silu != 0;
  const long long ldo = p.ldo;
#pragma unroll 1
  for (int q = cx.group; q < NQ; q += EPI_GROUPS) {
    const int mi = q / NCH, c0 = (q % NCH) * 32;
    const long long mb = cx.m0 + (long long)mi * BLOCK_M + rsub;
    const int n0 = cx.n_tile0 + c0 + c4;
    const float* src = p.out32 + mb * ldo + n0;
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v[i] = (FULL || mb + i * 4 < p.M) ? __ldcg(reinterpret_cast<const float4*>(src + (long long)(i * 4) * ldo)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 g4 = *reinterpret_cast<const float4*>(gx.gam + c0 + c4);
    const float4 b4 = *reinterpret_cast<const float4*>(gx.bet + c0 + c4);
    const int gi = (c0 + c4) / cpg;
    const int row_off = mi * BLOCK_M + quad * 32;
    const int img0 = (p.gn_xc > 1 || rpi >= TR) ? 0 : row_off / rpi;
    const float2 st0 = gx.gstat[img0 * GNF_GMAX + gi];
    float4 a0, o0, a1, o1;
    a0.x = st0.y * g4.x; a0.y = st0.y * g4.y; a0.z = st0.y * g4.z; a0.w = st0.y * g4.w;
    o0.x = b4.x - st0.x * a0.x; o0.y = b4.y - st0.x * a0.y; o0.z = b4.z - st0.x * a0.z; o0.w = b4.w - st0.x * a0.w;
    a1 = a0; o1 = o0;
    if (split) {
      const float2 st1 = gx.gstat[(img0 + 1) * GNF_GMAX + gi];
      a1.x = st1.y * g4.x; a1.y = st1.y * g4.y; a1.z = st1.y * g4.z; a1.w = st1.y * g4.w;
      o1.x = b4.x - st1.x * a1.x; o1.y = b4.y - st1.x * a1.y; o1.z = b4.z - st1.x * a1.z; o1.w = b4.w - st1.x * a1.w;
    }
    __half* o16 = p.out16 + mb * ldo + n0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 aa = (i >= 4) ? a1 : a0, oo = (i >= 4) ? o1 : o0;
      float4 y;
      y.x = fmaf(v[i].x, aa.x, oo.x); y.y = fmaf(v[i].y, aa.y, oo.y); y.z = fmaf(v[i].z, aa.z, oo.z); y.w = fmaf(v[i].w, aa.w, oo.w);
      if (silu) {
        y.x = gnf_silu_t(y.x, fmaf(v[i].x, aa.x * NL2E, oo.x * NL2E)); y.y = gnf_silu_t(y.y, fmaf(v[i].y, aa.y * NL2E, oo.y * NL2E));
        y.z = gnf_silu_t(y.z, fmaf(v[i].z, aa.z * NL2E, oo.z * NL2E)); y.w = gnf_silu_t(y.w, fmaf(v[i].w, aa.w * NL2E, oo.w * NL2E));
      }
      if (FULL || mb + i * 4 < p.M) {
        __half2 h0 = __floats2half2_rn(y.x, y.y);
        __half2 h1 = __floats2half2_rn(y.z, y.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(o16 + (long long)(i * 4) * ldo) = pk;
      }
    }
  }
}

}  // namespace gddim
