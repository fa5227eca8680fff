// K2: fused single-head self-attention for the 16x16 level (T = 256 tokens, C = 256 channels) on tcgen05.
//
// Replaces, per attention block, the chain  QK^T GEMM (+ row softmax epilogue, P to HBM) -> V transpose -> P.V GEMM
// that stands in for the two einsums of cld_jax/models/layerspp.py:74-78:
//     w = softmax_(hw)( einsum('bhwc,bHWc->bhwHW', q, k) * C^-0.5 ),   h = einsum('bhwHW,bHWc->bhwc', w, v)
// One CTA per (image, 128-query half), 9 warps:
//   warps 0-7  softmax + epilogue: thread = (query row = TMEM lane, half of the 256 columns; warps w and w + 4 share a
//              lane quadrant).  S is read from TMEM twice (row max, then exp -> fp16 P written into shared memory in
//              the K-major 128B-swizzle layout of a UMMA A operand); the two column halves exchange row max / row sum
//              through shared memory.  O is read from TMEM, scaled by 1/rowsum and transposed through swizzled shared
//              memory so that every global store covers full 128-byte lines
//   warp 8     one thread: TMA loads (Q, K per 64-channel block, each with its own barrier so that the first MMAs
//              start while the rest is still in flight; later V into the shared memory K occupied) and both MMA
//              sequences
//              S = Q K^T (K-major B) and O = P V, where V is consumed as it lies in the qkv tensor, [key][channel]:
//              an MN-major B operand (instruction-descriptor bit 16), so the V^T pass of the unfused chain disappears
// TMEM: S in columns [0,256), O in [256,512).  Scores and probabilities never leave the SM.
// PROJ variant (the network path): the block's output projection NIN_3 and the residual (layerspp.py:79-83) run in
// the same CTA -- O / rowsum is written as the fp16 A operand into the shared memory P occupied, the 256 x 256
// projection weights stream into the shared memory V occupied, Y = O W3^T accumulates in TMEM columns [0,256) and
// leaves through the GEMM kernels' own linear epilogue (gemm_epilogue.cuh: bias, residual, 1/sqrt(2), fp32 output,
// column statistics for the next GroupNorm) -- bit-identical to the separate projection GEMM it replaces.
// Numerics are those of the unfused chain: P = exp2(s*scale*log2e - max) rounded to fp16, row sum taken over the
// ROUNDED values, O = (P V) / sum in fp32, fp16 output.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"
#include "gemm_epilogue.cuh"

namespace gddim {

namespace {
constexpr int AT_T = 256;            // tokens (keys) per image
constexpr int AT_C = 256;            // channels
constexpr int AT_Q = 128;            // query rows per CTA
constexpr int AT_THREADS = 288;
constexpr int AT_CTRL = 256;         // the TMA + MMA thread
constexpr int AT_QBLK = AT_Q * 128;  // one 64-channel K-block of Q (or 64-key block of P): 128 rows x 128 B
constexpr int AT_KBLK = AT_T * 128;  // one 64-channel block of K (K-major) or of V (MN-major): 256 rows x 128 B
constexpr int AT_OFF_K = 4 * AT_QBLK;
constexpr int AT_OFF_BAR = AT_OFF_K + 4 * AT_KBLK;
constexpr int AT_OFF_XCH = AT_OFF_BAR + 128;        // [2 halves][128 rows] row max, then the same for row sums
constexpr int AT_OFF_BIAS = AT_OFF_XCH + 2048;      // [256] projection bias row (PROJ)
constexpr int AT_OFF_GNF = AT_OFF_BIAS + 1024;      // PROJ + GroupNorm of the output: gemm_epilogue.cuh GnfSmem<256>
constexpr int AT_SMEM = AT_OFF_GNF + GnfSmem<256>::BYTES + 1024;  // + alignment slack

// MN-major B operand (V: rows = keys, 128 B = 64 channels per row, 128B swizzle): 8-key groups 1024 B apart (SBO),
// 64-channel blocks AT_KBLK apart (LBO)
__device__ __forceinline__ uint64_t desc_v_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(AT_KBLK >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void tma_load_3d(const void* tmap, uint64_t* bar, void* smem, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(ptx::smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// eight fp32 accumulator values r[0..7] * s -> eight fp16 -> one 16-byte shared-memory store
__device__ __forceinline__ void sts_scaled8(uint32_t addr, const uint32_t* r, float s) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
               "r"(pack_h2(__uint_as_float(r[0]) * s, __uint_as_float(r[1]) * s)),
               "r"(pack_h2(__uint_as_float(r[2]) * s, __uint_as_float(r[3]) * s)),
               "r"(pack_h2(__uint_as_float(r[4]) * s, __uint_as_float(r[5]) * s)),
               "r"(pack_h2(__uint_as_float(r[6]) * s, __uint_as_float(r[7]) * s))
               : "memory");
}
}  // namespace

struct AttnArgs {
  __half* out16;     // [B, T, C]
  float sc;          // C^-0.5 * log2(e)
  int reverse;
  const float* bias3;   // PROJ: projection bias [C]
  GemmArgs g;           // PROJ: residual / out32 / colstats / ldo / scale / M of the projection epilogue
};

template <bool PROJ>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn256_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_w3, const AttnArgs p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                  // Q (4 x [128 x 64]), later P (4 x [128 x 64 keys]), later the output staging
  uint8_t* sK = smem + AT_OFF_K;       // K (4 x [256 x 64]), later V (4 x [256 keys x 64 channels])
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_OFF_BAR);
  uint64_t* bar_qk = bars;             // [4] Q and K of one 64-channel block landed
  uint64_t* bar_s = bars + 4;          // S = Q K^T complete (also: Q / K shared memory free)
  uint64_t* bar_v = bars + 5;          // V landed
  uint64_t* bar_p = bars + 6;          // P written by the 256 softmax threads
  uint64_t* bar_o = bars + 7;          // O = P V complete
  uint64_t* bar_w = bars + 9;          // PROJ: projection weights landed
  uint64_t* bar_o16 = bars + 10;       // PROJ: O / rowsum written as an fp16 A operand by the 256 epilogue threads
  uint64_t* bar_y = bars + 11;         // PROJ: Y = O W3^T complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* bias_s = reinterpret_cast<float*>(smem + AT_OFF_BIAS);
  float* xch_max = reinterpret_cast<float*>(smem + AT_OFF_XCH);
  float* xch_sum = xch_max + 2 * AT_Q;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = p.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int b = unit >> 1, half = unit & 1;

  if (threadIdx.x == AT_CTRL) {
    ptx::prefetch_tmap(&tm_qkv);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&bar_qk[i], 1);
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_v, 1);
    ptx::mbar_init(bar_p, 256);
    ptx::mbar_init(bar_o, 1);
    if (PROJ) { ptx::prefetch_tmap(&tm_w3); ptx::mbar_init(bar_w, 1); ptx::mbar_init(bar_o16, 256); ptx::mbar_init(bar_y, 1); }
    if (PROJ && p.g.gn_gamma != nullptr) {       // statistics exchange with the CTA that owns the image's other half
      ptx::mbar_init(reinterpret_cast<uint64_t*>(smem + AT_OFF_GNF + GnfSmem<256>::OFF_BAR), 1);
      ptx::mbar_init(reinterpret_cast<uint64_t*>(smem + AT_OFF_GNF + GnfSmem<256>::OFF_BAR) + 1, 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 8) { __syncwarp(); ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  pdl_wait();
  ptx::tc_fence_before();
  // (GroupNorm epilogue: launched as clusters of two CTAs = one image; the peer's barrier must exist before data is sent)
  if (PROJ && p.g.gn_gamma != nullptr) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == AT_CTRL) {
    // ---- Q (this CTA's 128 rows) and K (all 256 keys), one 64-channel block per barrier ----
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      ptx::mbar_arrive_expect_tx(&bar_qk[kb], AT_QBLK + AT_KBLK);
      tma_load_3d(&tm_qkv, &bar_qk[kb], sQ + kb * AT_QBLK, kb * 64, half * AT_Q, b);
      tma_load_3d(&tm_qkv, &bar_qk[kb], sK + kb * AT_KBLK, AT_C + kb * 64, 0, b);
      tma_load_3d(&tm_qkv, &bar_qk[kb], sK + kb * AT_KBLK + AT_QBLK, AT_C + kb * 64, 128, b);
    }
    // ---- S[128, 256] = Q K^T ----
    {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(128, 256);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        ptx::mbar_wait(&bar_qk[kb], 0);
        ptx::tc_fence_after();
        const uint64_t a_desc = ptx::umma_desc_sw128(ptx::smem_u32(sQ + kb * AT_QBLK));
        const uint64_t b_desc = ptx::umma_desc_sw128(ptx::smem_u32(sK + kb * AT_KBLK));
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
      }
      ptx::umma_commit(bar_s);
    }
    // ---- V into the shared memory K occupied (free once S is complete) ----
    ptx::mbar_wait(bar_s, 0);
    ptx::mbar_arrive_expect_tx(bar_v, 4 * AT_KBLK);
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      tma_load_3d(&tm_qkv, bar_v, sK + nb * AT_KBLK, 2 * AT_C + nb * 64, 0, b);
      tma_load_3d(&tm_qkv, bar_v, sK + nb * AT_KBLK + AT_QBLK, 2 * AT_C + nb * 64, 128, b);
    }
    ptx::mbar_wait(bar_v, 0);
    ptx::mbar_wait(bar_p, 0);
    ptx::tc_fence_after();
    // ---- O[128, 256] = P V : A = P (K-major, 64-key blocks), B = V (MN-major), K = 256 keys in 16 steps ----
    {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(128, 256) | (1u << 16);
#pragma unroll
      for (int s = 0; s < 16; ++s) {
        const uint64_t a_desc = ptx::umma_desc_sw128(ptx::smem_u32(sQ + (s >> 2) * AT_QBLK)) + 2 * (s & 3);
        const uint64_t b_desc = desc_v_mn(ptx::smem_u32(sK + s * 16 * 128));
        ptx::umma_f16(tmem_base + 256, a_desc, b_desc, idesc, s != 0);
      }
      ptx::umma_commit(bar_o);
    }
    if (PROJ) {
      // ---- projection weights [256 out, 256 in] (K-major, 64-channel blocks) into the shared memory V occupied ----
      ptx::mbar_wait(bar_o, 0);
      ptx::mbar_arrive_expect_tx(bar_w, 4 * AT_KBLK);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        ptx::tma_load_2d(&tm_w3, bar_w, sK + kb * AT_KBLK, kb * 64, 0);
        ptx::tma_load_2d(&tm_w3, bar_w, sK + kb * AT_KBLK + AT_QBLK, kb * 64, 128);
      }
      ptx::mbar_wait(bar_w, 0);
      ptx::mbar_wait(bar_o16, 0);
      ptx::tc_fence_after();
      // ---- Y[128, 256] = O W3^T into the TMEM columns S occupied ----
      constexpr uint32_t idesc = ptx::umma_idesc_f16(128, 256);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const uint64_t a_desc = ptx::umma_desc_sw128(ptx::smem_u32(sQ + kb * AT_QBLK));
        const uint64_t b_desc = ptx::umma_desc_sw128(ptx::smem_u32(sK + kb * AT_KBLK));
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
      }
      ptx::umma_commit(bar_y);
    }
  } else if (warp < 8) {
    const int quad = warp & 3, ch = warp >> 2;                // TMEM lane quadrant, column half
    const int row = quad * 32 + lane;                         // query row inside the CTA tile = TMEM lane
    const uint32_t t_s = tmem_base + (uint32_t(quad * 32) << 16) + ch * 128;
    uint32_t r[32];
    ptx::mbar_wait(bar_s, 0);
    ptx::tc_fence_after();
    float mx = -INFINITY;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      ptx::tmem_ld_32x32b_x32(t_s + c0, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]) * p.sc);
    }
    xch_max[ch * AT_Q + row] = mx;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    mx = fmaxf(mx, xch_max[(ch ^ 1) * AT_Q + row]);
    float sum = 0.f;
    const uint32_t p_row = ptx::smem_u32(sQ) + row * 128;
    const uint32_t sw = row & 7;
#pragma unroll 1
    for (int c0 = ch * 128; c0 < ch * 128 + 128; c0 += 32) {
      ptx::tmem_ld_32x32b_x32(tmem_base + (uint32_t(quad * 32) << 16) + c0, r);
      ptx::tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        // round to fp16 first so that the row sum matches what the P.V MMA actually consumes
        v[j] = __half2float(__float2half_rn(exp2f(__uint_as_float(r[j]) * p.sc - mx)));
        sum += v[j];
      }
      // keys c0 .. c0+31 -> 64-key block c0 / 64, 16-byte units (c0 % 64) / 8 .. + 3 of this row (128B swizzle)
      const uint32_t blk = p_row + (c0 >> 6) * AT_QBLK;
      const uint32_t u0 = (c0 & 63) >> 3;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t a = blk + (((u0 + u) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_h2(v[8 * u], v[8 * u + 1])),
                     "r"(pack_h2(v[8 * u + 2], v[8 * u + 3])), "r"(pack_h2(v[8 * u + 4], v[8 * u + 5])),
                     "r"(pack_h2(v[8 * u + 6], v[8 * u + 7]))
                     : "memory");
      }
    }
    xch_sum[ch * AT_Q + row] = sum;
    ptx::fence_proxy_async();          // generic-proxy writes of P -> visible to the tensor core (async proxy)
    ptx::mbar_arrive(bar_p);

    // ---- epilogue ----
    ptx::mbar_wait(bar_o, 0);
    ptx::tc_fence_after();
    asm volatile("bar.sync 1, 256;" ::: "memory");            // the other half's row sums are written
    // fixed order (half 0 + half 1) in both threads of a row: identical scale factors
    const float inv = 1.0f / (xch_sum[row] + xch_sum[AT_Q + row]);
    const uint32_t t_o = tmem_base + (uint32_t(quad * 32) << 16) + 256 + ch * 128;
    if (PROJ) {
      // O / rowsum -> fp16 A operand of the projection, in the layout P had (64-channel blocks, 128B swizzle)
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        ptx::tmem_ld_32x32b_x32(t_o + c0, r);
        ptx::tmem_ld_wait();
        const int cc = ch * 128 + c0;
        const uint32_t blk = p_row + (cc >> 6) * AT_QBLK;
        const uint32_t u0 = (cc & 63) >> 3;
#pragma unroll
        for (int u = 0; u < 4; ++u) sts_scaled8(blk + (((u0 + u) ^ sw) << 4), r + 8 * u, inv);
      }
      ptx::fence_proxy_async();
      ptx::mbar_arrive(bar_o16);
      bias_s[threadIdx.x] = __ldg(p.bias3 + threadIdx.x);                  // 256 epilogue threads = 256 output columns
      const bool gn = p.g.gn_gamma != nullptr;
      uint8_t* gnf = smem + AT_OFF_GNF;
      float* gb = reinterpret_cast<float*>(gnf + GnfSmem<256>::OFF_GB);
      if (gn) { gb[threadIdx.x] = __ldg(p.g.gn_gamma + threadIdx.x); gb[256 + threadIdx.x] = __ldg(p.g.gn_beta + threadIdx.x); }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // the GEMM kernels' linear epilogue on this warp's lane quadrant and alternate 32-column chunks; its staging
      // (4 KB per warp) reuses the A-operand memory, which it touches only after Y is complete
      float* stg_f = reinterpret_cast<float*>(sQ) + warp * 32 * SmemLayout<256, 1>::EPI_ROW_FLOATS;
      EpiCtx<256, 1> cx{p.g, stg_f, bias_s, bar_y, 0u, tmem_base + (uint32_t(quad * 32) << 16),
                        (long long)b * AT_T + half * AT_Q + quad * 32, 0, lane, ch};
      if (gn) {
        // the NEXT block's GroupNorm_0 + swish applied here (dual GroupNorm epilogue, gemm_epilogue.cuh): pass 1 = the
        // linear epilogue with group partials, statistics exchanged with the peer CTA through DSMEM, pass 2 re-reads
        // the fp32 result and writes the consumer's fp16 A operand
        GnfCtx gx{reinterpret_cast<float2*>(gnf), reinterpret_cast<float2*>(gnf + GnfSmem<256>::OFF_GSTAT),
                  reinterpret_cast<float2*>(gnf + GnfSmem<256>::OFF_XCHG), gb, gb + 256,
                  reinterpret_cast<uint64_t*>(gnf + GnfSmem<256>::OFF_BAR), 0u, 0u, ptx::cluster_ctarank(), warp, 0};
        epi_tile<256, 1, true, true, false, true, false, true, false, false, true>(cx, gx.pstat, quad);
        gnf_fold<256, 1>(p.g, gx, lane);
        epi_tile_gnf_dual_pass2<256, 1, true>(cx, gx);
      } else {
        epi_tile<256, 1, true, true, false, true, false, true, false>(cx);
      }
    } else {
      // O / rowsum -> fp16 -> swizzled staging (this warp: 32 rows x 128 columns) -> full-line global stores
      const uint32_t stg = ptx::smem_u32(sQ) + warp * (32 * 256);
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        ptx::tmem_ld_32x32b_x32(t_o + c0, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 4; ++u)       // 16-byte unit (c0 / 8 + u) of the 256-byte staging row
          sts_scaled8(stg + lane * 256 + ((((c0 >> 3) + u) ^ (lane & 15)) << 4), r + 8 * u, inv);
      }
      __syncwarp();
      // two rows per instruction: lanes 0-15 row 2i, lanes 16-31 row 2i + 1, 16 bytes each = 256 contiguous bytes per row
      __half* obase = p.out16 + ((long long)b * AT_T + half * AT_Q + quad * 32) * AT_C + ch * 128;
      const int rsel = lane >> 4, unit = lane & 15;
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int rr = 2 * i + rsel;
        uint4 val;
        const uint32_t a = stg + rr * 256 + ((unit ^ (rr & 15)) << 4);
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w) : "r"(a));
        *reinterpret_cast<uint4*>(obase + (long long)rr * AT_C + unit * 8) = val;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---- host ----------------------------------------------------------------------------------------------------------
int attn_fused_supported(int T, int C) { return T == AT_T && C == AT_C; }

int attn_fused_prepare(AttnOp* op) {
  op->prepared = 0;
  if (!attn_fused_supported(op->T, op->C)) return -1;
  const uint64_t dims[3] = {(uint64_t)3 * op->C, (uint64_t)op->T, (uint64_t)op->B};
  const uint32_t box[3] = {64, (uint32_t)AT_Q, 1};
  if (tmap_encode_f16(&op->tm_qkv, op->qkv, 3, dims, box)) return -2;
  op->tm_w3 = op->tm_qkv;
  if (op->w3) {
    if (!op->bias3 || !op->residual || !op->out32 || !op->colstats) return -3;
    const uint64_t wd[2] = {(uint64_t)op->C, (uint64_t)op->C};          // [C_out rows][C_in] fp16, K-major
    const uint32_t wb[2] = {64, 128};
    if (tmap_encode_f16(&op->tm_w3, op->w3, 2, wd, wb)) return -2;
  }
  op->prepared = 1;
  return 0;
}

int attn_fused_launch(const AttnOp* op, int batch, cudaStream_t st) {
  if (!op->prepared || batch < 1 || batch > op->B) return -1;
  static DeviceOnce attr_set;
  if (attr_set.need()) {
    if (cudaFuncSetAttribute(attn256_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM) != cudaSuccess) return -2;
    if (cudaFuncSetAttribute(attn256_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM) != cudaSuccess) return -2;
    attr_set.done();
  }
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.out16 = op->out16;
  a.sc = op->scale * 1.4426950408889634f;
  a.reverse = op->reverse;
  if (op->w3) {
    a.bias3 = op->bias3;
    a.g.residual = op->residual; a.g.out32 = op->out32; a.g.colstats = op->colstats;
    a.g.ldo = op->C; a.g.scale = op->out_scale; a.g.M = batch * op->T; a.g.N = op->C;
    if (op->gn_gamma != nullptr) {
      // + act(GroupNorm(out32)) for the next block: the two CTAs of an image form a cluster and exchange their statistics
      if (!op->gn_beta || !op->gn_out16 || op->gn_groups <= 0 || (op->C / op->gn_groups != 8 && op->C / op->gn_groups != 4 && op->C / op->gn_groups != 16))
        return -4;
      a.g.gn_gamma = op->gn_gamma; a.g.gn_beta = op->gn_beta; a.g.gn_eps = op->gn_eps; a.g.gn_cpg = op->C / op->gn_groups;
      a.g.gn_silu = op->gn_silu; a.g.gn_rpi = op->T; a.g.gn_xc = 2; a.g.gn_xg = nullptr; a.g.out16 = op->gn_out16;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * batch); cfg.blockDim = dim3(AT_THREADS); cfg.dynamicSmemBytes = AT_SMEM; cfg.stream = st;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = pdl_attr(at, 1);
      if (cudaLaunchKernelEx(&cfg, attn256_kernel<true>, op->tm_qkv, op->tm_w3, a) != cudaSuccess) return -3;
      return 0;
    }
    if (launch_k(attn256_kernel<true>, dim3(2 * batch), dim3(AT_THREADS), (size_t)AT_SMEM, st, op->tm_qkv, op->tm_w3, a) != cudaSuccess) return -3;
    return 0;
  }
  if (launch_k(attn256_kernel<false>, dim3(2 * batch), dim3(AT_THREADS), (size_t)AT_SMEM, st, op->tm_qkv, op->tm_w3, a) != cudaSuccess) return -3;
  return 0;
}

}  // namespace gddim
