// K1x (EXPERIMENTAL, off unless GDDIM_XF=1; not yet validated on hardware): 3x3 convolution of the 32x32 level with
// the preceding GroupNorm + swish applied ON LOAD -- the normalised fp16 tensor never exists in HBM and the separate
// gn_apply pass (HBM-bound, 33 % of an evaluation together with its siblings) disappears for these layers.
//
// Same tile geometry as the halo variant of conv_gemm_umma_kernel (BLOCK_N = 128, 256-row tiles = 8 whole image rows,
// CTA pairs with cta_group::2 MMAs), but the A operand is produced inside the CTA:
//   warps 0-7  for every 64-channel block kc: read the (8 + 2) x 32 x 64 fp32 halo box of the source once, apply the
//              per-(image, channel) scale / shift of gn_coef_kernel and swish, round to fp16 and write THREE copies --
//              the x-shifts dx = -1, 0, +1 with their zero columns -- in the K-major 128B-swizzle layout into three
//              40 KB buffers (the y-shifts are descriptor offsets of 32 rows, as in the halo kernel).  Between the two
//              transform phases of a tile the same warps drain the PREVIOUS tile's accumulator in slices (the shared
//              linear epilogue, split by chunk range), so the tensor pipe works on block kc while the epilogue of the
//              tile before and the loads of block kc + 1 proceed.
//   warp 8     TMA producer: weights only (ring of three slots = the three dy tiles of one (dx, kc))
//   warp 9     MMA issuer (leader CTA): per kc nine tap MMAs groups on the resident A buffers, then releases them
// Expected ("what comes next"): conv time about unchanged (transform ~3 k + MMA 4.6 k cycles per
// channel block against 14.7 k cycles per tile today), GroupNorm apply (7.7 k cycles per tile equivalent) gone.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"
#include "gemm_epilogue.cuh"

namespace gddim {

namespace {
constexpr int XF_BN = 128;
constexpr int XF_MT = 2;
constexpr int XF_W = 32;                                  // image width the thread mapping is written for
constexpr int XF_TILE_ROWS = XF_MT * BLOCK_M;             // 256 output pixels = 8 image rows
constexpr int XF_HALO = XF_TILE_ROWS / XF_W + 2;          // 10 image rows per A buffer
constexpr int XF_ABUF = XF_HALO * XF_W * 128;             // 40 KB: one x-shift of one 64-channel block
constexpr int XF_BTILE = (XF_BN / 2) * BLOCK_K * 2;       // 8 KB: this CTA's half of one tap's weight tile
constexpr int XF_BSLOT = 3 * XF_BTILE;                    // the three dy tiles of one (dx, kc)
constexpr int XF_NSLOT = 3;
constexpr int XF_OFF_B = 3 * XF_ABUF;
constexpr int XF_OFF_EPI = XF_OFF_B + XF_NSLOT * XF_BSLOT;
constexpr int XF_OFF_BIAS = XF_OFF_EPI + EPI_WARPS * 32 * 32 * 4;
constexpr int XF_OFF_BAR = XF_OFF_BIAS + 2 * XF_BN * 4;
constexpr int XF_SMEM = XF_OFF_BAR + 256;
static_assert(XF_SMEM <= SMEM_BUDGET, "conv_xf shared memory");
constexpr int XF_NCH = XF_BN / 32;                        // 32-column chunks per sub-tile
constexpr int XF_NJ = XF_MT * XF_NCH / EPI_GROUPS;        // chunk iterations per epilogue group and tile

// cluster-scope release / acquire: the follower CTA's generic-proxy writes of its A buffers are published through the
// LEADER's barrier, whose waiter (the MMA thread) then reads both CTAs' shared memory
__device__ __forceinline__ void xf_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void xf_wait_acquire_cluster(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "XF_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra XF_DONE;\n\t"
      "bra XF_WAIT_LOOP;\n\t"
      "XF_DONE:\n\t"
      "}\n" ::"r"(ptx::smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ float xf_silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ uint32_t xf_pack(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// The linear epilogue of gemm_epilogue.cuh restricted to fp32 output + column statistics (+ residual), full tiles,
// over the chunk iterations [j_lo, j_hi) of this warp's group; `first` waits for the accumulator.
template <bool RES>
__device__ __forceinline__ void xf_epi_range(const EpiCtx<XF_BN, XF_MT>& cx, int j_lo, int j_hi, bool first) {
  constexpr int RS = SmemLayout<XF_BN, XF_MT>::EPI_ROW_FLOATS;
  const GemmArgs& p = cx.p;
  const int lane = cx.lane;
  const int rsub = lane >> 3;
  const int c4 = (lane & 7) * 4;
  const uint32_t stg_w = ptx::smem_u32(cx.stg) + lane * RS * 4;
  const uint32_t wx = lane & 7;
  const uint32_t stg_r0 = ptx::smem_u32(cx.stg) + rsub * RS * 4 + (((lane & 7) ^ rsub) << 4);
  const uint32_t stg_r1 = ptx::smem_u32(cx.stg) + (rsub + 4) * RS * 4 + (((lane & 7) ^ (rsub + 4)) << 4);
  const uint32_t bias_a = ptx::smem_u32(cx.bias_s) + c4 * 4;
  const long long ldo = p.ldo;
  const float scale = p.scale;
  if (first) {
    ptx::mbar_wait(cx.tfull, cx.tfull_phase);
    ptx::tc_fence_after();
  }
  uint32_t r[32];
#pragma unroll 1
  for (int j = j_lo; j < j_hi; ++j) {
    const int q = cx.group + j * EPI_GROUPS;
    const int mi = q / XF_NCH, c0 = (q % XF_NCH) * 32;
    ptx::tmem_ld_32x32b_x32(cx.taddr + mi * XF_BN + c0, r);
    const long long mb = cx.m0 + (long long)mi * BLOCK_M + rsub;
    const int n0 = cx.n_tile0 + c0 + c4;
    float4 res[8];
    if (RES) {
      const float* rb = p.residual + mb * ldo + n0;
#pragma unroll
      for (int i = 0; i < 8; ++i) res[i] = __ldg(reinterpret_cast<const float4*>(rb + (long long)(i * 4) * ldo));
    }
    ptx::tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 8; ++k) sts128(stg_w + ((k ^ wx) << 4), r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
    __syncwarp();
    float4 bsum = lds128(bias_a + c0 * 4);
    bsum.x *= scale; bsum.y *= scale; bsum.z *= scale; bsum.w *= scale;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f), cq = make_float4(0.f, 0.f, 0.f, 0.f);
    float* o32 = p.out32 + mb * ldo + n0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 v = lds128(((i & 1) ? stg_r1 : stg_r0) + (i >> 1) * 8 * RS * 4);
      if (RES) { v.x += res[i].x; v.y += res[i].y; v.z += res[i].z; v.w += res[i].w; }
      v.x = fmaf(v.x, scale, bsum.x); v.y = fmaf(v.y, scale, bsum.y);
      v.z = fmaf(v.z, scale, bsum.z); v.w = fmaf(v.w, scale, bsum.w);
      cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
      cq.x = fmaf(v.x, v.x, cq.x); cq.y = fmaf(v.y, v.y, cq.y); cq.z = fmaf(v.z, v.z, cq.z); cq.w = fmaf(v.w, v.w, cq.w);
      *reinterpret_cast<float4*>(o32 + (long long)(i * 4) * ldo) = v;
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
      cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
      cq.x += __shfl_xor_sync(0xffffffffu, cq.x, o); cq.y += __shfl_xor_sync(0xffffffffu, cq.y, o);
      cq.z += __shfl_xor_sync(0xffffffffu, cq.z, o); cq.w += __shfl_xor_sync(0xffffffffu, cq.w, o);
    }
    const long long mrow0 = cx.m0 + (long long)mi * BLOCK_M;
    if (lane < 8) {
      float* cp = p.colstats + ((mrow0 >> 5) * 2) * ldo + n0;
      *reinterpret_cast<float4*>(cp) = cs;
      *reinterpret_cast<float4*>(cp + ldo) = cq;
    }
    __syncwarp();
  }
}
}  // namespace

struct XfArgs {
  GemmArgs g;
  const float* src1; int c1;      // fp32 NHWC sources, channel-concatenated (src2 optional); c1, c2 multiples of 64
  const float* src2; int c2;
  const float* coef;              // [B, 2, c1 + c2] scale then shift (gn_coef_kernel)
  int silu;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_xf_kernel(const __grid_constant__ CUtensorMap tmB, const XfArgs p) {
  pdl_launch_dependents();
  const GemmArgs& g = p.g;
  const uint32_t cta_rank = ptx::cluster_ctarank();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sA = smem;                                   // [3 x-shifts][10 rows][32 px][128 B]
  uint8_t* sB = smem + XF_OFF_B;
  float* epi_stage = reinterpret_cast<float*>(smem + XF_OFF_EPI);
  float* epi_bias = reinterpret_cast<float*>(smem + XF_OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + XF_OFF_BAR);
  uint64_t* bfull = bars;             // [3] weights landed (leader's barrier, both CTAs' bytes)
  uint64_t* bempty = bars + 3;        // [3] weights consumed
  uint64_t* tfull = bars + 6;         // [2]
  uint64_t* tempty = bars + 8;        // [2] (leader's barrier, both CTAs' epilogue warps)
  uint64_t* a_full = bars + 10;       // A buffers written (leader's barrier, both CTAs' transform warps)
  uint64_t* a_empty = bars + 11;      // A buffers consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.c1 + p.c2;
  const int KC = C / BLOCK_K;

  if (threadIdx.x == PRODUCER_THREAD) {
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < XF_NSLOT; ++s) { ptx::mbar_init(&bfull[s], 2); ptx::mbar_init(&bempty[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull[s], 1); ptx::mbar_init(&tempty[s], EPI_WARPS * 2); }
    ptx::mbar_init(a_full, EPI_WARPS * 2);
    ptx::mbar_init(a_empty, 1);
    ptx::fence_mbar_init();
  }
  if (warp == MMA_WARP) { ptx::tmem_alloc_2cta(tmem_slot, 512); ptx::tmem_relinquish_2cta(); }
  pdl_wait();
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int unit_m = (g.m_tiles + 1) / 2;               // pairs of vertically adjacent 256-row tiles
  const int num_tiles = unit_m;                          // N = BLOCK_N: one N tile
  const int tile0 = blockIdx.x / 2, tile_step = gridDim.x / 2;

  if (threadIdx.x == PRODUCER_THREAD) {
    // ================= weight producer =================
    int slot = 0;
    uint32_t phase = 0;
    const uint32_t nrow0 = cta_rank * (XF_BN / 2);
    for (int tile_i = tile0; tile_i < num_tiles; tile_i += tile_step) {
      for (int kc = 0; kc < KC; ++kc) {
        for (int dxi = 0; dxi < 3; ++dxi) {
          uint8_t* dst = sB + slot * XF_BSLOT;
          ptx::mbar_wait(&bempty[slot], phase ^ 1);
          const uint32_t fb = ptx::mapa(ptx::smem_u32(&bfull[slot]), 0);
          if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&bfull[slot], 2 * XF_BSLOT);
#pragma unroll
          for (int dyi = 0; dyi < 3; ++dyi)
            ptx::tma_load_4d_2cta(&tmB, fb, dst + dyi * XF_BTILE, g.w_koff + (dyi * 3 + dxi) * C + kc * BLOCK_K, (int)nrow0, 0, 0);
          if (cta_rank != 0) ptx::mbar_arrive_cluster(fb);
          if (++slot == XF_NSLOT) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (threadIdx.x == MMA_THREAD && cta_rank == 0) {
    // ================= MMA issuer (leader) =================
    constexpr uint32_t idesc = ptx::umma_idesc_f16(BLOCK_M * 2, XF_BN);
    int slot = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0, a_par = 0;
    for (int tile_i = tile0; tile_i < num_tiles; tile_i += tile_step) {
      ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * XF_MT * XF_BN;
      uint32_t accum = 0;
      for (int kc = 0; kc < KC; ++kc) {
        xf_wait_acquire_cluster(a_full, a_par);
        a_par ^= 1;
        ptx::tc_fence_after();
        for (int dxi = 0; dxi < 3; ++dxi) {
          ptx::mbar_wait(&bfull[slot], phase);
          ptx::tc_fence_after();
          const uint32_t a_base = ptx::smem_u32(sA + dxi * XF_ABUF);
          const uint32_t b_base = ptx::smem_u32(sB + slot * XF_BSLOT);
#pragma unroll
          for (int dyi = 0; dyi < 3; ++dyi) {
            const uint64_t b_desc = ptx::umma_desc_sw128(b_base + dyi * XF_BTILE);
#pragma unroll
            for (int mi = 0; mi < XF_MT; ++mi) {
              const uint64_t a_desc = ptx::umma_desc_sw128(a_base + dyi * XF_W * 128 + mi * A_TILE_BYTES);
#pragma unroll
              for (int k = 0; k < BLOCK_K / 16; ++k)
                ptx::umma_f16_2cta(d_tmem + mi * XF_BN, a_desc + 2 * k, b_desc + 2 * k, idesc, accum | k);
            }
            accum = 1;
          }
          ptx::umma_commit_2cta(&bempty[slot], 3);
          if (++slot == XF_NSLOT) { slot = 0; phase ^= 1; }
        }
        ptx::umma_commit_2cta(a_empty, 3);              // the three A buffers may be rewritten (both CTAs)
      }
      ptx::umma_commit_2cta(&tfull[acc], 3);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp < EPI_WARPS) {
    // ================= transform + epilogue warps =================
    const int quad = warp & 3, group = warp >> 2;
    const int px = threadIdx.x >> 3, gch = threadIdx.x & 7;     // this thread's pixel column and 8-channel group
    uint32_t a_cnt = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int bias_buf = 0;
    bool have_prev = false;
    long long prev_m0 = 0;
    const uint32_t a_full_leader = ptx::mapa(ptx::smem_u32(a_full), 0);
    const uint32_t tempty_leader0 = ptx::mapa(ptx::smem_u32(&tempty[0]), 0);
    const uint32_t tempty_leader1 = ptx::mapa(ptx::smem_u32(&tempty[1]), 0);

    auto epi_part = [&](int part, int parts) {
      // slice `part` of `parts` of the previous tile's epilogue
      float* bias_s = epi_bias + bias_buf * XF_BN;
      if (part == 0) {
        if (threadIdx.x < XF_BN) {
          float bsum = 0.f;
          if (g.bias != nullptr) bsum += __ldg(g.bias + threadIdx.x);
          if (g.bias2 != nullptr) bsum += __ldg(g.bias2 + threadIdx.x);
          bias_s[threadIdx.x] = bsum;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      }
      EpiCtx<XF_BN, XF_MT> cx{g, epi_stage + warp * 32 * 32, bias_s, &tfull[acc], acc_phase,
                               tmem_base + (uint32_t(quad * 32) << 16) + acc * XF_MT * XF_BN, prev_m0 + quad * 32, 0, lane, group};
      const int j_lo = part * XF_NJ / parts, j_hi = (part + 1) * XF_NJ / parts;
      if (g.residual != nullptr) xf_epi_range<true>(cx, j_lo, j_hi, part == 0);
      else xf_epi_range<false>(cx, j_lo, j_hi, part == 0);
      if (part == parts - 1) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(acc == 0 ? tempty_leader0 : tempty_leader1);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        bias_buf ^= 1;
      }
    };

    for (int tile_i = tile0; tile_i < num_tiles; tile_i += tile_step) {
      const int tile = g.reverse ? num_tiles - 1 - tile_i : tile_i;
      const int mt = tile * 2 + (int)cta_rank;
      const long long p0 = (long long)mt * XF_TILE_ROWS;           // first output pixel of this CTA's tile
      const int h0 = (int)((p0 / XF_W) % g.H);
      const long long b0 = p0 / ((long long)XF_W * g.H);
      const bool tile_valid = mt < g.m_tiles;
      for (int kc = 0; kc < KC; ++kc) {
        // ---- source of this 64-channel block ----
        const float* src; int cs, coff;
        if (kc * BLOCK_K < p.c1) { src = p.src1; cs = p.c1; coff = kc * BLOCK_K; }
        else { src = p.src2; cs = p.c2; coff = kc * BLOCK_K - p.c1; }
        const int ch = kc * BLOCK_K + gch * 8;
        float a[8], bb[8];
        {
          const long long bi = tile_valid ? b0 : 0;
          const float4* ca = reinterpret_cast<const float4*>(p.coef + (bi * 2) * C + ch);
          const float4* cb = reinterpret_cast<const float4*>(p.coef + (bi * 2 + 1) * C + ch);
          const float4 a0 = ca[0], a1 = ca[1], c0v = cb[0], c1v = cb[1];
          a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
          bb[0] = c0v.x; bb[1] = c0v.y; bb[2] = c0v.z; bb[3] = c0v.w; bb[4] = c1v.x; bb[5] = c1v.y; bb[6] = c1v.z; bb[7] = c1v.w;
        }
        // ---- loads: halo row i = image row h0 - 1 + i, this thread's pixel column, 8 channels; all in flight ----
        float4 v[XF_HALO][2];
#pragma unroll
        for (int i = 0; i < XF_HALO; ++i) {
          const int y = h0 - 1 + i;
          if (tile_valid && y >= 0 && y < g.H) {
            const float* q = src + ((b0 * g.H + y) * XF_W + px) * cs + coff + gch * 8;
            v[i][0] = __ldg(reinterpret_cast<const float4*>(q));
            v[i][1] = __ldg(reinterpret_cast<const float4*>(q + 4));
          } else {
            v[i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
            v[i][1] = v[i][0];
          }
        }
        // ---- the A buffers are free once the MMAs of the previous channel block have retired ----
        ptx::mbar_wait(a_empty, (a_cnt & 1) ^ 1);
        ++a_cnt;
        const uint32_t bufm = ptx::smem_u32(sA), buf0 = bufm + XF_ABUF, bufp = bufm + 2 * XF_ABUF;   // dx = -1, 0, +1
#pragma unroll
        for (int i = 0; i < XF_HALO; ++i) {
          const int y = h0 - 1 + i;
          uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
          if (tile_valid && y >= 0 && y < g.H) {
            const float xv[8] = {v[i][0].x, v[i][0].y, v[i][0].z, v[i][0].w, v[i][1].x, v[i][1].y, v[i][1].z, v[i][1].w};
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { t[j] = xv[j] * a[j] + bb[j]; if (p.silu) t[j] = xf_silu(t[j]); }
            w0 = xf_pack(t[0], t[1]); w1 = xf_pack(t[2], t[3]); w2 = xf_pack(t[4], t[5]); w3 = xf_pack(t[6], t[7]);
          }
          // buffer dx holds source pixel (x + dx) at position x: this pixel goes to position px - dx
          const int r0 = i * XF_W + px;
          sts128(buf0 + r0 * 128 + ((gch ^ (r0 & 7)) << 4), w0, w1, w2, w3);
          if (px + 1 < XF_W) { const int r = r0 + 1; sts128(bufm + r * 128 + ((gch ^ (r & 7)) << 4), w0, w1, w2, w3); }
          else { const int r = i * XF_W; sts128(bufm + r * 128 + ((gch ^ (r & 7)) << 4), 0u, 0u, 0u, 0u); }         // zero column x = 0
          if (px >= 1) { const int r = r0 - 1; sts128(bufp + r * 128 + ((gch ^ (r & 7)) << 4), w0, w1, w2, w3); }
          else { const int r = i * XF_W + XF_W - 1; sts128(bufp + r * 128 + ((gch ^ (r & 7)) << 4), 0u, 0u, 0u, 0u); }   // zero column x = W-1
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) xf_arrive_release_cluster(a_full_leader);
        // ---- a slice of the previous tile's epilogue while the tensor pipe works on this block ----
        if (have_prev) epi_part(kc, KC);
      }
      have_prev = true;
      prev_m0 = p0;
    }
    if (have_prev)
      for (int part = 0; part < KC; ++part) epi_part(part, KC);
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == MMA_WARP) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta(tmem_base, 512);
  }
}

// ---- host ----------------------------------------------------------------------------------------------------------
// `op` must have been prepared by gemm_prepare as a halo / CTA-pair layer (block_n 128, 256-row tiles): its weight
// tensor map (box = 64 channels x 64 rows) is reused.
int conv_xf_supported(const GemmOp* op) {
  return op->prepared && op->halo && op->cg == 2 && op->block_n == XF_BN && op->m_sub == XF_MT && op->nseg == 1 &&
         op->seg[0].taps == 9 && op->W == XF_W && op->H % (XF_TILE_ROWS / XF_W) == 0 && op->N == XF_BN && op->epi == EPI_LINEAR &&
         op->out32 != nullptr && op->colstats != nullptr && op->out16 == nullptr && op->rowscale == nullptr &&
         op->xf_src1 != nullptr && op->xf_c1 % BLOCK_K == 0 && op->xf_c2 % BLOCK_K == 0 && op->xf_c1 + op->xf_c2 == op->seg[0].c &&
         op->seg[0].c_off == 0 && op->seg[0].c_total == op->seg[0].c && op->m_tiles % 2 == 0;
}

int conv_xf_launch(const GemmOp* op, cudaStream_t st) {
  if (!conv_xf_supported(op)) return -1;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_xf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XF_SMEM) != cudaSuccess) return -2;
    attr_set = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  XfArgs a;
  memset(&a, 0, sizeof(a));
  a.g.H = op->H; a.g.W = op->W;
  a.g.M = op->B * op->H * op->W; a.g.N = op->N;
  a.g.m_tiles = op->m_tiles; a.g.n_tiles = 1;
  a.g.w_koff = op->w_koff;
  a.g.bias = op->bias; a.g.bias2 = op->bias2; a.g.residual = op->residual;
  a.g.scale = op->scale; a.g.out32 = op->out32; a.g.ldo = op->ldo; a.g.colstats = op->colstats;
  a.g.reverse = op->reverse;
  a.src1 = op->xf_src1; a.c1 = op->xf_c1; a.src2 = op->xf_src2; a.c2 = op->xf_c2;
  a.coef = op->xf_coef; a.silu = op->xf_silu;
  if ((long long)op->B * op->H * op->W % XF_TILE_ROWS != 0 || (op->m_tiles & 1)) return -3;     // whole tiles, whole pairs
  const int units = op->m_tiles / 2;
  const int pairs = units < num_sms / 2 ? units : num_sms / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = XF_SMEM; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_attr(at, 1);
  if (cudaLaunchKernelEx(&cfg, conv_xf_kernel, op->tmB, a) != cudaSuccess) return -4;
  return 0;
}

}  // namespace gddim
