// K3: GroupNorm apply fused into the q/k/v projection of the attention block (16x16 level: C = 256 channels).
//
// Replaces  h = GroupNorm(x)  (cld_jax/models/layerspp.py:69, no activation)  followed by the three NIN projections
// q, k, v = NIN(h) (layerspp.py:70-72), which the unfused plan runs as a gn_apply pass (x fp32 -> h fp16 through HBM)
// and one N = 768 GEMM that re-reads its A tile once per N tile.  Here one CTA owns 128 rows (tokens) of one image:
//   warps 0-7  (a) transform: read the 128 x 256 fp32 rows of x once, apply the per-(image, channel) scale / shift
//                  that gn_coef_kernel produced, round to fp16 and write the result straight into shared memory in the
//                  K-major 128B-swizzle layout of a UMMA A operand (four 64-channel blocks) -- h never exists in HBM;
//              (b) epilogue: the GEMM kernels' own linear epilogue (gemm_epilogue.cuh) for the three N tiles
//   warp 8     TMA producer: the 768 x 256 weight matrix streams through a ring of four 32 KB slots
//   warp 9     MMA issuer: per N tile 16 MMAs 128 x 256 x 16 on the RESIDENT A operand, two TMEM accumulator stages so
//              the epilogue of one N tile overlaps the MMAs of the next
// Numerics are those of the unfused pair (same x*a+b expression, same fp16 rounding, same accumulation order, same
// epilogue code): the qkv tensor is bit-identical.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"
#include "gemm_epilogue.cuh"

namespace gddim {

namespace {
constexpr int GQ_C = 256;                  // input channels = K
constexpr int GQ_NT = 3;                   // N tiles of 256 columns (q, k, v)
constexpr int GQ_ROWS = 128;               // rows per CTA
constexpr int GQ_THREADS = 320;
constexpr int GQ_ABLK = GQ_ROWS * 128;     // one 64-channel block of A: 128 rows x 128 B
constexpr int GQ_WBLK = 256 * 128;         // one 64-channel block of a weight tile: 256 rows x 128 B
constexpr int GQ_SLOTS = 4;
constexpr int GQ_OFF_W = 4 * GQ_ABLK;
constexpr int GQ_OFF_EPI = GQ_OFF_W + GQ_SLOTS * GQ_WBLK;
constexpr int GQ_OFF_BIAS = GQ_OFF_EPI + EPI_WARPS * 32 * 32 * 4;
constexpr int GQ_OFF_BAR = GQ_OFF_BIAS + 2 * 256 * 4;
constexpr int GQ_SMEM = GQ_OFF_BAR + 128;
static_assert(GQ_SMEM <= SMEM_BUDGET, "gn_qkv shared memory");
// persistent variant: two resident A tiles (the transform of tile i+1 runs while tile i is multiplied and stored), two
// weight slots
constexpr int GP_SLOTS = 2;
constexpr int GP_OFF_W = 2 * 4 * GQ_ABLK;
constexpr int GP_OFF_EPI = GP_OFF_W + GP_SLOTS * GQ_WBLK;
constexpr int GP_OFF_BIAS = GP_OFF_EPI + EPI_WARPS * 32 * 32 * 4;
constexpr int GP_OFF_BAR = GP_OFF_BIAS + 2 * 256 * 4;
constexpr int GP_SMEM = GP_OFF_BAR + 128;
static_assert(GP_SMEM <= SMEM_BUDGET, "gn_qkv (persistent) shared memory");
}  // namespace

struct GnQkvArgs {
  const float* x;        // [M, 256] fp32
  const float* coef;     // [B, 2, 256]: scale then shift per image and channel
  const float* bias;     // [768]
  int rows_per_image;    // T
  int reverse;
  GemmArgs g;            // out16 / ldo / scale / M of the epilogue
};

// transform of one 128-row tile: x -> a x + b -> fp16 -> K-major 128B-swizzle A operand at shared address `sa`
// (four 64-channel blocks); executed by the eight epilogue warps
__device__ __forceinline__ void gq_transform(const GnQkvArgs& p, long long row0, uint32_t sa, int warp, int lane) {
  const int b = (int)(row0 / p.rows_per_image);
  const int c = lane * 8;                                  // this lane's eight channels
  const float4* ca = reinterpret_cast<const float4*>(p.coef + ((long long)b * 2) * GQ_C + c);
  const float4* cb = reinterpret_cast<const float4*>(p.coef + ((long long)b * 2 + 1) * GQ_C + c);
  const float4 a0 = ca[0], a1 = ca[1], b0 = cb[0], b1 = cb[1];
  const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const uint32_t blk = sa + (lane >> 3) * GQ_ABLK;         // 64-channel block of this lane
  const uint32_t unit = lane & 7;
  // warp w handles rows w, w + 8, ...: one 1 KB row per instruction pair, eight rows (8 KB per warp) in flight --
  // the x tile comes from HBM
  constexpr int INFL = 8;
#pragma unroll 1
  for (int r0 = warp; r0 < GQ_ROWS; r0 += INFL * EPI_WARPS) {
    float4 v[INFL][2];
#pragma unroll
    for (int k = 0; k < INFL; ++k) {
      const float* q = p.x + (row0 + r0 + k * EPI_WARPS) * GQ_C + c;
      v[k][0] = __ldg(reinterpret_cast<const float4*>(q));
      v[k][1] = __ldg(reinterpret_cast<const float4*>(q + 4));
    }
#pragma unroll
    for (int k = 0; k < INFL; ++k) {
      const int r = r0 + k * EPI_WARPS;
      const float xv[8] = {v[k][0].x, v[k][0].y, v[k][0].z, v[k][0].w, v[k][1].x, v[k][1].y, v[k][1].z, v[k][1].w};
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = xv[j] * a[j] + bb[j];
      const __half2 h0 = __floats2half2_rn(y[0], y[1]), h1 = __floats2half2_rn(y[2], y[3]);
      const __half2 h2 = __floats2half2_rn(y[4], y[5]), h3 = __floats2half2_rn(y[6], y[7]);
      sts128(blk + r * 128 + ((unit ^ (uint32_t)(r & 7)) << 4), *reinterpret_cast<const uint32_t*>(&h0),
             *reinterpret_cast<const uint32_t*>(&h1), *reinterpret_cast<const uint32_t*>(&h2),
             *reinterpret_cast<const uint32_t*>(&h3));
    }
  }
}

__global__ void __launch_bounds__(GQ_THREADS, 1)
gn_qkv_kernel(const __grid_constant__ CUtensorMap tm_w, const GnQkvArgs p) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sA = smem;
  uint8_t* sW = smem + GQ_OFF_W;
  float* epi_stage = reinterpret_cast<float*>(smem + GQ_OFF_EPI);
  float* epi_bias = reinterpret_cast<float*>(smem + GQ_OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GQ_OFF_BAR);
  uint64_t* a_ready = bars;            // A operand written by the 256 transform threads
  uint64_t* wfull = bars + 1;          // [4] weight block landed
  uint64_t* wempty = bars + 5;         // [4] weight block consumed
  uint64_t* tfull = bars + 9;          // [2] accumulator complete
  uint64_t* tempty = bars + 11;        // [2] accumulator drained by the 8 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = p.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const long long row0 = (long long)tile * GQ_ROWS;

  if (threadIdx.x == 256) {
    ptx::prefetch_tmap(&tm_w);
    ptx::mbar_init(a_ready, 256);
    for (int i = 0; i < GQ_SLOTS; ++i) { ptx::mbar_init(&wfull[i], 1); ptx::mbar_init(&wempty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], EPI_WARPS); }
    ptx::fence_mbar_init();
  }
  if (warp == 9) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  pdl_wait();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 256) {
    // ================= weight producer: 12 blocks (N tile, 64-channel block) through 4 slots =================
    uint32_t phase = 0;
    for (int i = 0; i < GQ_NT * 4; ++i) {
      const int slot = i & 3, nt = i >> 2, kb = i & 3;
      ptx::mbar_wait(&wempty[slot], phase ^ 1);
      ptx::mbar_arrive_expect_tx(&wfull[slot], GQ_WBLK);
      ptx::tma_load_2d(&tm_w, &wfull[slot], sW + slot * GQ_WBLK, kb * 64, nt * 256);
      ptx::tma_load_2d(&tm_w, &wfull[slot], sW + slot * GQ_WBLK + GQ_ABLK, kb * 64, nt * 256 + 128);
      if (slot == 3) phase ^= 1;
    }
  } else if (threadIdx.x == 288) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = ptx::umma_idesc_f16(128, 256);
    ptx::mbar_wait(a_ready, 0);
    ptx::tc_fence_after();
    uint32_t wphase = 0, acc_phase = 0;
    int acc = 0;
    for (int nt = 0; nt < GQ_NT; ++nt) {
      ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      for (int kb = 0; kb < 4; ++kb) {
        ptx::mbar_wait(&wfull[kb], wphase);
        ptx::tc_fence_after();
        const uint64_t a_desc = ptx::umma_desc_sw128(ptx::smem_u32(sA + kb * GQ_ABLK));
        const uint64_t b_desc = ptx::umma_desc_sw128(ptx::smem_u32(sW + kb * GQ_WBLK));
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base + acc * 256, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
        ptx::umma_commit(&wempty[kb]);
      }
      wphase ^= 1;
      ptx::umma_commit(&tfull[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp < EPI_WARPS) {
    // ================= transform: x -> a x + b -> fp16 -> swizzled A operand =================
    gq_transform(p, row0, ptx::smem_u32(sA), warp, lane);
    ptx::fence_proxy_async();          // generic-proxy writes -> visible to the tensor core (async proxy)
    ptx::mbar_arrive(a_ready);
    // ================= epilogue: three N tiles through the shared linear epilogue (fp16 output) =================
    const int quad = warp & 3, group = warp >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int nt = 0; nt < GQ_NT; ++nt) {
      float* bias_s = epi_bias + (nt & 1) * 256;
      bias_s[threadIdx.x] = __ldg(p.bias + nt * 256 + threadIdx.x);
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      EpiCtx<256, 1> cx{p.g, epi_stage + warp * 32 * SmemLayout<256, 1>::EPI_ROW_FLOATS, bias_s, &tfull[acc], acc_phase,
                        tmem_base + (uint32_t(quad * 32) << 16) + acc * 256, row0 + quad * 32, nt * 256, lane, group};
      epi_tile<256, 1, false, false, true, false, false, true, false>(cx);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// Persistent variant (default): <= one CTA per SM walks the tiles c, c + grid, ...  The A operand is double-buffered, so the
// eight epilogue warps transform tile i+1 (HBM reads) while the tensor core multiplies tile i and before they drain its
// accumulators; the weight stream and the two TMEM stages run straight through tile boundaries.  The non-persistent kernel
// above serialises, per 128-row tile, barrier setup + TMEM allocation -> transform -> first MMA -> ... -> last store, four
// times per SM at batch 256.  Same arithmetic, same instruction order per tile: bit-identical results.
__global__ void __launch_bounds__(GQ_THREADS, 1)
gn_qkv_persist_kernel(const __grid_constant__ CUtensorMap tm_w, const GnQkvArgs p, const int tiles) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sW = smem + GP_OFF_W;
  float* epi_stage = reinterpret_cast<float*>(smem + GP_OFF_EPI);
  float* epi_bias = reinterpret_cast<float*>(smem + GP_OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GP_OFF_BAR);
  uint64_t* a_ready = bars;            // [2] A buffer written by the 256 transform threads
  uint64_t* a_free = bars + 2;         // [2] every MMA that reads the A buffer has completed
  uint64_t* wfull = bars + 4;          // [2] weight block landed
  uint64_t* wempty = bars + 6;         // [2] weight block consumed
  uint64_t* tfull = bars + 8;          // [2] accumulator complete
  uint64_t* tempty = bars + 10;        // [2] accumulator drained by the 8 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto tile_of = [&](int seq) { return p.reverse ? tiles - 1 - seq : seq; };

  if (threadIdx.x == 256) {
    ptx::prefetch_tmap(&tm_w);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&a_ready[i], 256); ptx::mbar_init(&a_free[i], 1);
      ptx::mbar_init(&wfull[i], 1); ptx::mbar_init(&wempty[i], 1);
      ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 9) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  pdl_wait();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 256) {
    // ================= weight producer: per tile 12 blocks (N tile, 64-channel block) through 2 slots =================
    uint32_t wi = 0;
    for (int seq = blockIdx.x; seq < tiles; seq += gridDim.x) {
      for (int i = 0; i < GQ_NT * 4; ++i, ++wi) {
        const int slot = wi & 1, nt = i >> 2, kb = i & 3;
        ptx::mbar_wait(&wempty[slot], ((wi >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&wfull[slot], GQ_WBLK);
        ptx::tma_load_2d(&tm_w, &wfull[slot], sW + slot * GQ_WBLK, kb * 64, nt * 256);
        ptx::tma_load_2d(&tm_w, &wfull[slot], sW + slot * GQ_WBLK + GQ_ABLK, kb * 64, nt * 256 + 128);
      }
    }
  } else if (threadIdx.x == 288) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = ptx::umma_idesc_f16(128, 256);
    uint32_t wi = 0, acc_phase = 0;
    int acc = 0, it = 0;
    for (int seq = blockIdx.x; seq < tiles; seq += gridDim.x, ++it) {
      const int buf = it & 1;
      ptx::mbar_wait(&a_ready[buf], (it >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t sa = ptx::smem_u32(smem) + buf * 4 * GQ_ABLK;
      for (int nt = 0; nt < GQ_NT; ++nt) {
        ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        for (int kb = 0; kb < 4; ++kb, ++wi) {
          const int slot = wi & 1;
          ptx::mbar_wait(&wfull[slot], (wi >> 1) & 1);
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::umma_desc_sw128(sa + kb * GQ_ABLK);
          const uint64_t b_desc = ptx::umma_desc_sw128(ptx::smem_u32(sW + slot * GQ_WBLK));
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base + acc * 256, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          ptx::umma_commit(&wempty[slot]);
        }
        ptx::umma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      ptx::umma_commit(&a_free[buf]);
    }
  } else if (warp < EPI_WARPS) {
    const int quad = warp & 3, group = warp >> 2;
    int acc = 0, it = 0, bias_buf = 0;
    uint32_t acc_phase = 0;
    if ((int)blockIdx.x < tiles) {
      gq_transform(p, (long long)tile_of(blockIdx.x) * GQ_ROWS, ptx::smem_u32(smem), warp, lane);
      ptx::fence_proxy_async();          // generic-proxy writes -> visible to the tensor core (async proxy)
      ptx::mbar_arrive(&a_ready[0]);
    }
    for (int seq = blockIdx.x; seq < tiles; seq += gridDim.x, ++it) {
      const long long row0 = (long long)tile_of(seq) * GQ_ROWS;
      // ---- the next tile's A operand first: its HBM reads overlap this tile's MMAs ----
      const int nseq = seq + gridDim.x;
      if (nseq < tiles) {
        const int nit = it + 1, nb = nit & 1;
        if (nit >= 2) ptx::mbar_wait(&a_free[nb], ((nit >> 1) - 1) & 1);      // tile nit - 2 has been multiplied
        gq_transform(p, (long long)tile_of(nseq) * GQ_ROWS, ptx::smem_u32(smem) + nb * 4 * GQ_ABLK, warp, lane);
        ptx::fence_proxy_async();
        ptx::mbar_arrive(&a_ready[nb]);
      }
      // ---- epilogue: three N tiles through the shared linear epilogue (fp16 output) ----
      for (int nt = 0; nt < GQ_NT; ++nt) {
        // (bias + staging rows alternate with every N tile across tile boundaries: a warp is at most one barrier ahead)
        float* bias_s = epi_bias + bias_buf * 256;
        bias_buf ^= 1;
        bias_s[threadIdx.x] = __ldg(p.bias + nt * 256 + threadIdx.x);
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        EpiCtx<256, 1> cx{p.g, epi_stage + warp * 32 * SmemLayout<256, 1>::EPI_ROW_FLOATS, bias_s, &tfull[acc], acc_phase,
                          tmem_base + (uint32_t(quad * 32) << 16) + acc * 256, row0 + quad * 32, nt * 256, lane, group};
        epi_tile<256, 1, false, false, true, false, false, true, false>(cx);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---- host ----------------------------------------------------------------------------------------------------------
int gn_qkv_supported(int T, int C, int N) { return C == GQ_C && N == GQ_NT * 256 && T % GQ_ROWS == 0; }

int gn_qkv_prepare(GnQkvOp* op) {
  op->prepared = 0;
  if (!gn_qkv_supported(op->T, op->C, op->N)) return -1;
  const uint64_t wd[2] = {(uint64_t)op->C, (uint64_t)op->N};            // [N rows][C] fp16, K-major
  const uint32_t wb[2] = {64, 128};
  if (tmap_encode_f16(&op->tm_w, op->w, 2, wd, wb)) return -2;
  op->prepared = 1;
  return 0;
}

int gn_qkv_launch(const GnQkvOp* op, int batch, cudaStream_t st) {
  if (!op->prepared || batch < 1 || batch > op->B) return -1;
  static DeviceOnce attr_set;
  if (attr_set.need()) {
    if (cudaFuncSetAttribute(gn_qkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GQ_SMEM) != cudaSuccess) return -2;
    if (cudaFuncSetAttribute(gn_qkv_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GP_SMEM) != cudaSuccess) return -2;
    attr_set.done();
  }
  GnQkvArgs a;
  memset(&a, 0, sizeof(a));
  a.x = op->x; a.coef = op->coef; a.bias = op->bias; a.rows_per_image = op->T; a.reverse = op->reverse;
  a.g.out16 = op->out16; a.g.ldo = op->N; a.g.scale = 1.f; a.g.M = batch * op->T; a.g.N = op->N;
  const int tiles = batch * op->T / GQ_ROWS;
  static const bool persist = [] { const char* e = getenv("GDDIM_GNQKV_PERSIST"); return !(e && e[0] == '0'); }();   // A/B switch
  if (persist) {
    const int sms = device_sm_count();
    if (launch_k(gn_qkv_persist_kernel, dim3(tiles < sms ? tiles : sms), dim3(GQ_THREADS), (size_t)GP_SMEM, st, op->tm_w, a, tiles) != cudaSuccess)
      return -3;
    return 0;
  }
  if (launch_k(gn_qkv_kernel, dim3(tiles), dim3(GQ_THREADS), (size_t)GQ_SMEM, st, op->tm_w, a) != cudaSuccess) return -3;
  return 0;
}

}  // namespace gddim
