// K1: implicit-GEMM convolution / GEMM on the 5th-generation tensor cores (sm_100a).
//
// Replaces the XLA conv_general_dilated / dot_general calls that the reference issues from
//   cld_jax/models/layers.py:66-107 (ddpm_conv1x1 / ddpm_conv3x3), layers.py:467-478 (NIN),
//   cld_jax/models/layerspp.py:74-78 (attention einsums).
//
// Design: persistent CTAs (one per SM), 10 warps:
//   warps 0-7 epilogue     - tcgen05.ld -> swizzled smem transposition -> bias / temb / residual / scale / column
//                            statistics (or row softmax) -> full-line global stores; two groups of four warps take
//                            alternate 32-column chunks of every tile
//   warp 8   TMA producer  - one 4-D box load per (tap, 64-channel chunk) builds the im2col A tile directly in
//                            shared memory (shifted box + hardware zero fill = SAME padding); HALO kernels load one
//                            (rows + 2)-row box per x-shift and take the y-shifts as descriptor offsets
//   warp 9   MMA issuer    - tcgen05.mma.kind::f16, 128 x BLOCK_N x 16 (cta_group::2: 256 x BLOCK_N x 16 across a
//                            CTA pair, each CTA staging half the weight tile), accumulators in TMEM (2 stages)
// Operands are fp16 with fp32 accumulation; smem tiles use the 128-byte swizzle (K-major).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"
#include "gemm_epilogue.cuh"

namespace gddim {

static thread_local char g_gemm_err[512] = "";
const char* gemm_last_error() { return g_gemm_err; }
#define GEMM_FAIL(...)                                          \
  do {                                                          \
    snprintf(g_gemm_err, sizeof(g_gemm_err), __VA_ARGS__);      \
    return -1;                                                  \
  } while (0)


// HALO (3x3 convolutions whose CTA tile is a block of whole image rows): one TMA box of (tile rows + 2) image rows is
// loaded per x-shift and channel block, and the three y-shifted operands are just descriptor start addresses W rows
// apart inside it -- the A operand crosses L2->smem three times per tile instead of nine.
template <int BLOCK_N, int EPI, int MT, int CG, bool HALO>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_umma_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                      const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmH, const GemmArgs p) {
  pdl_launch_dependents();   // the wait is after the prologue
  GDDIM_STAMP(p, threadIdx.x == 0, 0, 11);
  constexpr bool GNF = EPI == EPI_GNF;
  using L = SmemLayout<BLOCK_N, MT, CG, GNF ? GnfSmem<BLOCK_N>::BYTES : 0>;
  static_assert(CG == 1 || (EPI != EPI_SOFTMAX && (BLOCK_N / CG) % 16 == 0), "CTA pairs: linear / GNF epilogue only");
  static_assert(!HALO || EPI != EPI_SOFTMAX, "halo tiles: linear / GNF epilogue only");
  const uint32_t cta_rank = CG == 2 ? ptx::cluster_ctarank() : 0u;     // rank 0 = leader: issues the MMAs
  // GNF with images spanning several CTAs: the launch is a cluster of gn_xc CTAs (CG == 2: the MMA pair itself)
  const bool clustered = CG == 2 || (GNF && p.gn_xc > 1 && p.gn_xg == nullptr);
  constexpr int MAX_STAGES = 8;
  const int STAGES = HALO ? p.stages : L::STAGES;
  const int stage_bytes = HALO ? p.stage_bytes : L::STAGE_BYTES;
  const int a_bytes = HALO ? p.a_bytes : MT * A_TILE_BYTES;            // then up to three weight tiles
  constexpr uint32_t TMEM_COLS = (2 * MT * BLOCK_N) < 32 ? 32 : (2 * MT * BLOCK_N);   // 2 accumulator stages
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns must be a power of two <= 512");

  extern __shared__ __align__(1024) uint8_t smem[];      // SWIZZLE_128B tiles need 1024-byte alignment
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  // ring slot s: [A operand (a_bytes)] [weight tile(s)]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* full_bar = bars;                          // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + MAX_STAGES;            // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * MAX_STAGES;        // [2]       MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * MAX_STAGES + 2;   // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * stage_bytes + L::BAR_BYTES);
  float* epi_bias = reinterpret_cast<float*>(smem + STAGES * stage_bytes + L::BAR_BYTES + L::EPI_STAGE_BYTES);
  uint8_t* gnf_smem = smem + STAGES * stage_bytes + L::BAR_BYTES + L::EPI_STAGE_BYTES + 2 * BLOCK_N * 4;   // GNF only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == PRODUCER_THREAD) {
    ptx::prefetch_tmap(&tmA0);
    if (p.nseg > 1) ptx::prefetch_tmap(&tmA1);
    ptx::prefetch_tmap(&tmB);
    if (HALO) ptx::prefetch_tmap(&tmH);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], CG);             // one arrive per producer of the pair (on the leader's barrier)
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull_bar[s], 1);
      ptx::mbar_init(&tempty_bar[s], (EPI == EPI_SOFTMAX ? 4 : EPI_WARPS) * CG);   // both CTAs' epilogues (leader's barrier)
    }
    if (GNF) {    // statistics exchange: one local arrival (expect_tx) + the bytes of every CTA of the cluster per tile
      ptx::mbar_init(reinterpret_cast<uint64_t*>(gnf_smem + GnfSmem<BLOCK_N>::OFF_BAR), 1);
      ptx::mbar_init(reinterpret_cast<uint64_t*>(gnf_smem + GnfSmem<BLOCK_N>::OFF_BAR) + 1, 1);
    }
    ptx::fence_mbar_init();
    if (p.pf_bytes > 0) {
      const long long per = ((p.pf_bytes + gridDim.x - 1) / gridDim.x + 127) & ~127LL;
      long long off = (long long)blockIdx.x * per;
      const long long end = off + per < p.pf_bytes ? off + per : p.pf_bytes;
      for (; off < end; off += 8192)
        ptx::prefetch_l2_bulk(static_cast<const char*>(p.pf_ptr) + off, (uint32_t)(end - off < 8192 ? end - off : 8192));
    }
  }
  if (warp == MMA_WARP) {
    if (CG == 2) { ptx::tmem_alloc_2cta(tmem_slot, TMEM_COLS); ptx::tmem_relinquish_2cta(); }
    else { ptx::tmem_alloc(tmem_slot, TMEM_COLS); ptx::tmem_relinquish(); }
  }
  // everything above touched only this CTA's shared memory / TMEM; from here on global memory is read and written
  pdl_wait();
  ptx::tc_fence_before();
  if (clustered) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  GDDIM_STAMP(p, threadIdx.x == 0, 0, 12);

  // work unit = CG vertically adjacent CTA tiles x one N tile; CTA `cta_rank` of the pair owns M tile unit_m * CG + rank
  const int unit_m = (p.m_tiles + CG - 1) / CG;
  const int num_tiles = unit_m * p.n_tiles;
  const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;

  if (threadIdx.x == PRODUCER_THREAD) {
    // ================= TMA producer =================
    int stage = 0;
    uint32_t phase = 0;
    // both CTAs of a pair complete their loads on the LEADER's barrier, which expects the bytes of the whole pair
    auto arm = [&](int st, uint32_t bytes) -> uint32_t {
      if (CG == 1) { ptx::mbar_arrive_expect_tx(&full_bar[st], bytes); return ptx::smem_u32(&full_bar[st]); }
      const uint32_t fb = ptx::mapa(ptx::smem_u32(&full_bar[st]), 0);
      if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[st], 2 * bytes);
      return fb;
    };
    auto load = [&](const CUtensorMap* tm, uint32_t fb, uint64_t* own_bar, void* dst, int c0, int c1, int c2, int c3) {
      if (CG == 1) ptx::tma_load_4d(tm, own_bar, dst, c0, c1, c2, c3);
      else ptx::tma_load_4d_2cta(tm, fb, dst, c0, c1, c2, c3);
    };
    for (int tile_i = tile0; tile_i < num_tiles; tile_i += tile_step) {
      const int tile = p.reverse ? num_tiles - 1 - tile_i : tile_i;
      const int mt = (tile / p.n_tiles) * CG + cta_rank, nt = tile % p.n_tiles;
      int w0[MT], h0[MT], b0[MT];
#pragma unroll
      for (int mi = 0; mi < MT; ++mi) {
        const int p0 = (mt * MT + mi) * BLOCK_M;
        w0[mi] = p0 % p.W;
        h0[mi] = (p0 / p.W) % p.H;
        b0[mi] = p0 / (p.W * p.H);
      }
      const int bidx = p.tiles_per_batch > 0 ? mt / p.tiles_per_batch : 0;
      const int nrow0 = nt * BLOCK_N + cta_rank * (BLOCK_N / CG);
      for (int ws = 0; ws < p.wsplit; ++ws) {          // precise mode: A against W_hi, then A again against W_lo
      const int wk0 = p.w_koff + ws * p.ktot;
      int kb = 0;
      for (int s = 0; s < p.nseg; ++s) {
        const CUtensorMap* tm = (s == 0) ? &tmA0 : &tmA1;
        if (HALO && p.taps[s] == 9) {
          // ring slot = one x-shift of one 64-channel block: halo box + the three weight tiles of its y-shifts
          const int cseg = p.kch[s] * BLOCK_K;
          for (int dxi = 0; dxi < 3; ++dxi) {
            for (int kc = 0; kc < p.kch[s]; ++kc) {
              uint8_t* slot = smem + stage * stage_bytes;
              ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
              const uint32_t fb = arm(stage, (uint32_t)(a_bytes + 3 * L::B_TILE_BYTES));
              load(&tmH, fb, &full_bar[stage], slot, p.coff[s] + kc * BLOCK_K, dxi - 1, h0[0] - 1, b0[0]);
#pragma unroll
              for (int dyi = 0; dyi < 3; ++dyi)
                load(&tmB, fb, &full_bar[stage], slot + a_bytes + dyi * L::B_TILE_BYTES,
                     wk0 + kb * BLOCK_K + (dyi * 3 + dxi) * cseg + kc * BLOCK_K, nrow0, bidx, 0);
              if (CG == 2 && cta_rank != 0) ptx::mbar_arrive_cluster(fb);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
          kb += 9 * p.kch[s];
          continue;
        }
        for (int tap = 0; tap < p.taps[s]; ++tap) {
          const int dy = (p.taps[s] == 9) ? tap / 3 - 1 : 0;
          const int dx = (p.taps[s] == 9) ? tap % 3 - 1 : 0;
          for (int kc = 0; kc < p.kch[s]; ++kc, ++kb) {
            uint8_t* slot = smem + stage * stage_bytes;
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            const uint32_t fb = arm(stage, (uint32_t)(MT * A_TILE_BYTES + L::B_TILE_BYTES));
#pragma unroll
            for (int mi = 0; mi < MT; ++mi)
              load(tm, fb, &full_bar[stage], slot + mi * A_TILE_BYTES, p.coff[s] + kc * BLOCK_K, w0[mi] + dx, h0[mi] + dy,
                   b0[mi]);
            load(&tmB, fb, &full_bar[stage], slot + a_bytes, wk0 + kb * BLOCK_K, nrow0, bidx, 0);
            if (CG == 2 && cta_rank != 0) ptx::mbar_arrive_cluster(fb);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
      }
    }
  } else if (threadIdx.x == MMA_THREAD && cta_rank == 0) {
    // ================= MMA issuer (leader CTA of a pair) =================
    constexpr uint32_t idesc = ptx::umma_idesc_f16(BLOCK_M * CG, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int mma_seq = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++mma_seq) {
      GDDIM_STAMP(p, true, mma_seq, 8);
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      GDDIM_STAMP(p, true, mma_seq, 9);
      const uint32_t d_tmem = tmem_base + acc * MT * BLOCK_N;
      uint32_t accum = 0;                             // 0 for the first MMA of every accumulator of the tile
      auto issue = [&](uint32_t a_addr, uint32_t b_addr, uint32_t acc_flag) {
        const uint64_t b_desc = ptx::umma_desc_sw128(b_addr);
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
          const uint64_t a_desc = ptx::umma_desc_sw128(a_addr + mi * A_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            // advance 16 fp16 = 32 bytes inside the swizzle row: +2 in the 16-byte-granular address field
            if (CG == 2) ptx::umma_f16_2cta(d_tmem + mi * BLOCK_N, a_desc + 2 * k, b_desc + 2 * k, idesc, acc_flag | k);
            else ptx::umma_f16(d_tmem + mi * BLOCK_N, a_desc + 2 * k, b_desc + 2 * k, idesc, acc_flag | k);
          }
        }
      };
      for (int sw = 0; sw < p.nseg * p.wsplit; ++sw) {
        const int s = sw % p.nseg;
        const bool halo_seg = HALO && p.taps[s] == 9;
        const int slots = halo_seg ? 3 * p.kch[s] : p.taps[s] * p.kch[s];
        for (int g = 0; g < slots; ++g) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          GDDIM_STAMP(p, sw == 0 && g == 0, mma_seq, 15);
          const uint32_t slot = ptx::smem_u32(smem + stage * stage_bytes);
          if (halo_seg) {
            // y-shift dyi = rows [dyi * W, dyi * W + 128 * MT) of the halo box
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) {
              issue(slot + dyi * p.W * 128, slot + a_bytes + dyi * L::B_TILE_BYTES, accum);
              accum = 1;
            }
          } else {
            issue(slot, slot + a_bytes, accum);
            accum = 1;
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if (CG == 2) ptx::umma_commit_2cta(&empty_bar[stage], 3); else ptx::umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      // accumulator ready for the epilogue (of both CTAs)
      GDDIM_STAMP(p, true, mma_seq, 10);
      if (CG == 2) ptx::umma_commit_2cta(&tfull_bar[acc], 3); else ptx::umma_commit(&tfull_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp < (EPI == EPI_SOFTMAX ? 4 : EPI_WARPS)) {
    // ================= epilogue =================
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
    const int group = warp >> 2;
    const int row = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    int bias_buf = 0;
    int tile_seq = 0;
    for (int tile_i = tile0; tile_i < num_tiles; tile_i += tile_step, bias_buf ^= 1, ++tile_seq) {
      const int tile = p.reverse ? num_tiles - 1 - tile_i : tile_i;
      const int mt = (tile / p.n_tiles) * CG + cta_rank, nt = tile % p.n_tiles;
      const long long m = (long long)mt * MT * BLOCK_M + row;
      const bool valid = m < p.M;
      uint32_t r[32];
      bool tmem_released = false;
      if (GNF) {
        constexpr int RS = L::EPI_ROW_FLOATS;
        float* stg = epi_stage + warp * 32 * RS;
        float* bias_s = epi_bias + bias_buf * BLOCK_N;
        float* gb = reinterpret_cast<float*>(gnf_smem + GnfSmem<BLOCK_N>::OFF_GB) + bias_buf * 2 * BLOCK_N;
        for (int j = threadIdx.x; j < BLOCK_N; j += EPI_WARPS * 32) {
          float b = 0.f;
          if (p.bias != nullptr) b += __ldg(p.bias + nt * BLOCK_N + j);
          if (p.bias2 != nullptr) b += __ldg(p.bias2 + nt * BLOCK_N + j);
          bias_s[j] = p.out32 != nullptr ? b : b * p.scale;      // (dual mode: the linear epilogue of pass 1 scales it itself)
          gb[j] = __ldg(p.gn_gamma + nt * BLOCK_N + j);
          gb[BLOCK_N + j] = __ldg(p.gn_beta + nt * BLOCK_N + j);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        EpiCtx<BLOCK_N, MT> cx{p, stg, bias_s, &tfull_bar[acc], acc_phase, tmem_base + (uint32_t(quad * 32) << 16) + acc * MT * BLOCK_N,
                               (long long)mt * MT * BLOCK_M + quad * 32, nt * BLOCK_N, lane, group};
        GnfCtx gx{reinterpret_cast<float2*>(gnf_smem), reinterpret_cast<float2*>(gnf_smem + GnfSmem<BLOCK_N>::OFF_GSTAT),
                  reinterpret_cast<float2*>(gnf_smem + GnfSmem<BLOCK_N>::OFF_XCHG), gb, gb + BLOCK_N,
                  reinterpret_cast<uint64_t*>(gnf_smem + GnfSmem<BLOCK_N>::OFF_BAR), (uint32_t)(tile_seq & 1),
                  (uint32_t)((tile_seq >> 1) & 1), clustered ? ptx::cluster_ctarank() : 0u, warp, tile_seq};
        if (p.out32 != nullptr) {
          // dual mode: pass 1 = the plain linear epilogue (fp32 result + column statistics) with the group-partials hook;
          // the accumulator is free afterwards, pass 2 re-reads the fp32 result
          const bool full = ((long long)(mt + 1) * MT * BLOCK_M <= (long long)p.M);
          if (!full) epi_tile<BLOCK_N, MT, true, true, false, true, false, false, true, false, true>(cx, gx.pstat, quad);
          else if (p.residual != nullptr) epi_tile<BLOCK_N, MT, true, true, false, true, false, true, false, false, true>(cx, gx.pstat, quad);
          else epi_tile<BLOCK_N, MT, false, true, false, true, false, true, false, false, true>(cx, gx.pstat, quad);
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tempty_bar[acc]), 0));
            else ptx::mbar_arrive(&tempty_bar[acc]);
          }
          tmem_released = true;
          gnf_fold<BLOCK_N, MT>(p, gx, lane);
          if (full) epi_tile_gnf_dual_pass2<BLOCK_N, MT, true>(cx, gx); else epi_tile_gnf_dual_pass2<BLOCK_N, MT, false>(cx, gx);
        } else if ((long long)(mt + 1) * MT * BLOCK_M <= (long long)p.M) {
          epi_tile_gnf<BLOCK_N, MT, true>(cx, gx);
        } else {
          epi_tile_gnf<BLOCK_N, MT, false>(cx, gx);
        }
      } else if (EPI == EPI_LINEAR) {
        // TMEM -> registers (thread = row) -> padded smem -> registers (8 lanes = one 32-column row segment),
        // so that every global access of the epilogue is a full 128-byte line.  Everything that does not depend
        // on the accumulator (bias row, first residual chunk) is fetched before waiting for the MMAs, and the
        // residual of chunk q+1 is in flight while chunk q is processed: the epilogue warps have no peers to
        // hide latency behind (one warp per scheduler), so the overlap has to be explicit.
        constexpr int RS = L::EPI_ROW_FLOATS;
        float* stg = epi_stage + warp * 32 * RS;
        // one (bias + bias2) row per tile, filled by the 256 epilogue threads together; double-buffered so that a
        // warp already on the next tile never overwrites the row a slower warp still reads (the barrier below keeps
        // them within one tile of each other)
        float* bias_s = epi_bias + bias_buf * BLOCK_N;
        for (int j = threadIdx.x * 4; j < BLOCK_N; j += EPI_WARPS * 32 * 4) {
          float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias != nullptr) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + nt * BLOCK_N + j));
            bsum.x += t.x; bsum.y += t.y; bsum.z += t.z; bsum.w += t.w;
          }
          if (p.bias2 != nullptr) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias2 + nt * BLOCK_N + j));
            bsum.x += t.x; bsum.y += t.y; bsum.z += t.z; bsum.w += t.w;
          }
          *reinterpret_cast<float4*>(bias_s + j) = bsum;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        const bool full = ((long long)(mt + 1) * MT * BLOCK_M <= (long long)p.M);
        const unsigned mode = (p.residual ? 1u : 0u) | (p.out32 ? 2u : 0u) | (p.out16 ? 4u : 0u) |
                              (p.colstats ? 8u : 0u) | (p.rowscale ? 16u : 0u);
        EpiCtx<BLOCK_N, MT> cx{p, stg, bias_s, &tfull_bar[acc], acc_phase, tmem_base + (uint32_t(quad * 32) << 16) + acc * MT * BLOCK_N,
                               (long long)mt * MT * BLOCK_M + quad * 32, nt * BLOCK_N, lane, group, tile_seq, warp};
        if (p.n_store > 0) {                                                                                    // few-channel output
          bool done = false;
          if constexpr (BLOCK_N == 32 && MT == 1 && CG == 1 && !HALO) {
            if (p.upd_on) { epi_head_update<BLOCK_N, MT>(cx); done = true; }      // head + gDDIM update
          }
          if (!done) epi_tile<BLOCK_N, MT, false, true, false, false, false, false, false, true>(cx);
        }
        else if (!full) epi_tile<BLOCK_N, MT, true, true, true, true, true, false, true>(cx);     // ragged last tile: generic path
        else if (mode == (2u | 8u)) epi_tile<BLOCK_N, MT, false, true, false, true, false, true, false>(cx);           // conv1, shortcut conv2, stem
        else if (mode == (1u | 2u | 8u)) epi_tile<BLOCK_N, MT, true, true, false, true, false, true, false>(cx);       // conv2 / proj / pyramid
        else if (mode == 4u) epi_tile<BLOCK_N, MT, false, false, true, false, false, true, false>(cx);                  // qkv
        else if (mode == (4u | 16u)) epi_tile<BLOCK_N, MT, false, false, true, false, true, true, false>(cx);           // P.V
        else epi_tile<BLOCK_N, MT, true, true, true, true, true, true, true>(cx);
      } else {
        // row softmax over the BLOCK_N columns of this tile (requires N == BLOCK_N, MT == 1)
        ptx::mbar_wait(&tfull_bar[acc], acc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * BLOCK_N;
        const float sc = p.scale * 1.4426950408889634f;     // exp(x) = exp2(x * log2 e)
        float mx = -INFINITY;
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
          ptx::tmem_ld_32x32b_x32(taddr + c0, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]) * sc);
        }
        float sum = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
          ptx::tmem_ld_32x32b_x32(taddr + c0, r);
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            // round to fp16 first so that the row sum matches what the PV GEMM will actually consume
            v[j] = __half2float(__float2half_rn(exp2f(__uint_as_float(r[j]) * sc - mx)));
            sum += v[j];
          }
          if (valid) {
            uint4* op = reinterpret_cast<uint4*>(p.out16 + m * p.ldo + c0);
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              __half2 h0 = __floats2half2_rn(v[j], v[j + 1]);
              __half2 h1 = __floats2half2_rn(v[j + 2], v[j + 3]);
              __half2 h2 = __floats2half2_rn(v[j + 4], v[j + 5]);
              __half2 h3 = __floats2half2_rn(v[j + 6], v[j + 7]);
              uint4 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&h0);
              pk.y = *reinterpret_cast<uint32_t*>(&h1);
              pk.z = *reinterpret_cast<uint32_t*>(&h2);
              pk.w = *reinterpret_cast<uint32_t*>(&h3);
              op[j >> 3] = pk;
            }
          }
        }
        if (valid) p.row_out[m] = 1.0f / sum;
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0 && !tmem_released) {
        if (CG == 2) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tempty_bar[acc]), 0));   // the leader's barrier
        else ptx::mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  if (clustered) ptx::cluster_sync(); else __syncthreads();     // pair / cluster: no CTA may retire while a peer still works
  GDDIM_STAMP(p, threadIdx.x == 0, 0, 13);
  if (warp == MMA_WARP) {
    ptx::tc_fence_after();
    if (CG == 2) ptx::tmem_dealloc_2cta(tmem_base, TMEM_COLS); else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// CUDA-core reference with identical semantics (validation of the tcgen05 path on the GPU; never the
// default).  64x64 output tile, 16-wide K slices, 256 threads x (4x4) outputs.
// ---------------------------------------------------------------------------------------------------
struct RefArgs {
  const __half* a[2];
  int ctot[2], coff[2], c[2], taps[2];
  int nseg;
  int B, H, W, M, N;
  const __half* w;
  int w_ld, w_koff;
  int w_lo_off;                  // > 0: second weight half (W = W_hi + W_lo) this many columns further
  long long w_batch_stride;
  int tiles_per_batch_rows;      // rows of M per batch matrix (0 = shared)
  const float* bias;
  const float* bias2;
  const float* residual;
  const float* rowscale;
  float scale;
  float* out32;
  __half* out16;
  float* row_out;
  int ldo;
  int epi;
  float* softmax_tmp;            // [M, N] scratch for the softmax epilogue
  float* colstats;
  int n_store;
};

__global__ void __launch_bounds__(256) conv_gemm_ref_kernel(const RefArgs p) {
  pdl_entry();
  __shared__ float sA[16][64 + 1];
  __shared__ float sW[16][64 + 1];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const long long m0 = (long long)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  int kglobal = 0;
  for (int s = 0; s < p.nseg; ++s) {
    for (int tap = 0; tap < p.taps[s]; ++tap) {
      const int dy = (p.taps[s] == 9) ? tap / 3 - 1 : 0;
      const int dx = (p.taps[s] == 9) ? tap % 3 - 1 : 0;
      for (int c0 = 0; c0 < p.c[s]; c0 += 16, kglobal += 16) {
        // load A slice: 64 rows x 16 k
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
          const int r = i / 16, kk = i % 16;
          const long long m = m0 + r;
          float v = 0.f;
          if (m < p.M) {
            const int x = int(m % p.W), y = int((m / p.W) % p.H);
            const long long b = m / ((long long)p.W * p.H);
            const int yy = y + dy, xx = x + dx;
            if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W)
              v = __half2float(p.a[s][((b * p.H + yy) * p.W + xx) * p.ctot[s] + p.coff[s] + c0 + kk]);
          }
          sA[kk][r] = v;
        }
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
          const int r = i / 16, kk = i % 16;
          const int n = n0 + r;
          float v = 0.f;
          if (n < p.N) {
            long long boff = 0;
            if (p.w_batch_stride != 0) boff = (m0 / p.tiles_per_batch_rows) * p.w_batch_stride;
            v = __half2float(p.w[boff + (long long)n * p.w_ld + p.w_koff + kglobal + kk]);
            if (p.w_lo_off > 0) v += __half2float(p.w[boff + (long long)n * p.w_ld + p.w_koff + p.w_lo_off + kglobal + kk]);
          }
          sW[kk][r] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          float a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = sW[kk][tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N || (p.n_store > 0 && n >= p.n_store)) continue;
      float v = acc[i][j];
      if (p.epi == EPI_SOFTMAX) {
        p.softmax_tmp[m * p.N + n] = v * p.scale;
        continue;
      }
      if (p.rowscale) v *= p.rowscale[m];
      if (p.bias) v += p.bias[n];
      if (p.bias2) v += p.bias2[n];
      if (p.residual) v += p.residual[m * p.ldo + n];
      v *= p.scale;
      if (p.out32) p.out32[m * p.ldo + n] = v;
      if (p.out16) p.out16[m * p.ldo + n] = __float2half_rn(v);
    }
  }
}

// column statistics of the reference path's fp32 output (same layout as the fused epilogue's)
__global__ void colstats_ref_kernel(const float* out32, float* colstats, int M, int N, int ldo) {
  pdl_entry();
  const int n = blockIdx.y * blockDim.x + threadIdx.x;
  const long long slab = blockIdx.x;
  if (n >= N) return;
  float s = 0.f, q = 0.f;
  for (int r = 0; r < 32; ++r) {
    const long long m = slab * 32 + r;
    if (m >= M) break;
    const float v = out32[m * ldo + n];
    s += v; q += v * v;
  }
  colstats[(slab * 2) * ldo + n] = s;
  colstats[(slab * 2 + 1) * ldo + n] = q;
}

__global__ void softmax_ref_kernel(const float* s, __half* out16, float* row_out, int M, int N, int ldo) {
  pdl_entry();
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float mx = -INFINITY;
  for (int n = 0; n < N; ++n) mx = fmaxf(mx, s[m * N + n]);
  float sum = 0.f;
  for (int n = 0; n < N; ++n) {
    const __half e = __float2half_rn(expf(s[m * N + n] - mx));
    sum += __half2float(e);
    out16[m * ldo + n] = e;
  }
  row_out[m] = 1.0f / sum;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  });
  return fn;
}

static int encode_4d(CUtensorMap* tm, const void* base, const uint64_t dims[4], const uint32_t box[4]) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) GEMM_FAIL("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstr[3] = {dims[0] * 2, dims[0] * dims[1] * 2, dims[0] * dims[1] * dims[2] * 2};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    GEMM_FAIL("cuTensorMapEncodeTiled failed (%d) dims=%llu,%llu,%llu,%llu box=%u,%u,%u,%u", (int)r,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
              (unsigned long long)dims[3], box[0], box[1], box[2], box[3]);
  return 0;
}

int tmap_encode_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint32_t* box) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) GEMM_FAIL("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  if (rank < 2 || rank > 4) GEMM_FAIL("tmap_encode_f16: rank %d", rank);
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  cuuint64_t stride = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i]; bx[i] = box[i];
    stride *= dims[i];
    if (i < rank - 1) gstr[i] = stride;
  }
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) GEMM_FAIL("cuTensorMapEncodeTiled failed (%d), rank %d", (int)r, rank);
  return 0;
}

static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// EPI_GNF tile rule: with a CTA tile of `tr` rows and `bn` columns, an image of `rpi` rows must either fit the tile a
// whole number of times (>= 16 rows each, <= GNF_IMGS images) or span 2 / 4 CTAs of one cluster, which then must own all
// N columns (bn == N); a group's channels never straddle N tiles and a tile holds at most GNF_GMAX groups.
static int gnf_cluster(int rpi, int tr, int bn, int N, int cpg) {
  if (bn % cpg != 0 || bn / cpg > GNF_GMAX) return 0;
  if (rpi >= tr) {
    const int xc = rpi / tr;
    if (xc == 1) return 1;
    return ((xc == 2 || xc == 4) && bn == N) ? xc : 0;
  }
  if (rpi < 16 || tr % rpi != 0 || tr / rpi > GNF_IMGS) return 0;
  return 1;
}

int gemm_gnf_supported(int H, int W, int N, int groups) {
  if (groups <= 0 || N % groups != 0 || N % 32 != 0 || !is_pow2(H) || !is_pow2(W)) return 0;
  const int cpg = N / groups, rpi = H * W;
  if (cpg != 4 && cpg != 8 && cpg != 16) return 0;
  for (int bn = 256; bn >= 32; bn >>= 1) {
    if (N % bn != 0) continue;
    for (int ms = 1; ms <= 2; ++ms) {
      if (ms == 2 && bn != 128 && bn != 64) continue;
      if (gnf_cluster(rpi, BLOCK_M * ms, bn, N, cpg)) return 1;
    }
  }
  return 0;
}

int gemm_prepare(GemmOp* op, int force_block_n, int force_m_sub, int force_cg) {
  op->prepared = 0;
  op->m_sub = 1;
  const long long M = (long long)op->B * op->H * op->W;
  if (!is_pow2(op->W) || !is_pow2(op->H)) GEMM_FAIL("conv_gemm: H and W must be powers of two (got %d x %d)", op->H, op->W);
  if (op->nseg < 1 || op->nseg > 2) GEMM_FAIL("conv_gemm: 1 or 2 A segments");
  int ktot = 0;
  op->cuda_core = 0;
  for (int s = 0; s < op->nseg; ++s) {
    const GemmSeg& g = op->seg[s];
    if (g.c % 16 != 0 || g.c_off % 8 != 0 || g.c_total % 8 != 0)
      GEMM_FAIL("conv_gemm: segment %d channels (%d of %d at %d) must be a multiple of 16", s, g.c, g.c_total, g.c_off);
    if (g.c % BLOCK_K != 0) op->cuda_core = 1;       // no 64-wide K block to feed the UMMA (the network planner pairs pixels instead)
    if (g.taps != 1 && g.taps != 9) GEMM_FAIL("conv_gemm: taps must be 1 or 9");
    ktot += g.taps * g.c;
  }
  if (op->cuda_core) {
    // CUDA-core kernel (conv_gemm_ref_kernel): any N, linear epilogue with every option but the fused GroupNorm / softmax
    if (op->epi != EPI_LINEAR) GEMM_FAIL("conv_gemm: channel counts that are not multiples of %d support the linear epilogue only", BLOCK_K);
    if (op->w_koff + ktot * (op->wsplit == 2 ? 2 : 1) > op->w_ld) GEMM_FAIL("conv_gemm: K range exceeds weight row stride");
    if (op->w_batch_stride != 0) GEMM_FAIL("conv_gemm: batched B operand needs channel counts that are multiples of %d", BLOCK_K);
    op->block_n = 0; op->cg = 1; op->halo = 0; op->gn_xc = 0; op->m_tiles = op->n_tiles = op->tiles_per_batch = 0;
    op->prepared = 1;
    return 0;
  }
  if (op->wsplit != 0 && op->wsplit != 1 && op->wsplit != 2) GEMM_FAIL("conv_gemm: wsplit must be 1 or 2");
  if (op->w_koff + ktot * (op->wsplit == 2 ? 2 : 1) > op->w_ld) GEMM_FAIL("conv_gemm: K range exceeds weight row stride");
  if (op->wsplit == 2 && (op->w_batch_stride != 0 || op->epi == EPI_SOFTMAX)) GEMM_FAIL("conv_gemm: split weights need a shared weight matrix");
  const bool gnf = op->epi == EPI_GNF;
  const int rpi = op->H * op->W;
  int cpg = 0;
  if (gnf) {
    if (!op->gn_gamma || !op->gn_beta || !op->out16 || op->rowscale || op->n_store != 0 || op->w_batch_stride != 0)
      GEMM_FAIL("conv_gemm: the GroupNorm epilogue needs gamma / beta / out16 and excludes rowscale / n_store / batched B");
    if (!op->out32 && (op->colstats || op->residual))
      GEMM_FAIL("conv_gemm: GroupNorm epilogue: colstats / residual belong to the dual mode (out32 + out16)");
    if (!gemm_gnf_supported(op->H, op->W, op->N, op->gn_groups))
      GEMM_FAIL("conv_gemm: GroupNorm epilogue not available for %dx%d, N=%d, groups=%d", op->H, op->W, op->N, op->gn_groups);
    cpg = op->N / op->gn_groups;
  }
  int bn = force_block_n;
  if (bn == 0) {
    if (op->epi == EPI_SOFTMAX) bn = op->N;
    else if (op->N % 256 == 0) bn = 256;
    else if (op->N % 128 == 0) bn = 128;
    else if (op->N % 64 == 0) bn = 64;
    else if (op->N % 32 == 0) bn = 32;
    else GEMM_FAIL("conv_gemm: N=%d must be a multiple of 32", op->N);
    // Pick the widest tile the layer can feed: wide tiles halve the L2->smem traffic per FLOP (the A tile is
    // re-read once per N tile) and are MMA-issue efficient; narrow tiles only win when the layer has too few
    // tiles to occupy the 148 SMs.  Cost model in tensor-pipe cycles per CTA: waves x (mainloop + epilogue).
    if (op->epi != EPI_SOFTMAX) {
      const long long mt = (M + BLOCK_M - 1) / BLOCK_M;
      const int kb = ktot / BLOCK_K;
      double best = 1e30;
      int best_bn = bn, best_ms = 1;
      for (int cand = bn; cand >= 32; cand >>= 1) {
        if (op->N % cand != 0) continue;
        for (int ms = 1; ms <= 2; ++ms) {
          if (ms == 2 && (cand > 128 || cand < 64 || op->w_batch_stride != 0)) continue;
          if (gnf && !gnf_cluster(rpi, BLOCK_M * ms, cand, op->N, cpg)) continue;
          const long long tiles = ((mt + ms - 1) / ms) * (op->N / cand);
          const long long waves = (tiles + 147) / 148;
          const double width = (double)cand * ms;         // output columns x row-tiles sharing one operand load
          const double ineff = width >= 256 ? 1.0 : (width >= 128 ? 1.25 : 1.5);
          const double mma = (double)kb * 4.0 * ms * (cand >= 64 ? cand / 2.0 : 32.0) * ineff;
          const double cost = (double)waves * (mma + 6.0 * cand * ms + 1500.0);
          if (cost < best * 0.97) { best = cost; best_bn = cand; best_ms = ms; }
        }
      }
      op->m_sub = best_ms;
      bn = best_bn;
    }
  }
  if (op->N % bn != 0) GEMM_FAIL("conv_gemm: N=%d not a multiple of block_n=%d", op->N, bn);
  if (op->epi == EPI_SOFTMAX && (bn != op->N || bn != 256)) GEMM_FAIL("conv_gemm: softmax epilogue needs N == 256");
  if (op->epi == EPI_SOFTMAX && (!op->out16 || !op->row_out)) GEMM_FAIL("conv_gemm: softmax needs out16 and row_out");
  if (op->n_store == 0 && op->ldo % 8 != 0) GEMM_FAIL("conv_gemm: ldo must be a multiple of 8");
  if (op->n_store != 0 && (op->residual || op->out16 || op->colstats || op->rowscale || !op->out32 || op->epi != EPI_LINEAR))
    GEMM_FAIL("conv_gemm: n_store supports bias + scale + fp32 output only");
  op->block_n = bn;
  if (force_block_n != 0) op->m_sub = (force_m_sub == 2) ? 2 : 1;
  op->gn_xc = 0;
  if (gnf) {
    op->gn_xc = gnf_cluster(rpi, BLOCK_M * op->m_sub, bn, op->N, cpg);
    if (op->gn_xc == 0) GEMM_FAIL("conv_gemm: GroupNorm epilogue: no valid tile for block_n %d, m_sub %d at %dx%d", bn, op->m_sub, op->H, op->W);
  }
  if (op->m_sub == 2 && ((bn != 128 && bn != 64) || op->w_batch_stride != 0 || op->epi == EPI_SOFTMAX))
    GEMM_FAIL("conv_gemm: 256-row tiles need block_n 64/128, shared weights and the linear epilogue");
  op->m_tiles = int((M + (long long)BLOCK_M * op->m_sub - 1) / ((long long)BLOCK_M * op->m_sub));
  op->n_tiles = op->N / bn;
  // ---- CTA pairs (cta_group::2) and halo tiles ------------------------------------------------------------------
  {
    static int no_pairs = -1, no_halo = -1, halo128_cg = -1;   // A/B timing switches, not product options
    if (no_pairs < 0) { const char* e = getenv("GDDIM_NO_CTA_PAIRS"); no_pairs = (e && e[0] == '1') ? 1 : 0; }
    if (no_halo < 0) { const char* e = getenv("GDDIM_NO_HALO"); no_halo = (e && e[0] == '1') ? 1 : 0; }
    if (halo128_cg < 0) { const char* e = getenv("GDDIM_HALO128_CG"); halo128_cg = e ? atoi(e) : 2; }
    static int halo256 = -1;               // GDDIM_HALO256=1: halo tiles for N = 256 too (measured slower: two ring slots)
    if (halo256 < 0) { const char* e = getenv("GDDIM_HALO256"); halo256 = (e && e[0] == '1') ? 1 : 0; }
    const bool plain = op->epi != EPI_SOFTMAX && op->w_batch_stride == 0 && op->n_store == 0;
    // (GroupNorm epilogue: a pair must be exactly one image; images spanning 4 CTAs run as clusters of single-CTA MMAs)
    const bool can_pair = plain && bn == 256 && op->m_sub == 1 && (!gnf || op->gn_xc == 2);
    // images spanning four CTAs: two cta_group::2 pairs with the statistics exchanged through global memory (all SMs
    // busy) when the halo pair kernel applies, else a cluster of four single-CTA MMAs (DSMEM exchange, 132 SMs)
    static int gnf_pairs4 = -1;              // GDDIM_GNF_PAIRS4=0: A/B switch back to 4-CTA clusters
    if (gnf_pairs4 < 0) { const char* e = getenv("GDDIM_GNF_PAIRS4"); gnf_pairs4 = (e && e[0] == '0') ? 0 : 1; }
    // halo tiles: 3x3 convolution whose CTA tile is a block of whole image rows of ONE image
    const int tile_px = BLOCK_M * op->m_sub;
    // (measured: 256-row N = 128 tiles 1.06 -> 1.29 PFLOP/s; N = 256 tiles lose, their ring shrinks to two slots)
    const bool halo_ok = !no_halo && plain && ((bn == 128 && op->m_sub == 2) || (bn == 256 && op->m_sub == 1 && halo256)) &&
                         op->seg[0].taps == 9 && (op->nseg == 1 || op->seg[1].taps == 1) && op->W >= 16 && op->W <= 128 &&
                         tile_px % op->W == 0 && op->H % (tile_px / op->W) == 0;
    // pairs, measured (profiles/): +8..11 % on the K >= 1152, N = 256 layers with at least two waves of tiles, a loss on
    // K = 256 GEMMs and single-wave layers, which therefore stay on single-CTA MMAs
    const bool shape_ok = ktot / BLOCK_K >= 16 && op->m_tiles >= 256;
    op->cg = (!no_pairs && shape_ok && can_pair) ? 2 : 1;
    if (halo_ok && bn == 256 && op->m_tiles >= 2 && !no_pairs) op->cg = 2;   // three 32 KB weight tiles per slot do not fit
    if (halo_ok && bn == 128 && op->m_tiles >= 2 && !no_pairs) op->cg = halo128_cg == 2 ? 2 : 1;
    if (gnf && op->gn_xc != 2 && !(op->gn_xc == 4 && gnf_pairs4 && halo_ok && bn == 128 && op->cg == 2)) op->cg = 1;
    if (force_cg == 1) op->cg = 1;
    if (force_cg == 2) {
      if ((!can_pair && !halo_ok) || (gnf && op->gn_xc != 2 && !(op->gn_xc == 4 && halo_ok && bn == 128)))
        GEMM_FAIL("conv_gemm: CTA pairs need block_n 256 (or halo tiles), shared weights, linear epilogue");
      op->cg = 2;
    }
    op->halo = 0;
    if (halo_ok && !(bn == 256 && op->cg == 1)) {
      using L0 = SmemLayout<32, 1, 1>;      // BAR / epilogue staging sizes do not depend on the tile shape ...
      const int epi_bytes = L0::EPI_STAGE_BYTES + 2 * bn * 4 +                // ... except for the bias rows
                            (gnf ? (bn == 128 ? GnfSmem<128>::BYTES : GnfSmem<256>::BYTES) : 0);
      const int b_tile = (bn / op->cg) * BLOCK_K * 2;
      op->a_bytes = (tile_px + 2 * op->W) * BLOCK_K * 2;
      op->stage_bytes = op->a_bytes + 3 * b_tile;
      int st = (SMEM_BUDGET - L0::BAR_BYTES - epi_bytes) / op->stage_bytes;
      if (st > 8) st = 8;
      if (st >= 2) { op->halo = 1; op->stages = st; }
    }
    if (op->cg == 2 && !op->halo && !can_pair) op->cg = 1;
  }
  op->tiles_per_batch = 0;
  if (op->w_batch_stride != 0) {
    const int hw = op->H * op->W;
    if (hw % BLOCK_M != 0) GEMM_FAIL("conv_gemm: batched B operand needs H*W %% 128 == 0");
    op->tiles_per_batch = hw / BLOCK_M;
  }
  // A boxes: 128 consecutive pixels in NHW order
  const uint32_t bw = op->W < BLOCK_M ? op->W : BLOCK_M;
  const uint32_t bh = (op->H < int(BLOCK_M / bw)) ? op->H : BLOCK_M / bw;
  const uint32_t bb = BLOCK_M / (bw * bh);
  for (int s = 0; s < op->nseg; ++s) {
    const GemmSeg& g = op->seg[s];
    const uint64_t dims[4] = {(uint64_t)g.c_total, (uint64_t)op->W, (uint64_t)op->H, (uint64_t)op->B};
    const uint32_t box[4] = {BLOCK_K, bw, bh, bb};
    if (encode_4d(&op->tmA[s], g.ptr, dims, box)) return -1;
  }
  if (op->nseg == 1) op->tmA[1] = op->tmA[0];
  op->tmH = op->tmA[0];
  if (op->halo) {
    const GemmSeg& g = op->seg[0];
    const uint64_t dims[4] = {(uint64_t)g.c_total, (uint64_t)op->W, (uint64_t)op->H, (uint64_t)op->B};
    const uint32_t box[4] = {BLOCK_K, (uint32_t)op->W, (uint32_t)(BLOCK_M * op->m_sub / op->W + 2), 1};
    if (encode_4d(&op->tmH, g.ptr, dims, box)) return -1;
  }
  {
    uint64_t nb = 1, rows = op->N;
    if (op->w_batch_stride != 0) {
      rows = op->w_rows_per_batch;
      if ((long long)rows * op->w_ld != op->w_batch_stride) GEMM_FAIL("conv_gemm: batched B operand must be dense");
      nb = op->B;
    }
    const uint64_t dims[4] = {(uint64_t)op->w_ld, rows, nb, 1};
    const uint32_t box[4] = {BLOCK_K, (uint32_t)(bn / op->cg), 1, 1};      // a CTA of a pair stages half the weight tile
    if (encode_4d(&op->tmB, op->w, dims, box)) return -1;
  }
  op->prepared = 1;
  return 0;
}

template <int BN, int EPI, int MT, int CG = 1, bool HALO = false>
static int launch_umma(const GemmOp* op, const GemmArgs& a, cudaStream_t st) {
  using L = SmemLayout<BN, MT, CG, EPI == EPI_GNF ? GnfSmem<BN>::BYTES : 0>;
  static DeviceOnce attr_set;
  auto kern = conv_gemm_umma_kernel<BN, EPI, MT, CG, HALO>;
  const int smem_total = HALO ? op->stages * op->stage_bytes + L::BAR_BYTES + L::EPI_BYTES : L::TOTAL;
  if (attr_set.need()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO ? SMEM_BUDGET : L::TOTAL);
    if (e != cudaSuccess) GEMM_FAIL("cudaFuncSetAttribute(smem=%d): %s", L::TOTAL, cudaGetErrorString(e));
    attr_set.done();
  }
  const int num_sms = device_sm_count();
  int num_sms_eff = num_sms;
  {
    static int cap = -1;                      // GDDIM_GEMM_MAX_CTAS: SM-partitioning experiments only
    if (cap < 0) { const char* e = getenv("GDDIM_GEMM_MAX_CTAS"); cap = e ? atoi(e) : 0; }
    if (cap > 0 && cap < num_sms) num_sms_eff = cap & ~1;
  }
  if (CG == 2) {
    const int units = ((a.m_tiles + 1) / 2) * a.n_tiles;
    int pairs = units < num_sms_eff / 2 ? units : num_sms_eff / 2;
    if (EPI == EPI_GNF && a.gn_xg != nullptr) {
      // the two pairs of an image must work on it in the same round: an even number of pairs, whole images
      if (units % 2 != 0) GEMM_FAIL("conv_gemm: GroupNorm epilogue needs whole images (%d pair tiles)", units);
      pairs &= ~1;
      if (pairs < 2) GEMM_FAIL("conv_gemm: GroupNorm epilogue over two pairs needs at least four SMs");
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = smem_total; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_attr(at, 1);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, op->tmA[0], op->tmA[1], op->tmB, op->tmH, a);
    if (e != cudaSuccess) GEMM_FAIL("conv_gemm_umma pair launch: %s", cudaGetErrorString(e));
    return 0;
  }
  const int tiles = a.m_tiles * a.n_tiles;
  if (EPI == EPI_GNF && a.gn_xc > 1) {
    // images span gn_xc CTAs: clusters of gn_xc single-CTA MMAs, CTA r of a cluster owns tile (k * gn_xc + r) = part r of
    // image k.  The grid is what can be co-resident (a persistent cluster that had to wait for a free GPC slot would
    // run its whole share after everyone else): cudaOccupancyMaxActiveClusters, cached per device.
    const int xc = a.gn_xc;
    if (tiles % xc != 0) GEMM_FAIL("conv_gemm: GroupNorm epilogue needs whole images (tiles %d, cluster %d)", tiles, xc);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = smem_total; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = xc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    static int max_clusters[64][5] = {};
    const int dev = DeviceOnce::dev() & 63;
    if (max_clusters[dev][xc] == 0) {
      cfg.gridDim = dim3(num_sms / xc * xc); cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = num_sms / xc / 2 > 0 ? num_sms / xc / 2 : 1; }
      max_clusters[dev][xc] = n;
      if (getenv("GDDIM_VERBOSE")) fprintf(stderr, "gddim: conv_gemm<%d,GNF,%d,%d,%d>: %d co-resident clusters of %d CTAs (%d SMs)\n", BN, MT, CG, (int)HALO, n, xc, num_sms);
    }
    int clusters = tiles / xc;
    if (clusters > max_clusters[dev][xc]) clusters = max_clusters[dev][xc];
    if (clusters > num_sms_eff / xc) clusters = num_sms_eff / xc;
    cfg.gridDim = dim3(clusters * xc);
    cfg.numAttrs = pdl_attr(at, 1);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, op->tmA[0], op->tmA[1], op->tmB, op->tmH, a);
    if (e != cudaSuccess) GEMM_FAIL("conv_gemm_umma cluster launch (x%d): %s", xc, cudaGetErrorString(e));
    return 0;
  }
  const int grid = tiles < num_sms_eff ? tiles : num_sms_eff;
  launch_k(kern, dim3(grid), dim3(NUM_THREADS), smem_total, st, op->tmA[0], op->tmA[1], op->tmB, op->tmH, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) GEMM_FAIL("conv_gemm_umma launch: %s", cudaGetErrorString(e));
  return 0;
}

static float* g_softmax_tmp = nullptr;
static size_t g_softmax_tmp_bytes = 0;

// reference path of EPI_GNF: GroupNorm (+ swish) of the fp32 GEMM result, one block per (image, group)
__global__ void __launch_bounds__(256) gnf_ref_kernel(const float* __restrict__ v, __half* __restrict__ out16, int rpi, int N,
                                                      int cpg, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float eps, int silu, int ldo) {
  pdl_entry();
  const long long row0 = (long long)blockIdx.x * rpi;
  const int c0 = blockIdx.y * cpg;
  const int n = rpi * cpg;
  __shared__ double sh[2][256];
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float x = v[(row0 + i / cpg) * N + c0 + i % cpg];
    s += x; q += (double)x * x;
  }
  sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
    __syncthreads();
  }
  const float mean = (float)(sh[0][0] / n);
  const float var = fmaxf((float)(sh[1][0] / n - (sh[0][0] / n) * (sh[0][0] / n)), 0.f);
  const float rstd = rsqrtf(var + eps);
  for (int i = threadIdx.x; i < n; i += 256) {
    const int c = c0 + i % cpg;
    const long long m = row0 + i / cpg;
    float y = (v[m * N + c] - mean) * rstd * gamma[c] + beta[c];
    if (silu) y = y / (1.0f + expf(-y));
    out16[m * ldo + c] = __float2half_rn(y);
  }
}

// the epilogue of this (prepared) launch can apply the CLD update `u`: the six-column head convolution on the 128 x 32 tcgen05
// tile, three data channels, no noise term, at most four eps terms, eps_0 = this launch's own output
int gemm_head_update_supported(const GemmOp* op, const CldStepArgs* u) {
  return op->prepared && !op->cuda_core && op->epi == EPI_LINEAR && op->n_store == 6 && op->ldo == 6 && op->block_n == 32 &&
         op->m_sub == 1 && op->cg == 1 && !op->halo && op->out32 != nullptr && u != nullptr && u->C == 3 && u->noise_mode == 0 &&
         u->n_eps >= 1 && u->n_eps <= 4 && u->eps[0] == op->out32 && (!u->mixed || u->eps_store == op->out32) &&
         u->n_pix == (long long)op->B * op->H * op->W;
}

int gemm_launch(const GemmOp* op, int impl, cudaStream_t st) {
  const long long M = (long long)op->B * op->H * op->W;
  if (impl == 0 && !(op->prepared && op->cuda_core)) {
    if (!op->prepared) GEMM_FAIL("conv_gemm: op not prepared");
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    for (int s = 0; s < op->nseg; ++s) {
      a.taps[s] = op->seg[s].taps;
      a.kch[s] = op->seg[s].c / BLOCK_K;
      a.coff[s] = op->seg[s].c_off;
    }
    a.nseg = op->nseg;
    a.H = op->H; a.W = op->W;
    a.M = (int)M; a.N = op->N;
    a.m_tiles = op->m_tiles; a.n_tiles = op->n_tiles; a.tiles_per_batch = op->tiles_per_batch;
    a.w_koff = op->w_koff;
    a.wsplit = op->wsplit == 2 ? 2 : 1;
    a.ktot = 0;
    for (int s = 0; s < op->nseg; ++s) a.ktot += op->seg[s].taps * op->seg[s].c;
    a.bias = op->bias; a.bias2 = op->bias2; a.residual = op->residual; a.rowscale = op->rowscale;
    a.scale = op->scale; a.out32 = op->out32; a.out16 = op->out16; a.row_out = op->row_out; a.ldo = op->ldo;
    a.colstats = op->colstats;
    a.n_store = op->n_store;
    a.reverse = op->reverse;
    if (op->upd != nullptr) {
      const CldStepArgs& u = *op->upd;
      if (!gemm_head_update_supported(op, op->upd)) GEMM_FAIL("conv_gemm: head update: unsupported layer / step (see gemm_head_update_supported)");
      a.upd_on = 1;
      a.upd.u = u.u; a.upd.u_out = u.u_out;
      // the history loads are unconditional (registers), the sums are not: unused slots read rows of this launch's own A
      // operand -- valid (>= 128 bytes per pixel), never written by this launch (the loads are ld.global.nc)
      for (int j = 0; j < 4; ++j) a.upd.eps[j] = (j >= 1 && j < u.n_eps) ? u.eps[j] : reinterpret_cast<const float*>(op->seg[0].ptr);
      a.upd.n_eps = u.n_eps; a.upd.mixed = u.mixed;
      for (int j = 0; j < 5; ++j) for (int k = 0; k < 4; ++k) a.upd.coef[j][k] = u.coef[j][k];
      for (int k = 0; k < 4; ++k) a.upd.mixm[k] = u.mixm[k];
    }
    {
      static int wpf = -1;                    // GDDIM_NO_WPF=1: A/B switch for the weight prefetch
      if (wpf < 0) { const char* e = getenv("GDDIM_NO_WPF"); wpf = (e && e[0] == '1') ? 0 : 1; }
      if (wpf && op->w_batch_stride == 0) { a.pf_ptr = op->w; a.pf_bytes = (long long)op->N * op->w_ld * 2; }
    }
#ifdef GDDIM_ABLATE      // timing-only epilogue ablations (results INVALID): compile with -DGDDIM_ABLATE, never in the product build
    {
      static int dbg = -1;
      if (dbg < 0) { const char* e = getenv("GDDIM_GEMM_DBG"); dbg = e ? atoi(e) : 0; }
      a.dbg = dbg;
      static long long* clk = nullptr;
      if (getenv("GDDIM_CLK")) {
        if (!clk) cudaMalloc(&clk, 16 * 16 * sizeof(long long));
        cudaMemsetAsync(clk, 0, 16 * 16 * sizeof(long long), st);
        a.dbg_clk = clk;
      }
    }
    struct ClkDump {
      long long* clk; cudaStream_t st;
      ~ClkDump() {
        if (!clk) return;
        long long h[256];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
        long long t0 = h[11] ? h[11] : h[8];
        fprintf(stderr, "CLK timeline CTA 0 (cycles since kernel entry): prologue done %lld, kernel end %lld\n", h[12] - t0, h[13] - t0);
        for (int t = 0; t < 10; ++t) {
          if (h[t * 16 + 8] == 0) break;
          fprintf(stderr, " tile %d  mma: first operands %lld wait %lld..%lld issue_done %lld | epi: wait %lld..%lld p1 %lld bar %lld fold %lld bar %lld p2 %lld\n", t,
                  h[t * 16 + 15] - t0, h[t * 16 + 8] - t0, h[t * 16 + 9] - t0, h[t * 16 + 10] - t0, h[t * 16 + 0] - t0, h[t * 16 + 1] - t0,
                  h[t * 16 + 2] - t0, h[t * 16 + 3] - t0, h[t * 16 + 4] - t0, h[t * 16 + 5] - t0, h[t * 16 + 6] - t0);
        }
      }
    } clk_dump{a.dbg_clk, st};
#endif
    if (op->epi == EPI_SOFTMAX) return launch_umma<256, EPI_SOFTMAX, 1>(op, a, st);
    a.stages = op->stages; a.stage_bytes = op->stage_bytes; a.a_bytes = op->a_bytes;
    if (op->epi == EPI_GNF) {
      a.gn_gamma = op->gn_gamma; a.gn_beta = op->gn_beta; a.gn_eps = op->gn_eps; a.gn_cpg = op->N / op->gn_groups;
      a.gn_silu = op->gn_silu; a.gn_rpi = op->H * op->W; a.gn_xc = op->gn_xc;
      if (op->halo) {
        if (op->block_n == 128 && op->m_sub == 2 && op->cg == 2) {
          // two pairs per image, exchange through global memory: [4-CTA group][round parity][rank][group][2] tagged words
          static unsigned long long* xg[64] = {};
          const int dev = DeviceOnce::dev() & 63;
          if (!xg[dev]) {
            const size_t bytes = (size_t)(device_sm_count() / 4 + 1) * 2 * 4 * GNF_GMAX * 2 * sizeof(unsigned long long);
            if (cudaMalloc(&xg[dev], bytes) != cudaSuccess) GEMM_FAIL("conv_gemm: GroupNorm exchange buffer allocation failed");
            cudaMemset(xg[dev], 0, bytes);
          }
          a.gn_xg = xg[dev];
          // tags restart at 1 in every launch (a captured graph replays the same arguments): clear the words first
          const size_t used = (size_t)(device_sm_count() / 4 + 1) * 2 * 4 * GNF_GMAX * 2 * sizeof(unsigned long long);
          if (cudaMemsetAsync(xg[dev], 0, used, st) != cudaSuccess) GEMM_FAIL("conv_gemm: clearing the GroupNorm exchange buffer failed");
          return launch_umma<128, EPI_GNF, 2, 2, true>(op, a, st);
        }
        if (op->block_n == 128 && op->m_sub == 2 && op->cg == 1) return launch_umma<128, EPI_GNF, 2, 1, true>(op, a, st);
        GEMM_FAIL("conv_gemm: no GroupNorm-epilogue halo kernel for block_n %d, m_sub %d, cg %d", op->block_n, op->m_sub, op->cg);
      }
      if (op->cg == 2) {
        if (op->block_n == 256 && op->m_sub == 1) return launch_umma<256, EPI_GNF, 1, 2>(op, a, st);
        GEMM_FAIL("conv_gemm: CTA pairs need block_n 256");
      }
      if (op->m_sub == 2) {
        switch (op->block_n) {
          case 128: return launch_umma<128, EPI_GNF, 2>(op, a, st);
          case 64: return launch_umma<64, EPI_GNF, 2>(op, a, st);
          default: GEMM_FAIL("conv_gemm: m_sub=2 needs block_n 64 or 128 (got %d)", op->block_n);
        }
      }
      switch (op->block_n) {
        case 256: return launch_umma<256, EPI_GNF, 1>(op, a, st);
        case 128: return launch_umma<128, EPI_GNF, 1>(op, a, st);
        case 64: return launch_umma<64, EPI_GNF, 1>(op, a, st);
        case 32: return launch_umma<32, EPI_GNF, 1>(op, a, st);
        default: GEMM_FAIL("conv_gemm: unsupported block_n %d", op->block_n);
      }
    }
    if (op->halo) {
      if (op->block_n == 256 && op->m_sub == 1 && op->cg == 2) return launch_umma<256, EPI_LINEAR, 1, 2, true>(op, a, st);
      if (op->block_n == 128 && op->m_sub == 2 && op->cg == 1) return launch_umma<128, EPI_LINEAR, 2, 1, true>(op, a, st);
      if (op->block_n == 128 && op->m_sub == 2 && op->cg == 2) return launch_umma<128, EPI_LINEAR, 2, 2, true>(op, a, st);
      GEMM_FAIL("conv_gemm: no halo kernel for block_n %d, m_sub %d, cg %d", op->block_n, op->m_sub, op->cg);
    }
    if (op->cg == 2) {
      if (op->block_n == 256 && op->m_sub == 1) return launch_umma<256, EPI_LINEAR, 1, 2>(op, a, st);
      GEMM_FAIL("conv_gemm: CTA pairs need block_n 256");
    }
    if (op->m_sub == 2) {
      switch (op->block_n) {
        case 128: return launch_umma<128, EPI_LINEAR, 2>(op, a, st);
        case 64: return launch_umma<64, EPI_LINEAR, 2>(op, a, st);
        default: GEMM_FAIL("conv_gemm: m_sub=2 needs block_n 64 or 128 (got %d)", op->block_n);
      }
    }
    switch (op->block_n) {
      case 256: return launch_umma<256, EPI_LINEAR, 1>(op, a, st);
      case 128: return launch_umma<128, EPI_LINEAR, 1>(op, a, st);
      case 64: return launch_umma<64, EPI_LINEAR, 1>(op, a, st);
      case 32: return launch_umma<32, EPI_LINEAR, 1>(op, a, st);
      default: GEMM_FAIL("conv_gemm: unsupported block_n %d", op->block_n);
    }
  }
  // reference path
  RefArgs r;
  memset(&r, 0, sizeof(r));
  for (int s = 0; s < op->nseg; ++s) {
    r.a[s] = op->seg[s].ptr; r.ctot[s] = op->seg[s].c_total; r.coff[s] = op->seg[s].c_off;
    r.c[s] = op->seg[s].c; r.taps[s] = op->seg[s].taps;
  }
  r.nseg = op->nseg; r.B = op->B; r.H = op->H; r.W = op->W; r.M = (int)M; r.N = op->N;
  r.w = op->w; r.w_ld = op->w_ld; r.w_koff = op->w_koff; r.w_batch_stride = op->w_batch_stride;
  r.w_lo_off = 0;
  if (op->wsplit == 2) for (int s = 0; s < op->nseg; ++s) r.w_lo_off += op->seg[s].taps * op->seg[s].c;
  r.tiles_per_batch_rows = op->H * op->W;
  r.bias = op->bias; r.bias2 = op->bias2; r.residual = op->residual; r.rowscale = op->rowscale;
  r.scale = op->scale; r.out32 = op->out32; r.out16 = op->out16; r.row_out = op->row_out; r.ldo = op->ldo;
  r.epi = op->epi;
  r.n_store = op->n_store;
  if (op->epi == EPI_GNF) {
    if (!op->gn_gamma || !op->gn_beta || !op->out16 || op->gn_groups <= 0 || op->N % op->gn_groups != 0)
      GEMM_FAIL("conv_gemm ref: bad GroupNorm epilogue arguments");
    const size_t need = (size_t)M * op->N * sizeof(float);
    if (need > g_softmax_tmp_bytes) {
      if (g_softmax_tmp) cudaFree(g_softmax_tmp);
      if (cudaMalloc(&g_softmax_tmp, need) != cudaSuccess) GEMM_FAIL("conv_gemm ref: scratch alloc failed");
      g_softmax_tmp_bytes = need;
    }
    r.epi = EPI_LINEAR; r.out16 = nullptr;
    const float* vsrc = op->out32;
    int vld = op->ldo;
    if (op->out32 == nullptr) { r.out32 = g_softmax_tmp; r.ldo = op->N; vsrc = g_softmax_tmp; vld = op->N; }   // out16 only
    dim3 grid0((unsigned)((M + 63) / 64), (unsigned)((op->N + 63) / 64));
    launch_k(conv_gemm_ref_kernel, dim3(grid0), dim3(256), 0, st, r);
    if (op->colstats != nullptr && op->out32 != nullptr) {
      dim3 g2((unsigned)((M + 31) / 32), (unsigned)((op->N + 127) / 128));
      launch_k(colstats_ref_kernel, dim3(g2), dim3(128), 0, st, op->out32, op->colstats, (int)M, op->N, op->ldo);
    }
    launch_k(gnf_ref_kernel, dim3((unsigned)op->B, (unsigned)op->gn_groups), dim3(256), 0, st, vsrc, op->out16,
             op->H * op->W, vld, op->N / op->gn_groups, op->gn_gamma, op->gn_beta, op->gn_eps, op->gn_silu, op->ldo);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) GEMM_FAIL("conv_gemm_ref (GroupNorm epilogue) launch: %s", cudaGetErrorString(e));
    return 0;
  }
  if (op->w_batch_stride != 0 && (op->H * op->W) % 64 != 0) GEMM_FAIL("conv_gemm ref: batched B needs H*W %% 64 == 0");
  if (op->epi == EPI_SOFTMAX) {
    const size_t need = (size_t)M * op->N * sizeof(float);
    if (need > g_softmax_tmp_bytes) {
      if (g_softmax_tmp) cudaFree(g_softmax_tmp);
      if (cudaMalloc(&g_softmax_tmp, need) != cudaSuccess) GEMM_FAIL("conv_gemm ref: softmax scratch alloc failed");
      g_softmax_tmp_bytes = need;
    }
    r.softmax_tmp = g_softmax_tmp;
  }
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)((op->N + 63) / 64));
  launch_k(conv_gemm_ref_kernel, dim3(grid), dim3(256), 0, st, r);
  if (op->colstats != nullptr && op->out32 != nullptr) {
    dim3 g2((unsigned)((M + 31) / 32), (unsigned)((op->N + 127) / 128));
    launch_k(colstats_ref_kernel, dim3(g2), dim3(128), 0, st, op->out32, op->colstats, (int)M, op->N, op->ldo);
  }
  if (op->epi == EPI_SOFTMAX)
    launch_k(softmax_ref_kernel, dim3((unsigned)((M + 127) / 128)), dim3(128), 0, st, g_softmax_tmp, op->out16, op->row_out, (int)M, op->N, op->ldo);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) GEMM_FAIL("conv_gemm_ref launch: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace gddim
