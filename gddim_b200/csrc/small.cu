// Small CUDA-core kernels around the tensor-core GEMMs: stem / head convolutions (6 or 3 channels on one
// side: < 0.1 % of the FLOPs, no tensor-core path), the FIR + stride-2 window gather of the input pyramid,
// V transposition for the PV GEMM, 16-token attention of the 4x4 middle block, and the tiny dense layers of
// the time embedding.
//
// Reference semantics: cld_jax/models/ncsnpp.py:146 (stem conv3x3), :237 (head conv3x3, init_scale),
// layerspp.py:115-143 + up_or_down_sampling.py:168-209 (Downsample fir+with_conv = conv_downsample_2d),
// layerspp.py:61-83 (AttnBlockpp), ncsnpp.py:87-89 + layerspp.py:216 (Dense).
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "launch.cuh"

namespace gddim {

static int ceil_div_ll(long long a, long long b) { return int((a + b - 1) / b); }

// ---- input pyramid: FIR (pad 2) then 3x3 stride-2 VALID window gather -> GEMM A operand ---------------------
__global__ void __launch_bounds__(256) im2col_fir_down_kernel(const float* __restrict__ in, __half* __restrict__ a16,
                                                             int B, int H, int W, int c, int kpad, int use_fir,
                                                             float out_scale) {
  pdl_entry();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * Ho * Wo * kpad;
  const float kf[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = int(idx % kpad);
    const long long opix = idx / kpad;
    float v = 0.f;
    if (k < 9 * c) {
      const int ch = k % c, tap = k / c;
      const int ky = tap / 3, kx = tap % 3;
      const int ox = int(opix % Wo), oy = int((opix / Wo) % Ho);
      const long long b = opix / ((long long)Wo * Ho);
      if (use_fir) {
        // FIR output grid is (H+1)x(W+1): f[py][px] = sum_{i,j} k[i]k[j] in[py + i - 2][px + j - 2]
        const int py = 2 * oy + ky, px = 2 * ox + kx;
        for (int i = 0; i < 4; ++i) {
          const int iy = py + i - 2;
          if (iy < 0 || iy >= H) continue;
          for (int j = 0; j < 4; ++j) {
            const int ix = px + j - 2;
            if (ix < 0 || ix >= W) continue;
            v += kf[i] * kf[j] * in[((b * H + iy) * W + ix) * c + ch];
          }
        }
      } else {
        // plain stride-2 SAME 3x3 (pad (0,1)): used only when fir is off
        const int iy = 2 * oy + ky, ix = 2 * ox + kx;
        if (iy < H && ix < W) v = in[((b * H + iy) * W + ix) * c + ch];
      }
    }
    a16[idx] = __float2half_rn(v * out_scale);
  }
}

// 8 channels per thread (c % 8 == 0, kpad == 9*c): float4 loads, one 16-byte store
__global__ void __launch_bounds__(256) im2col_fir_down_vec8_kernel(const float* __restrict__ in, __half* __restrict__ a16,
                                                                  int B, int H, int W, int c, float out_scale) {
  pdl_entry();
  const int Ho = H / 2, Wo = W / 2;
  const int cv = c / 8;
  const long long total = (long long)B * Ho * Wo * 9 * cv;
  const float kf[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ch = int(idx % cv) * 8;
    const int tap = int((idx / cv) % 9);
    const long long opix = idx / (9 * cv);
    const int ky = tap / 3, kx = tap % 3;
    const int ox = int(opix % Wo), oy = int((opix / Wo) % Ho);
    const long long b = opix / ((long long)Wo * Ho);
    const int py = 2 * oy + ky, px = 2 * ox + kx;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int iy = py + i - 2;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ix = px + j - 2;
        if (ix < 0 || ix >= W) continue;
        const float w = kf[i] * kf[j];
        const float* q = in + ((b * H + iy) * W + ix) * c + ch;
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(q)), v1 = __ldg(reinterpret_cast<const float4*>(q + 4));
        acc[0] += w * v0.x; acc[1] += w * v0.y; acc[2] += w * v0.z; acc[3] += w * v0.w;
        acc[4] += w * v1.x; acc[5] += w * v1.y; acc[6] += w * v1.z; acc[7] += w * v1.w;
      }
    }
    __half h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = __float2half_rn(acc[j] * out_scale);
    *reinterpret_cast<uint4*>(a16 + opix * (9 * c) + tap * c + ch) = *reinterpret_cast<const uint4*>(h);
  }
}

// Same result through shared memory: the FIR output grid f (H + 1) x (W + 1) of the rows a CTA needs is computed ONCE
// (the direct kernels recompute the 16-tap FIR for each of the 9 window positions it belongs to and are bound by
// instruction issue: 72 us for the 16x16x128 level whose bytes take 11 us), then the nine stride-2 windows are copied
// out.  CTA = (output-row pair, 32-channel chunk, image); staged input rows carry two zero columns on either side and
// zero rows outside the image, so no tap needs a bounds test.  Tap order and weights as in the direct kernels.
__global__ void __launch_bounds__(256) im2col_fir_down_tiled_kernel(const float* __restrict__ in, __half* __restrict__ a16,
                                                                   int H, int W, int c, int ro, float out_scale) {
  pdl_entry();
  extern __shared__ __align__(16) float ism[];
  constexpr int CC = 32;
  const int Ho = H / 2, Wo = W / 2;
  const int nxr = 2 * ro + 4, nfr = 2 * ro + 1;         // staged input rows, FIR rows
  const int Wx = W + 4, Wf = W + 1;
  float* X = ism;                                         // [nxr][Wx][CC]
  float* F = ism + (size_t)nxr * Wx * CC;                 // [nfr][Wf][CC]
  const int oy0 = blockIdx.x * ro, c0 = blockIdx.y * CC;
  const long long b = blockIdx.z;
  const int v = threadIdx.x & 7, pr = threadIdx.x >> 3;   // 4-channel vector, pixel slot
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // input row r <-> image row 2 oy0 - 2 + r, column x <-> image column x - 2
  for (int e = pr; e < nxr * Wx; e += 32) {
    const int r = e / Wx, x = e - r * Wx;
    const int iy = 2 * oy0 - 2 + r, ix = x - 2;
    float4 t = z4;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) t = __ldg(reinterpret_cast<const float4*>(in + ((b * H + iy) * W + ix) * c + c0 + v * 4));
    *reinterpret_cast<float4*>(X + (size_t)e * CC + v * 4) = t;
  }
  __syncthreads();
  // f[py][px] = sum_{i, j} k[i] k[j] in[py + i - 2][px + j - 2]; FIR row q <-> py = 2 oy0 + q: input rows q + i, columns px + j
  const float kf[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  for (int e = pr; e < nfr * Wf; e += 32) {
    const int q = e / Wf, px = e - q * Wf;
    float4 acc = z4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w = kf[i] * kf[j];
        const float4 t = *reinterpret_cast<const float4*>(X + ((size_t)(q + i) * Wx + px + j) * CC + v * 4);
        acc.x += w * t.x; acc.y += w * t.y; acc.z += w * t.z; acc.w += w * t.w;
      }
    *reinterpret_cast<float4*>(F + (size_t)e * CC + v * 4) = acc;
  }
  __syncthreads();
  // windows: a16[opix][tap * c + ch] = f[2 oy + ky][2 ox + kx][ch] * out_scale
  for (int e = pr; e < ro * Wo * 9; e += 32) {
    const int tap = e % 9, op = e / 9;
    const int orow = op / Wo, ox = op - orow * Wo;
    if (oy0 + orow >= Ho) break;
    const int ky = tap / 3, kx = tap - ky * 3;
    const float4 t = *reinterpret_cast<const float4*>(F + ((size_t)(2 * orow + ky) * Wf + 2 * ox + kx) * CC + v * 4);
    const __half2 h0 = __floats2half2_rn(t.x * out_scale, t.y * out_scale), h1 = __floats2half2_rn(t.z * out_scale, t.w * out_scale);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&h0);
    pk.y = *reinterpret_cast<const uint32_t*>(&h1);
    const long long opix = (b * Ho + oy0 + orow) * Wo + ox;
    *reinterpret_cast<uint2*>(a16 + opix * (9 * c) + tap * c + c0 + v * 4) = pk;
  }
}

// few channels (the c = 6 input of the first pyramid level, kpad = 64): one CTA per image, the whole image and its FIR grid
// in shared memory, then one thread per output element of the [Ho * Wo][kpad] operand (zero beyond 9 c)
__global__ void __launch_bounds__(256) im2col_fir_down_image_kernel(const float* __restrict__ in, __half* __restrict__ a16,
                                                                   int H, int W, int c, int kpad, float out_scale) {
  pdl_entry();
  extern __shared__ __align__(16) float ism[];
  const int Wx = W + 4, Hx = H + 4, Wf = W + 1, Hf = H + 1;
  float* X = ism;                                   // [Hx][Wx][c], image at offset (2, 2), zeros around
  float* F = ism + (size_t)Hx * Wx * c;             // [Hf][Wf][c]
  const long long b = blockIdx.x;
  for (int e = threadIdx.x; e < Hx * Wx * c; e += blockDim.x) {
    const int ch = e % c, x = (e / c) % Wx, y = e / (c * Wx);
    const int iy = y - 2, ix = x - 2;
    X[e] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(in + ((b * H + iy) * W + ix) * c + ch) : 0.f;
  }
  __syncthreads();
  const float kf[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  for (int e = threadIdx.x; e < Hf * Wf * c; e += blockDim.x) {
    const int ch = e % c, px = (e / c) % Wf, py = e / (c * Wf);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc += kf[i] * kf[j] * X[((py + i) * Wx + px + j) * c + ch];
    F[e] = acc;
  }
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2;
  for (int e = threadIdx.x; e < Ho * Wo * kpad; e += blockDim.x) {
    const int k = e % kpad, op = e / kpad;
    float vv = 0.f;
    if (k < 9 * c) {
      const int ch = k % c, tap = k / c;
      const int ky = tap / 3, kx = tap - ky * 3;
      const int oy = op / Wo, ox = op - oy * Wo;
      vv = F[((2 * oy + ky) * Wf + 2 * ox + kx) * c + ch];
    }
    a16[(b * Ho * Wo + op) * kpad + k] = __float2half_rn(vv * out_scale);
  }
}

int im2col_fir_down_launch(const float* in, __half* a16, int B, int H, int W, int c, int kpad, int use_fir,
                           float out_scale, cudaStream_t st) {
  static int tiled = -1;                        // GDDIM_NO_IM2COL_TILED=1: A/B switch back to the direct kernels
  if (tiled < 0) { const char* e = getenv("GDDIM_NO_IM2COL_TILED"); tiled = (e && e[0] == '1') ? 0 : 1; }
  if (tiled && use_fir && c % 32 == 0 && kpad == 9 * c && H % 4 == 0 && W % 2 == 0) {
    const int ro = 2;
    const size_t smem = ((size_t)(2 * ro + 4) * (W + 4) + (size_t)(2 * ro + 1) * (W + 1)) * 32 * sizeof(float);
    if (smem <= 96 * 1024) {
      static DeviceOnce attr_set;
      if (attr_set.need()) {
        cudaFuncSetAttribute(im2col_fir_down_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr_set.done();
      }
      launch_k(im2col_fir_down_tiled_kernel, dim3(H / 2 / ro, c / 32, B), dim3(256), smem, st, in, a16, H, W, c, ro, out_scale);
      return cudaGetLastError() == cudaSuccess ? 0 : -2;
    }
  }
  if (tiled && use_fir && c <= 8 && H % 2 == 0 && W % 2 == 0) {
    const size_t smem = ((size_t)(H + 4) * (W + 4) + (size_t)(H + 1) * (W + 1)) * c * sizeof(float);
    if (smem <= 96 * 1024) {
      static DeviceOnce attr_set2;
      if (attr_set2.need()) {
        cudaFuncSetAttribute(im2col_fir_down_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr_set2.done();
      }
      launch_k(im2col_fir_down_image_kernel, dim3(B), dim3(256), smem, st, in, a16, H, W, c, kpad, out_scale);
      return cudaGetLastError() == cudaSuccess ? 0 : -2;
    }
  }
  if (use_fir && c % 8 == 0 && kpad == 9 * c) {
    const long long total = (long long)B * (H / 2) * (W / 2) * 9 * (c / 8);
    int grid = ceil_div_ll(total, 256);
    if (grid > 148 * 32) grid = 148 * 32;
    launch_k(im2col_fir_down_vec8_kernel, dim3(grid), dim3(256), 0, st, in, a16, B, H, W, c, out_scale);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
  }
  const long long total = (long long)B * (H / 2) * (W / 2) * kpad;
  int grid = ceil_div_ll(total, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  launch_k(im2col_fir_down_kernel, dim3(grid), dim3(256), 0, st, in, a16, B, H, W, c, kpad, use_fir, out_scale);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ---- stem: 3x3 SAME window gather so that the 6 -> nf convolution runs as one K=64 GEMM k-block ---------------
__global__ void __launch_bounds__(256) im2col_same3x3_kernel(const float* __restrict__ in, __half* __restrict__ a16,
                                                            int B, int H, int W, int c, int kpad, float out_scale,
                                                            int split, int kseg) {
  pdl_entry();
  // one thread per (pixel, 8 consecutive k): 16-byte stores
  const int kv = kpad / 8;
  const long long total = (long long)B * H * W * kv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k0 = int(idx % kv) * 8;
    const long long pix = idx / kv;
    const int x = int(pix % W), y = int((pix / W) % H);
    const long long b = pix / ((long long)W * H);
    // split mode (kpad == 3*kseg): columns [0,kseg) = hi(x), [kseg,2kseg) = lo(x) = x - hi(x), [2kseg,3kseg) = hi(x)
    // again; with weight rows (hi(w), hi(w), lo(w)) the GEMM evaluates x*w to ~2^-22 relative with fp16 operands.
    const int seg = split ? k0 / kseg : 0;
    __half h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j - seg * kseg;
      float v = 0.f;
      if (k < 9 * c) {
        const int ch = k % c, tap = k / c;
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(in + ((b * H + yy) * W + xx) * c + ch) * out_scale;
      }
      const __half hi = __float2half_rn(v);
      h[j] = (seg == 1) ? __float2half_rn(v - __half2float(hi)) : hi;
    }
    *reinterpret_cast<uint4*>(a16 + pix * kpad + k0) = *reinterpret_cast<const uint4*>(h);
  }
}

// split layout, 9*c <= 64: one thread per pixel builds the 64-wide hi and lo rows in registers and writes the three
// 128-byte segments (hi, lo, hi) with 16-byte stores
template <int C>
__global__ void __launch_bounds__(128) im2col_stem_split_kernel(const float* __restrict__ in, __half* __restrict__ a16,
                                                               int B, int H, int W, float out_scale) {
  pdl_entry();
  const long long npix = (long long)B * H * W;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int x = int(pix % W), y = int((pix / W) % H);
  const long long b = pix / ((long long)W * H);
  __half hi[64], lo[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) { hi[k] = __float2half(0.f); lo[k] = __float2half(0.f); }
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const float* q = in + ((b * H + yy) * W + xx) * C;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
      const float v = __ldg(q + ch) * out_scale;
      const __half h = __float2half_rn(v);
      hi[tap * C + ch] = h;
      lo[tap * C + ch] = __float2half_rn(v - __half2float(h));
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(a16 + pix * 192);
  const uint4* ph = reinterpret_cast<const uint4*>(hi);
  const uint4* pl = reinterpret_cast<const uint4*>(lo);
#pragma unroll
  for (int j = 0; j < 8; ++j) { dst[j] = ph[j]; dst[8 + j] = pl[j]; dst[16 + j] = ph[j]; }
}

int im2col_same3x3_launch(const float* in, __half* a16, int B, int H, int W, int c, int kpad, float out_scale,
                          int split, cudaStream_t st) {
  if (kpad % 8 != 0 || (split && kpad % 24 != 0)) return -1;
  if (split && kpad == 192 && (c == 6 || c == 3)) {
    const long long npix = (long long)B * H * W;
    const int grid = ceil_div_ll(npix, 128);
    if (c == 6) launch_k(im2col_stem_split_kernel<6>, dim3(grid), dim3(128), 0, st, in, a16, B, H, W, out_scale);
    else launch_k(im2col_stem_split_kernel<3>, dim3(grid), dim3(128), 0, st, in, a16, B, H, W, out_scale);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
  }
  const int kseg = split ? kpad / 3 : kpad;
  const long long total = (long long)B * H * W * (kpad / 8);
  int grid = ceil_div_ll(total, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  launch_k(im2col_same3x3_kernel, dim3(grid), dim3(256), 0, st, in, a16, B, H, W, c, kpad, out_scale, split, kseg);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ---- V^T: qkv16 [B,T,ld] (V at voff) -> vT [B,C,T]; 32x32 tiles through shared memory ------------------------
__global__ void __launch_bounds__(256) transpose_v_kernel(const __half* __restrict__ qkv, __half* __restrict__ vT, int T,
                                                         int C, int ld, int voff) {
  pdl_entry();
  __shared__ __half tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, c = c0 + tx;
    tile[r][tx] = (t < T && c < C) ? qkv[((long long)b * T + t) * ld + voff + c] : __float2half(0.f);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, t = t0 + tx;
    if (c < C && t < T) vT[((long long)b * C + c) * T + t] = tile[tx][r];
  }
}

// 64x64 tiles, 16-byte global accesses on both sides (T, C, ld, voff multiples of 8; T, C multiples of 64)
__global__ void __launch_bounds__(256) transpose_v64_kernel(const __half* __restrict__ qkv, __half* __restrict__ vT, int T,
                                                           int C, int ld, int voff) {
  pdl_entry();
  __shared__ __align__(16) __half tile[64][72];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int v = threadIdx.x + k * 256;          // 512 vectors: row = v / 8, 8-channel group = v % 8
    const int r = v >> 3, g = v & 7;
    *reinterpret_cast<uint4*>(&tile[r][g * 8]) =
        *reinterpret_cast<const uint4*>(qkv + ((long long)b * T + t0 + r) * ld + voff + c0 + g * 8);
  }
  __syncthreads();
  const int c = threadIdx.x & 63;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int tv = (threadIdx.x >> 6) + k * 4;    // 8-token group 0..7
    __half h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = tile[tv * 8 + j][c];
    *reinterpret_cast<uint4*>(vT + ((long long)b * C + c0 + c) * T + t0 + tv * 8) = *reinterpret_cast<const uint4*>(h);
  }
}

int transpose_v_launch(const __half* qkv, __half* vT, int B, int T, int C, int ld, int voff, cudaStream_t st) {
  if (T % 64 == 0 && C % 64 == 0 && ld % 8 == 0 && voff % 8 == 0) {
    dim3 grid(T / 64, C / 64, B);
    launch_k(transpose_v64_kernel, dim3(grid), dim3(256), 0, st, qkv, vT, T, C, ld, voff);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
  }
  dim3 grid((T + 31) / 32, (C + 31) / 32, B);
  launch_k(transpose_v_kernel, dim3(grid), dim3(256), 0, st, qkv, vT, T, C, ld, voff);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ---- attention for very short sequences (the 4x4 middle block: T = 16) ---------------------------------------
__global__ void __launch_bounds__(256) small_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ o16, int T,
                                                        int C, float scale) {
  pdl_entry();
  extern __shared__ float sm[];
  float* sq = sm;                 // [T][C]
  float* sk = sq + T * C;
  float* sv = sk + T * C;
  float* ss = sv + T * C;         // [T][T]
  const int b = blockIdx.x;
  const __half* base = qkv + (long long)b * T * 3 * C;
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) {
    const int t = i / C, c = i % C;
    sq[i] = __half2float(base[t * 3 * C + c]);
    sk[i] = __half2float(base[t * 3 * C + C + c]);
    sv[i] = __half2float(base[t * 3 * C + 2 * C + c]);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int e = warp; e < T * T; e += nw) {
    const int i = e / T, j = e % T;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc += sq[i * C + c] * sk[j * C + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) ss[e] = acc * scale;
  }
  __syncthreads();
  if (threadIdx.x < T) {
    const int i = threadIdx.x;
    float mx = -INFINITY;
    for (int j = 0; j < T; ++j) mx = fmaxf(mx, ss[i * T + j]);
    float sum = 0.f;
    for (int j = 0; j < T; ++j) { const float e = expf(ss[i * T + j] - mx); ss[i * T + j] = e; sum += e; }
    const float inv = 1.0f / sum;
    for (int j = 0; j < T; ++j) ss[i * T + j] *= inv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) {
    const int t = i / C, c = i % C;
    float acc = 0.f;
    for (int j = 0; j < T; ++j) acc += ss[t * T + j] * sv[j * C + c];
    o16[((long long)b * T + t) * C + c] = __float2half_rn(acc);
  }
}

int small_attn_launch(const __half* qkv, __half* o16, int B, int T, int C, float scale, cudaStream_t st) {
  const size_t smem = (size_t)(3 * T * C + T * T) * sizeof(float);
  if (smem > 200 * 1024 || T > 256) return -1;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaFuncSetAttribute(small_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  launch_k(small_attn_kernel, dim3(B), dim3(256), smem, st, qkv, o16, T, C, scale);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ---- row softmax for long sequences (T > 256 keys: the 32x32-token middle attention of the 256x256 config) ------
// One warp per query row: p16 = exp(s - max) in fp16, rowinv = 1 / sum of the ROUNDED numerators (what the P.V GEMM
// will actually add up), same convention as the fused N = 256 epilogue.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s32, __half* __restrict__ p16,
                                                          float* __restrict__ rowinv, long long rows, int T) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(s32 + row * T);
  float mx = -INFINITY;
  for (int j = lane; j < T / 4; j += 32) {
    const float4 v = __ldg(src + j);
    mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  uint2* dst = reinterpret_cast<uint2*>(p16 + row * T);
  for (int j = lane; j < T / 4; j += 32) {
    const float4 v = __ldg(src + j);
    const __half2 a = __floats2half2_rn(__expf(v.x - mx), __expf(v.y - mx));
    const __half2 b = __floats2half2_rn(__expf(v.z - mx), __expf(v.w - mx));
    const float2 fa = __half22float2(a), fb = __half22float2(b);
    sum += (fa.x + fa.y) + (fb.x + fb.y);
    uint2 o;
    o.x = *reinterpret_cast<const unsigned*>(&a);
    o.y = *reinterpret_cast<const unsigned*>(&b);
    dst[j] = o;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) rowinv[row] = 1.f / sum;
}

int softmax_rows_launch(const float* s32, __half* p16, float* rowinv, long long rows, int T, cudaStream_t st) {
  if (T % 4 != 0 || rows <= 0) return -1;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  launch_k(softmax_rows_kernel, dim3(grid), dim3(256), 0, st, s32, p16, rowinv, rows, T);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ---- dense: y[r,n] = bias[n] + sum_k act(x[r,k]) w[k,n]  (w in flax (in,out) layout) ------------------------
__global__ void __launch_bounds__(256) dense_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                   const float* __restrict__ bias, float* __restrict__ y, int rows, int K,
                                                   int N, int silu_in) {
  pdl_entry();
  const long long total = (long long)rows * N;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = int(idx % N);
    const long long r = idx / N;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      float v = x[r * K + k];
      if (silu_in) v = v / (1.0f + expf(-v));
      acc += v * w[(long long)k * N + n];
    }
    y[idx] = acc + (bias ? bias[n] : 0.f);
  }
}

int dense_launch(const float* x, const float* w, const float* bias, float* y, int rows, int K, int N, int silu_in,
                 cudaStream_t st) {
  const long long total = (long long)rows * N;
  int grid = ceil_div_ll(total, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  launch_k(dense_kernel, dim3(grid), dim3(256), 0, st, x, w, bias, y, rows, K, N, silu_in);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

__global__ void add_vec_kernel(const float* a, const float* b, float* y, int n) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] + b[i];
}
int add_vec_launch(const float* a, const float* b, float* y, int n, cudaStream_t st) {
  launch_k(add_vec_kernel, dim3((n + 255) / 256), dim3(256), 0, st, a, b, y, n);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace gddim
