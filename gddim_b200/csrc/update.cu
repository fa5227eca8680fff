// K5 / K6: the gDDIM / DEIS linear-algebra update and the blur-diffusion DCT-space update.
// HBM-bound elementwise work, no tensor-core path: coalesced float4 traffic, one pass.
//
// Reference semantics:
//   cld_jax/deis.py:141-151           multistep_ab_step   u' = Psi u + sum_j C_j eps_j  (2x2 on the (x,v) pair)
//   cld_jax/sampling.py:30-39         denoising step (same algebraic form with A = I - eps F, C = -eps GG R^-T)
//   cld_jax/models/utils.py:153,158   '(d g) <-> (g d)' relayout, :174-176 mixed_score
//   blur_jax/sampling.py:60-75        order-0 DDIM update in DCT space
//   blur_jax/blur.py:11-107           orthonormal DCT-II / DCT-III (here: dense 32x32 transforms in smem)
//   blur_jax/multistep.py:94-98       scalar-coefficient ab_step
#include <cmath>
#include <cstdio>

#include "kernels.h"
#include "launch.cuh"
#include "cld_update.cuh"

namespace gddim {

static int grid_for(long long n, int per_block, int cap = 148 * 16) {
  long long g = (n + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

// ---- CLD step in net layout --------------------------------------------------------------------------
struct CldStepDev {
  const float* u; float* u_out;
  const float* eps[6];
  int n_eps;
  float coef[7][4];
  int mixed; float mixm[4];
  float* eps_store;
  long long n_pix; int C;
  int noise_mode;
  const float* noise;
  float nfac[4];
  unsigned long long seed, stream_id;
};

// Philox4x32-10 (Salmon et al.): counter-based, one call = 4 uniform words
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
// 4 standard normals from one Philox call (Box-Muller on (0,1] uniforms)
__device__ __forceinline__ float4 philox_normal4(unsigned long long idx, unsigned long long stream, unsigned int sub,
                                                 unsigned long long seed) {
  const uint4 r = philox4x32(make_uint4((unsigned int)idx, (unsigned int)(idx >> 32), (unsigned int)stream, sub),
                             make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)));
  const float u0 = ((float)r.x + 1.0f) * 2.3283064e-10f, u1 = (float)r.y * 2.3283064e-10f;
  const float u2 = ((float)r.z + 1.0f) * 2.3283064e-10f, u3 = (float)r.w * 2.3283064e-10f;
  const float m0 = sqrtf(-2.0f * logf(u0)), m1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  return make_float4(m0 * c0, m0 * s0, m1 * c1, m1 * s1);
}

// Specialisation for C = 3 (pixel = 6 floats): a thread owns 2 pixels = 3 float4 per array.  The per-pixel algebra is
// cld_update.cuh's (shared with the head convolution's epilogue, which applies the same update in place of this kernel
// for the deterministic samplers).
__global__ void __launch_bounds__(256) cld_step_c3_kernel(const CldStepDev p) {
  pdl_entry();
  const long long npair = p.n_pix / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npair;
       i += (long long)gridDim.x * blockDim.x) {
    float u[2][6], accp[2][6];
    {
      const float4* q = reinterpret_cast<const float4*>(p.u) + i * 3;
      const float4 a = q[0], b = q[1], c = q[2];     // plain loads: u is updated in place by this launch
      u[0][0] = a.x; u[0][1] = a.y; u[0][2] = a.z; u[0][3] = a.w; u[0][4] = b.x; u[0][5] = b.y;
      u[1][0] = b.z; u[1][1] = b.w; u[1][2] = c.x; u[1][3] = c.y; u[1][4] = c.z; u[1][5] = c.w;
    }
    cld_px_apply(accp[0], u[0], p.coef[0][0], p.coef[0][1], p.coef[0][2], p.coef[0][3]);
    cld_px_apply(accp[1], u[1], p.coef[0][0], p.coef[0][1], p.coef[0][2], p.coef[0][3]);
    for (int j = 0; j < p.n_eps; ++j) {
      float e[2][6];
      const float4* q = reinterpret_cast<const float4*>(p.eps[j]) + i * 3;
      const float4 a = q[0], b = q[1], c = q[2];     // plain loads: eps[0] may be eps_store (mixed score)
      e[0][0] = a.x; e[0][1] = a.y; e[0][2] = a.z; e[0][3] = a.w; e[0][4] = b.x; e[0][5] = b.y;
      e[1][0] = b.z; e[1][1] = b.w; e[1][2] = c.x; e[1][3] = c.y; e[1][4] = c.z; e[1][5] = c.w;
      if (j == 0 && p.mixed) {
        cld_px_mix(e[0], u[0], p.mixm[0], p.mixm[1], p.mixm[2], p.mixm[3]);
        cld_px_mix(e[1], u[1], p.mixm[0], p.mixm[1], p.mixm[2], p.mixm[3]);
        float4* s = reinterpret_cast<float4*>(p.eps_store) + i * 3;
        s[0] = make_float4(e[0][0], e[0][1], e[0][2], e[0][3]);
        s[1] = make_float4(e[0][4], e[0][5], e[1][0], e[1][1]);
        s[2] = make_float4(e[1][2], e[1][3], e[1][4], e[1][5]);
      }
      cld_px_acc(accp[0], e[0], p.coef[1 + j][0], p.coef[1 + j][1], p.coef[1 + j][2], p.coef[1 + j][3]);
      cld_px_acc(accp[1], e[1], p.coef[1 + j][0], p.coef[1 + j][1], p.coef[1 + j][2], p.coef[1 + j][3]);
    }
    float acc[12];
#pragma unroll
    for (int k = 0; k < 6; ++k) { acc[k] = accp[0][k]; acc[6 + k] = accp[1][k]; }
    if (p.noise_mode != 0) {
      float z[12];                 // reference layout of the two pixels: (px, d, g)
      if (p.noise_mode == 1) {
        const float4* q = reinterpret_cast<const float4*>(p.noise) + i * 3;
        const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
        z[0] = a.x; z[1] = a.y; z[2] = a.z; z[3] = a.w; z[4] = b.x; z[5] = b.y; z[6] = b.z; z[7] = b.w;
        z[8] = c.x; z[9] = c.y; z[10] = c.z; z[11] = c.w;
      } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float4 n = philox_normal4((unsigned long long)i, p.stream_id, k, p.seed);
          z[4 * k] = n.x; z[4 * k + 1] = n.y; z[4 * k + 2] = n.z; z[4 * k + 3] = n.w;
        }
      }
#pragma unroll
      for (int px = 0; px < 2; ++px)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float zx = z[px * 6 + d * 2], zv = z[px * 6 + d * 2 + 1];
          acc[px * 6 + d] += p.nfac[0] * zx + p.nfac[1] * zv;
          acc[px * 6 + 3 + d] += p.nfac[2] * zx + p.nfac[3] * zv;
        }
    }
    float4* o = reinterpret_cast<float4*>(p.u_out) + i * 3;
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    o[2] = make_float4(acc[8], acc[9], acc[10], acc[11]);
  }
}

// Generic channel count: one thread per (pixel, d) pair.
__global__ void __launch_bounds__(256) cld_step_generic_kernel(const CldStepDev p) {
  pdl_entry();
  const long long n = p.n_pix * p.C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / p.C;
    const int d = int(i % p.C);
    const long long ix = pix * 2 * p.C + d, iv = ix + p.C;
    const float x = p.u[ix], v = p.u[iv];
    float ax = p.coef[0][0] * x + p.coef[0][1] * v;
    float av = p.coef[0][2] * x + p.coef[0][3] * v;
    for (int j = 0; j < p.n_eps; ++j) {
      float ex = p.eps[j][ix], ev = p.eps[j][iv];
      if (j == 0 && p.mixed) {
        ex += p.mixm[0] * x + p.mixm[1] * v;
        ev += p.mixm[2] * x + p.mixm[3] * v;
        p.eps_store[ix] = ex;
        p.eps_store[iv] = ev;
      }
      ax += p.coef[1 + j][0] * ex + p.coef[1 + j][1] * ev;
      av += p.coef[1 + j][2] * ex + p.coef[1 + j][3] * ev;
    }
    if (p.noise_mode != 0) {
      float zx, zv;
      if (p.noise_mode == 1) { zx = p.noise[i * 2]; zv = p.noise[i * 2 + 1]; }
      else { const float4 n = philox_normal4((unsigned long long)i, p.stream_id, 7u, p.seed); zx = n.x; zv = n.y; }
      ax += p.nfac[0] * zx + p.nfac[1] * zv;
      av += p.nfac[2] * zx + p.nfac[3] * zv;
    }
    p.u_out[ix] = ax;
    p.u_out[iv] = av;
  }
}

int cld_step_launch(const CldStepArgs* a, cudaStream_t st) {
  CldStepDev d;
  d.u = a->u; d.u_out = a->u_out;
  for (int j = 0; j < 6; ++j) d.eps[j] = a->eps[j];
  d.n_eps = a->n_eps;
  for (int j = 0; j < 7; ++j) for (int k = 0; k < 4; ++k) d.coef[j][k] = a->coef[j][k];
  d.mixed = a->mixed;
  for (int k = 0; k < 4; ++k) d.mixm[k] = a->mixm[k];
  d.eps_store = a->eps_store; d.n_pix = a->n_pix; d.C = a->C;
  d.noise_mode = a->noise_mode; d.noise = a->noise; d.seed = a->seed; d.stream_id = a->stream_id;
  for (int k = 0; k < 4; ++k) d.nfac[k] = a->nfac[k];
  if (a->n_eps < 0 || a->n_eps > 6) return -1;
  if (a->C == 3 && a->n_pix % 2 == 0) {
    launch_k(cld_step_c3_kernel, dim3(grid_for(a->n_pix / 2, 256)), dim3(256), 0, st, d);
  } else {
    launch_k(cld_step_generic_kernel, dim3(grid_for(a->n_pix * a->C, 256)), dim3(256), 0, st, d);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

__global__ void cld_split_kernel(const float* __restrict__ u, float* __restrict__ x, float* __restrict__ v,
                                 long long n_pix, int C, float mul, float add) {
  pdl_entry();
  const long long n = n_pix * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / C;
    const int d = int(i % C);
    const float xv = u[pix * 2 * C + d], vv = u[pix * 2 * C + C + d];
    x[i] = xv * mul + add;
    v[i] = vv;
  }
}
int cld_split_launch(const float* u, float* x, float* v, long long n_pix, int C, float mul, float add, cudaStream_t st) {
  launch_k(cld_split_kernel, dim3(grid_for(n_pix * C, 256)), dim3(256), 0, st, u, x, v, n_pix, C, mul, add);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// reference layout [pix, d, g] <-> net layout [pix, g*C + d]
__global__ void relayout_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n_pix, int C,
                                int to_net) {
  pdl_entry();
  const long long n = n_pix * 2 * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / (2 * C);
    const int k = int(i % (2 * C));
    if (to_net) {
      const int g = k / C, d = k % C;                 // destination index k = g*C + d
      dst[i] = src[pix * 2 * C + d * 2 + g];
    } else {
      const int d = k / 2, g = k % 2;                 // destination index k = d*2 + g
      dst[i] = src[pix * 2 * C + g * C + d];
    }
  }
}
int relayout_launch(const float* src, float* dst, long long n_pix, int C, int to_net, cudaStream_t st) {
  launch_k(relayout_kernel, dim3(grid_for(n_pix * 2 * C, 256)), dim3(256), 0, st, src, dst, n_pix, C, to_net);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// deis.multistep_ab_step on reference-layout arrays: pairs (x, v) are adjacent (last axis of size 2).
__global__ void __launch_bounds__(256) ab_step_ref_kernel(const float* __restrict__ x, const float* __restrict__ coef,
                                                         const float* __restrict__ new_eps,
                                                         const float* __restrict__ hist, float* __restrict__ x_out,
                                                         float* __restrict__ hist_out, int order, long long n_pairs) {
  pdl_entry();
  __shared__ float sc[8 * 4];
  if (threadIdx.x < (order + 3) * 4) sc[threadIdx.x] = coef[threadIdx.x];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs;
       i += (long long)gridDim.x * blockDim.x) {
    const float2 u = reinterpret_cast<const float2*>(x)[i];
    float ax = sc[0] * u.x + sc[1] * u.y, av = sc[2] * u.x + sc[3] * u.y;
    const float2 e0 = reinterpret_cast<const float2*>(new_eps)[i];
    ax += sc[4] * e0.x + sc[5] * e0.y;
    av += sc[6] * e0.x + sc[7] * e0.y;
    reinterpret_cast<float2*>(hist_out)[i] = e0;                       // full[:-1][0] = new_eps
    for (int j = 0; j <= order; ++j) {
      const float2 e = reinterpret_cast<const float2*>(hist)[(long long)j * n_pairs + i];
      const float* c = sc + (2 + j) * 4;
      ax += c[0] * e.x + c[1] * e.y;
      av += c[2] * e.x + c[3] * e.y;
      if (j < order) reinterpret_cast<float2*>(hist_out)[(long long)(j + 1) * n_pairs + i] = e;
    }
    reinterpret_cast<float2*>(x_out)[i] = make_float2(ax, av);
  }
}
int ab_step_ref_layout_launch(const float* x, const float* coef, const float* new_eps, const float* hist, float* x_out,
                              float* hist_out, int order, long long n_pairs, cudaStream_t st) {
  if (order < 0 || order > 5) return -1;
  launch_k(ab_step_ref_kernel, dim3(grid_for(n_pairs, 256)), dim3(256), 0, st, x, coef, new_eps, hist, x_out, hist_out, order, n_pairs);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ---- DCT on 32x32 planes -------------------------------------------------------------------------------
__constant__ float c_dct[32 * 32];   // D[k][n] = s_k cos(pi (2n+1) k / 64), orthonormal DCT-II matrix
static bool g_dct_ready = false;

static int ensure_dct_matrix() {
  if (g_dct_ready) return 0;
  float h[32 * 32];
  for (int k = 0; k < 32; ++k)
    for (int n = 0; n < 32; ++n) {
      const double s = (k == 0) ? std::sqrt(1.0 / 32.0) : std::sqrt(2.0 / 32.0);
      h[k * 32 + n] = (float)(s * std::cos(M_PI * (2.0 * n + 1.0) * k / 64.0));
    }
  if (cudaMemcpyToSymbol(c_dct, h, sizeof(h)) != cudaSuccess) return -1;
  g_dct_ready = true;
  return 0;
}

// Separable 32x32 transform of a [32][32][C] image held in shared memory, register-tiled: a thread owns a 4 x 4 block of
// outputs (4 transform rows x 4 consecutive (pixel, channel) columns) and reads one float4 of the transposed transform
// matrix and one float4 of the image per reduction step -- 2 shared-memory loads per 16 FMAs (the first version issued
// 2 loads per FMA and was bound by the shared-memory pipe: 55 us per blur step at batch 256).
// One pass:  dst[((j / C) * 32 + r) * C + j % C] = sum_h M[r][h] X[h][j]   (X: 32 rows of NC = 32 C columns; MT = M^T)
// i.e. the result is stored with the row index and the column's pixel index swapped, so the second pass (over the other
// image axis) is the same routine again and restores the [h][w][c] order.
__device__ __forceinline__ void dct_pass(const float* X, float* dst, const float* MT, int C) {
  const int NC = 32 * C;
  const int kq = threadIdx.x & 7;
  for (int jq = threadIdx.x >> 3; jq < NC / 4; jq += blockDim.x >> 3) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 8
    for (int h = 0; h < 32; ++h) {
      const float4 m = *reinterpret_cast<const float4*>(MT + h * 32 + 4 * kq);
      const float4 x = *reinterpret_cast<const float4*>(X + h * NC + 4 * jq);
      const float mm[4] = {m.x, m.y, m.z, m.w}, xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(mm[a], xx[b], acc[a][b]);
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = 4 * jq + b, w = j / C, c = j - w * C;
#pragma unroll
      for (int a = 0; a < 4; ++a) dst[(w * 32 + 4 * kq + a) * C + c] = acc[a][b];
    }
  }
}
// fwd:  Y = D X D^T  (over h, then over w);  inverse: X = D^T Y D.  sD = D, sDT = D^T (both row-major in shared memory).
// In place: img -> tmp -> img; ends with a barrier.
__device__ void dct_planes(float* img, float* tmp, const float* sD, const float* sDT, int C, int fwd) {
  const float* MT = fwd ? sDT : sD;
  dct_pass(img, tmp, MT, C);          // tmp[w][k][c]
  __syncthreads();
  dct_pass(tmp, img, MT, C);          // img[k][l][c]
  __syncthreads();
}
__device__ __forceinline__ void dct_load_matrix(float* sD, float* sDT) {
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    const float v = c_dct[i];
    sD[i] = v;
    sDT[(i & 31) * 32 + (i >> 5)] = v;
  }
}

__global__ void __launch_bounds__(256) dct32_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int fwd) {
  pdl_entry();
  extern __shared__ __align__(16) float sm[];
  float* sD = sm;
  float* sDT = sm + 1024;
  float* img = sm + 2048;
  float* tmp = img + 1024 * C;
  const int n4 = 256 * C;
  dct_load_matrix(sD, sDT);
  const float4* src = reinterpret_cast<const float4*>(in + (long long)blockIdx.x * 1024 * C);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) reinterpret_cast<float4*>(img)[i] = src[i];
  __syncthreads();
  dct_planes(img, tmp, sD, sDT, C, fwd);
  float4* dst = reinterpret_cast<float4*>(out + (long long)blockIdx.x * 1024 * C);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = reinterpret_cast<const float4*>(img)[i];
}

int dct32_launch(const float* in, float* out, int B, int C, int fwd, cudaStream_t st) {
  if (ensure_dct_matrix()) return -1;
  const size_t smem = (size_t)(2048 + 2 * 1024 * C) * sizeof(float);
  if (smem > 48 * 1024) return -3;
  launch_k(dct32_kernel, dim3(B), dim3(256), smem, st, in, out, C, fwd);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// y' = a .* y + b .* DCT(eps_x);  x_next = IDCT(y').  One CTA per image; eps, y read once, y', x written once.
__global__ void __launch_bounds__(256) blur_step_kernel(const float* __restrict__ y, const float* __restrict__ eps_x,
                                                       const float* __restrict__ a, const float* __restrict__ bcoef,
                                                       float* __restrict__ y_out, float* __restrict__ x_next, int C) {
  pdl_entry();
  extern __shared__ __align__(16) float sm[];
  float* sD = sm;
  float* sDT = sm + 1024;
  float* img = sm + 2048;
  float* tmp = img + 1024 * C;
  const int n = 1024 * C;
  const long long off = (long long)blockIdx.x * n;
  dct_load_matrix(sD, sDT);
  for (int i = threadIdx.x; i < n / 4; i += blockDim.x)
    reinterpret_cast<float4*>(img)[i] = reinterpret_cast<const float4*>(eps_x + off)[i];
  __syncthreads();
  dct_planes(img, tmp, sD, sDT, C, 1);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int f = i / C;                            // frequency index h*32 + w
    const float v = a[f] * y[off + i] + bcoef[f] * img[i];
    y_out[off + i] = v;
    img[i] = v;
  }
  __syncthreads();
  if (x_next != nullptr) {
    dct_planes(img, tmp, sD, sDT, C, 0);
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x)
      reinterpret_cast<float4*>(x_next + off)[i] = reinterpret_cast<const float4*>(img)[i];
  }
}

int blur_step_launch(const float* y, const float* eps_x, const float* a, const float* b, float* y_out, float* x_next,
                     int B, int C, cudaStream_t st) {
  if (ensure_dct_matrix()) return -1;
  const size_t smem = (size_t)(2048 + 2 * 1024 * C) * sizeof(float);
  if (smem > 48 * 1024) return -3;
  launch_k(blur_step_kernel, dim3(B), dim3(256), smem, st, y, eps_x, a, b, y_out, x_next, C);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

__global__ void scale_shift_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, float mul,
                                   float add) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = in[i] * mul + add;
}
int scale_shift_launch(const float* in, float* out, long long n, float mul, float add, cudaStream_t st) {
  launch_k(scale_shift_kernel, dim3(grid_for(n, 256)), dim3(256), 0, st, in, out, n, mul, add);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

__global__ void scalar_ab_step_kernel(const float* __restrict__ x, const float* __restrict__ coef,
                                      const float* __restrict__ new_eps, const float* __restrict__ hist,
                                      float* __restrict__ x_out, float* __restrict__ hist_out, int n_hist, long long n) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float e0 = new_eps[i];
    float acc = coef[0] * x[i] + coef[1] * e0;
    if (n_hist > 0) hist_out[i] = e0;
    for (int j = 0; j < n_hist; ++j) {
      const float e = hist[(long long)j * n + i];
      acc += coef[2 + j] * e;
      if (j + 1 < n_hist) hist_out[(long long)(j + 1) * n + i] = e;
    }
    x_out[i] = acc;
  }
}
int scalar_ab_step_launch(const float* x, const float* coef, const float* new_eps, const float* hist, float* x_out,
                          float* hist_out, int n_hist, long long n, cudaStream_t st) {
  launch_k(scalar_ab_step_kernel, dim3(grid_for(n, 256)), dim3(256), 0, st, x, coef, new_eps, hist, x_out, hist_out, n_hist, n);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace gddim
