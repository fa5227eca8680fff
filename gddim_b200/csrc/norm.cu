// K2 / K2b: GroupNorm (+ swish) (+ FIR or naive 2x resampling) producing the fp16 GEMM operand.
//
// Replaces flax nn.GroupNorm + nn.swish + up_or_down_sampling.{upsample_2d,downsample_2d,naive_*}
// as used by ResnetBlockBigGANpp (cld_jax/models/layerspp.py:196-213,218) and AttnBlockpp (layerspp.py:69),
// ncsnpp.py:236.  Statistics: per (image, group) over (H, W, C/G), variance = E[x^2] - E[x]^2,
// eps = 1e-6 (flax default), groups are contiguous channel blocks.
//
// Memory-bound, no tensor-core path: pass 1 reads the fp32 source once (float4, coalesced) and writes
// deterministic per-slab partial sums; pass 2 reads it again, applies scale/shift/swish in registers,
// optionally gathers the 4x4 (down) or 2x2 (up) FIR footprint, and writes 16-byte fp16 vectors.
#include <cstdio>

#include "kernels.h"

namespace gddim {

static int ceil_div(long long a, long long b) { return int((a + b - 1) / b); }

int norm_splits(int B, int H, int W) {
  const int P = H * W;
  int s = 1;
  // aim for >= 2 waves of 148 SMs, keep >= 16 pixels per slab
  while (s < 32 && (long long)B * s < 592 && P / (s * 2) >= 16) s *= 2;
  return s;
}

__global__ void gn_partial_kernel(const float* __restrict__ src1, int c1, const float* __restrict__ src2, int c2,
                                  int P, int groups, int splits, int rows, float* __restrict__ partial) {
  extern __shared__ float sm[];   // [threads][8]: per-channel sum[4], sumsq[4]
  const int C = c1 + c2;
  const int nv = C / 4;
  const int b = blockIdx.y, split = blockIdx.x;
  const int vi = threadIdx.x % nv, row = threadIdx.x / nv;
  const int pp = P / splits;
  const int pbeg = split * pp;
  const int c = vi * 4;
  const float* base;
  int cs, cc;
  if (c < c1) { base = src1 + (long long)b * P * c1; cs = c1; cc = c; }
  else { base = src2 + (long long)b * P * c2; cs = c2; cc = c - c1; }
  float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
  for (int pix = pbeg + row; pix < pbeg + pp; pix += rows) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(base + (long long)pix * cs + cc));
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    ss[0] += v.x * v.x; ss[1] += v.y * v.y; ss[2] += v.z * v.z; ss[3] += v.w * v.w;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sm[threadIdx.x * 8 + j] = s[j];
    sm[threadIdx.x * 8 + 4 + j] = ss[j];
  }
  __syncthreads();
  // fixed-order (deterministic) reduction: one thread per group walks its channels and pixel rows
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    const int cpg = C / groups;
    float ts = 0.f, tss = 0.f;
    for (int ch = g * cpg; ch < (g + 1) * cpg; ++ch) {
      const int v = ch >> 2, j = ch & 3;
      for (int r = 0; r < rows; ++r) {
        const int t = r * nv + v;
        ts += sm[t * 8 + j];
        tss += sm[t * 8 + 4 + j];
      }
    }
    float* o = partial + (((long long)b * splits + split) * groups + g) * 2;
    o[0] = ts;
    o[1] = tss;
  }
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

struct ApplyArgs {
  const float* src1; int c1;
  const float* src2; int c2;
  int H, W, Ho, Wo;
  int groups, splits;
  const float* gamma; const float* beta;
  const float* partial;
  float eps;
  int silu, resample, do_norm;
  float raw_scale;
  __half* dst16; __half* raw16;
};

__device__ __forceinline__ void store8(__half* p, const float (&v)[8]) {
  __half2 h0 = __floats2half2_rn(v[0], v[1]);
  __half2 h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]);
  __half2 h3 = __floats2half2_rn(v[6], v[7]);
  uint4 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&h0);
  pk.y = *reinterpret_cast<uint32_t*>(&h1);
  pk.z = *reinterpret_cast<uint32_t*>(&h2);
  pk.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(p) = pk;
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const ApplyArgs p) {
  extern __shared__ float sm[];   // a[C], b[C]
  const int C = p.c1 + p.c2;
  float* sa = sm;
  float* sb = sm + C;
  const int b = blockIdx.y;
  if (p.do_norm) {
    const int cpg = C / p.groups;
    const float inv_n = 1.0f / (float(p.H) * float(p.W) * float(cpg));
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const int g = c / cpg;
      float s = 0.f, ss = 0.f;
      for (int k = 0; k < p.splits; ++k) {
        const float* q = p.partial + (((long long)b * p.splits + k) * p.groups + g) * 2;
        s += q[0];
        ss += q[1];
      }
      const float mean = s * inv_n;
      const float var = fmaxf(ss * inv_n - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
      const float a = rstd * p.gamma[c];
      sa[c] = a;
      sb[c] = p.beta[c] - mean * a;
    }
    __syncthreads();
  }
  const int nv = C / 8;
  const long long total = (long long)p.Ho * p.Wo * nv;
  const float* base1 = p.src1 + (long long)b * p.H * p.W * p.c1;
  const float* base2 = p.src2 ? p.src2 + (long long)b * p.H * p.W * p.c2 : nullptr;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int vi = int(idx % nv);
    const int opix = int(idx / nv);
    const int ox = opix % p.Wo, oy = opix / p.Wo;
    const int c = vi * 8;
    const float* base; int cs, cc;
    if (c < p.c1) { base = base1; cs = p.c1; cc = c; } else { base = base2; cs = p.c2; cc = c - p.c1; }
    float a[8], bb[8];
    if (p.do_norm) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { a[j] = sa[c + j]; bb[j] = sb[c + j]; }
    }
    float accn[8], accr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { accn[j] = 0.f; accr[j] = 0.f; }

    // tap enumeration
    int ny, nx, y0, x0;
    float wy[4], wx[4];
    switch (p.resample) {
      case RS_FIR_DOWN:
        ny = nx = 4; y0 = 2 * oy - 1; x0 = 2 * ox - 1;
        wy[0] = wx[0] = 0.125f; wy[1] = wx[1] = 0.375f; wy[2] = wx[2] = 0.375f; wy[3] = wx[3] = 0.125f;
        break;
      case RS_FIR_UP:
        ny = nx = 2;
        if (oy & 1) { y0 = oy >> 1; wy[0] = 0.75f; wy[1] = 0.25f; } else { y0 = (oy >> 1) - 1; wy[0] = 0.25f; wy[1] = 0.75f; }
        if (ox & 1) { x0 = ox >> 1; wx[0] = 0.75f; wx[1] = 0.25f; } else { x0 = (ox >> 1) - 1; wx[0] = 0.25f; wx[1] = 0.75f; }
        break;
      case RS_NAIVE_DOWN:
        ny = nx = 2; y0 = 2 * oy; x0 = 2 * ox; wy[0] = wy[1] = wx[0] = wx[1] = 0.5f;
        break;
      case RS_NAIVE_UP:
        ny = nx = 1; y0 = oy >> 1; x0 = ox >> 1; wy[0] = wx[0] = 1.f;
        break;
      default:
        ny = nx = 1; y0 = oy; x0 = ox; wy[0] = wx[0] = 1.f;
        break;
    }
    for (int i = 0; i < ny; ++i) {
      const int iy = y0 + i;
      if (iy < 0 || iy >= p.H) continue;
      for (int j = 0; j < nx; ++j) {
        const int ix = x0 + j;
        if (ix < 0 || ix >= p.W) continue;
        const float w = wy[i] * wx[j];
        const float* q = base + ((long long)iy * p.W + ix) * cs + cc;
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(q));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(q + 4));
        const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          accr[k] += w * x[k];
          if (p.do_norm) {
            float t = x[k] * a[k] + bb[k];
            if (p.silu) t = silu_f(t);
            accn[k] += w * t;
          }
        }
      }
    }
    const long long o = (((long long)b * p.Ho + oy) * p.Wo + ox) * C + c;
    if (p.dst16) store8(p.dst16 + o, accn);
    if (p.raw16) {
#pragma unroll
      for (int k = 0; k < 8; ++k) accr[k] *= p.raw_scale;
      store8(p.raw16 + o, accr);
    }
  }
}

int norm_launch(const NormOp* op, cudaStream_t st) {
  const int C = op->c1 + op->c2;
  const int do_norm = op->dst16 != nullptr;
  if (C % 8 != 0 || op->c1 % 8 != 0) return -1;
  if (do_norm) {
    if (C % op->groups != 0 || C % 4 != 0) return -2;
    const int nv = C / 4;
    int rows = 256 / nv;
    if (rows < 1) rows = 1;
    const int P = op->H * op->W;
    const int pp = P / op->splits;
    if (rows > pp) rows = pp;
    if (nv * rows > 1024 || P % op->splits != 0) return -3;
    const int threads = nv * rows;
    dim3 grid(op->splits, op->B);
    gn_partial_kernel<<<grid, threads, threads * 8 * sizeof(float), st>>>(op->src1, op->c1, op->src2, op->c2, P,
                                                                          op->groups, op->splits, rows, op->partial);
  }
  ApplyArgs a;
  a.src1 = op->src1; a.c1 = op->c1; a.src2 = op->src2; a.c2 = op->c2;
  a.H = op->H; a.W = op->W;
  switch (op->resample) {
    case RS_FIR_DOWN: case RS_NAIVE_DOWN: a.Ho = op->H / 2; a.Wo = op->W / 2; break;
    case RS_FIR_UP: case RS_NAIVE_UP: a.Ho = op->H * 2; a.Wo = op->W * 2; break;
    default: a.Ho = op->H; a.Wo = op->W; break;
  }
  a.groups = op->groups; a.splits = op->splits; a.gamma = op->gamma; a.beta = op->beta; a.partial = op->partial;
  a.eps = op->eps; a.silu = op->silu; a.resample = op->resample; a.do_norm = do_norm;
  a.dst16 = op->dst16; a.raw16 = op->raw16;
  a.raw_scale = op->raw_scale;
  const long long total = (long long)a.Ho * a.Wo * (C / 8);
  int gx = ceil_div(total, 256 * 2);
  if (gx < 1) gx = 1;
  dim3 grid(gx, op->B);
  gn_apply_kernel<<<grid, 256, 2 * C * sizeof(float), st>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace gddim
