// The gDDIM / DEIS linear-algebra update on ONE pixel of a CLD state in net layout (C = 3: six floats x0 x1 x2 v0 v1 v2):
//   u' = A u + sum_j C_j eps_j      with 2x2 matrices acting on every (x_d, v_d) pair
// (cld_jax/deis.py:141-151 multistep_ab_step; cld_jax/sampling.py:30-39 denoising step; models/utils.py:174-176 mixed
// score).  Shared by the standalone update kernel (update.cu) and by the head convolution's epilogue
// (gemm_epilogue.cuh: epi_head_update), which applies the same update to the eps values it still holds in registers.
// Every operation is an explicit round-to-nearest intrinsic, so both callers produce bit-identical states whatever the
// compiler would otherwise contract.
#pragma once

namespace gddim {

// what the head convolution's epilogue needs to apply the update (kernel parameter, by value)
struct CldUpd {
  const float* u;          // [n_pix, 6] state in net layout
  float* u_out;            // may alias u (each thread reads its pixel before it writes it)
  const float* eps[4];     // eps[1..n_eps-1]: earlier evaluations (ring slots); eps[0] is the value in registers
  int n_eps;               // 1 .. 4
  int mixed;               // eps_0 += M u before it is stored and used
  float coef[5][4];        // coef[0] = A, coef[1 + j] = C_j (row-major 2x2)
  float mixm[4];
};

__device__ __forceinline__ void cld_px_apply(float (&acc)[6], const float (&u)[6], float a00, float a01, float a10, float a11) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    acc[d] = __fmaf_rn(a00, u[d], __fmul_rn(a01, u[3 + d]));
    acc[3 + d] = __fmaf_rn(a10, u[d], __fmul_rn(a11, u[3 + d]));
  }
}
// e += M u
__device__ __forceinline__ void cld_px_mix(float (&e)[6], const float (&u)[6], float m00, float m01, float m10, float m11) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float ex = __fadd_rn(e[d], __fmaf_rn(m00, u[d], __fmul_rn(m01, u[3 + d])));
    const float ev = __fadd_rn(e[3 + d], __fmaf_rn(m10, u[d], __fmul_rn(m11, u[3 + d])));
    e[d] = ex; e[3 + d] = ev;
  }
}
// acc += C e
__device__ __forceinline__ void cld_px_acc(float (&acc)[6], const float (&e)[6], float c00, float c01, float c10, float c11) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    acc[d] = __fadd_rn(acc[d], __fmaf_rn(c00, e[d], __fmul_rn(c01, e[3 + d])));
    acc[3 + d] = __fadd_rn(acc[3 + d], __fmaf_rn(c10, e[d], __fmul_rn(c11, e[3 + d])));
  }
}

}  // namespace gddim
