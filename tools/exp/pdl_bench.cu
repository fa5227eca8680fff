// Experiment: per-node cost of a captured chain of dependent kernels with the launch shapes of one network
// evaluation (persistent GEMM CTAs with ~200 KB smem alternating with GroupNorm-shaped grids), with and without
// programmatic dependent launch.  Kernels do ~no work: what is measured is launch + drain + ramp per node.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__global__ void __launch_bounds__(192, 1) big_kernel(float* p, int pdl) {
  extern __shared__ unsigned char sm[];
  if (pdl) { asm volatile("griddepcontrol.launch_dependents;"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
  if (threadIdx.x == 0) { sm[0] = 1; p[blockIdx.x] += 1.f; }
}
__global__ void __launch_bounds__(256) small_kernel(float* p, int pdl) {
  if (pdl) { asm volatile("griddepcontrol.launch_dependents;"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
  if (threadIdx.x == 0) p[blockIdx.x + blockIdx.y * gridDim.x] += 1.f;
}

static void launch(void* fn, dim3 g, dim3 b, size_t smem, cudaStream_t st, float* p, int pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = g; cfg.blockDim = b; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  void* args[] = {&p, &pdl};
  cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
  if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); }
}

int main() {
  float* p; cudaMalloc(&p, 1 << 20); cudaMemset(p, 0, 1 << 20);
  cudaFuncSetAttribute(big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  for (int pdl = 0; pdl < 2; ++pdl) {
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    const int pairs = 270;
    for (int i = 0; i < pairs; ++i) {
      launch((void*)big_kernel, dim3(148), dim3(192), 200 * 1024, st, p, pdl);
      launch((void*)small_kernel, dim3(8, 256), dim3(256), 0, st, p, pdl);
    }
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (e != cudaSuccess) { printf("capture: %s\n", cudaGetErrorString(e)); return 1; }
    e = cudaGraphInstantiate(&ge, g, 0);
    if (e != cudaSuccess) { printf("instantiate: %s\n", cudaGetErrorString(e)); return 1; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) cudaGraphLaunch(ge, st);
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    const int reps = 20;
    for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("pdl=%d: %.3f ms per graph of %d nodes = %.2f us per node (%s)\n", pdl, ms / reps, 2 * pairs,
           1e3 * ms / reps / (2 * pairs), cudaGetErrorString(cudaGetLastError()));
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  }
  return 0;
}
