// Kernel launch helper: programmatic dependent launch (PDL) for the ~460-node kernel chain of one network evaluation.
//
// Every kernel of the library starts with pdl_entry(): `griddepcontrol.launch_dependents` lets the NEXT kernel of the
// stream (or captured graph) be scheduled as soon as all CTAs of this grid have started, so its launch latency, CTA
// ramp and prologue (barrier init, TMEM allocation, tensor-map prefetch) overlap this grid's tail; the following
// `griddepcontrol.wait` blocks until every prerequisite grid has COMPLETED and its memory is visible, so no kernel
// reads or overwrites anything before its producers/consumers are done (the workspace arena recycles buffers, so
// the wait also covers write-after-read).  Because every kernel waits before it can complete, completion is
// transitive along the chain.  Without the launch attribute both instructions are no-ops.
// Measured on B200 (tools/ab.sh, profiles/r01_ab_scheduling.txt): inside a captured graph a kernel node already costs
// only 1.5 us, and with the attribute it costs 1.7 us (empty-kernel chain of 540 nodes) -- the sampler is 2.4 % SLOWER
// with PDL (763.6 vs 745.0 ms per call, bit-identical samples).  It is therefore OFF by default; GDDIM_PDL=1 turns
// it on.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <utility>

namespace gddim {

inline int pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("GDDIM_PDL"); on = (e && e[0] == '1') ? 1 : 0; }
  return on;
}

// sets the programmatic-serialization attribute in slot `at[n]`; returns the new attribute count
inline int pdl_attr(cudaLaunchAttribute* at, int n) {
  if (!pdl_enabled()) return n;
  at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[n].val.programmaticStreamSerializationAllowed = 1;
  return n + 1;
}

// cudaFuncSetAttribute (max dynamic shared memory) and the SM count are PER DEVICE: one process may drive several
// devices through several contexts, so "done once" flags are kept per (kernel, device), not per process.
struct DeviceOnce {
  unsigned long long mask = 0;
  static int dev() { int d = 0; cudaGetDevice(&d); return d; }
  bool need() const { const int d = dev(); return d >= 64 || !((mask >> d) & 1ull); }
  void done() { const int d = dev(); if (d < 64) mask |= 1ull << d; }
};
inline int device_sm_count() {
  static int n[64] = {0};
  const int d = DeviceOnce::dev();
  if (d >= 64) { int v = 0; cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d); return v; }
  if (n[d] == 0) cudaDeviceGetAttribute(&n[d], cudaDevAttrMultiProcessorCount, d);
  return n[d];
}

template <typename... P, typename... A>
inline cudaError_t launch_k(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  cfg.attrs = at; cfg.numAttrs = pdl_attr(at, 0);
  return cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() { pdl_launch_dependents(); pdl_wait(); }
#endif

}  // namespace gddim
