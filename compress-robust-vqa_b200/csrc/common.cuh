// Shared host-side helpers for the C-ABI translation units.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/crvqa.h"

namespace crv {

extern thread_local int g_last_cuda_error;
extern unsigned long long g_launch_count;  // kernels launched by this library (all threads; racy by design)

inline int record(cudaError_t e) {
  if (e != cudaSuccess) {
    g_last_cuda_error = static_cast<int>(e);
    return static_cast<int>(e);
  }
  return CRV_OK;
}

#define CRV_CUDA(expr)                        \
  do {                                        \
    int _rc = ::crv::record((expr));          \
    if (_rc != CRV_OK) return _rc;            \
  } while (0)

inline int launch_status(int kernels = 1) {
  g_launch_count += static_cast<unsigned long long>(kernels);
  return record(cudaGetLastError());
}

// Programmatic dependent launch.  Every kernel of the step is a short link of one long dependency chain, so the
// launch latency and prologue of link i+1 are overlapped with the tail of link i: kernels are launched with the
// programmatic-stream-serialization attribute, call pdl_wait() before they touch anything a predecessor wrote
// (it returns at once when the launch carried no programmatic edge) and pdl_launch_dependents() right away, which
// lets the next grid be scheduled as soon as all of this grid's CTAs are resident.  pdl_wait() is the first
// statement of every such kernel -- before any early return -- so completion stays transitive along the chain.
// CRVQA_PDL=0 launches without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("CRVQA_PDL");
    return e ? atoi(e) != 0 : true;
  }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace crv
