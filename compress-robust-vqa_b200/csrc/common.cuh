// Shared host-side helpers for the C-ABI translation units.
#pragma once
#include <cstdint>
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
