// erf GELU device functions shared by the standalone GELU kernels (fused_ops.cu) and the masked-GEMM epilogues
// (gemm_sm100.cu: FF1 forward writes gelu(u) next to u; the dX of FF2 multiplies by gelu'(u)).
#pragma once
#include <cuda_runtime.h>

namespace crv {

// erf GELU (the reference's: 0.5 u (1 + erf(u / sqrt 2))).  libdevice erff + expf make these kernels ALU-bound
// (measured 45 % of HBM bandwidth); Abramowitz-Stegun 7.1.26 needs one exp2 and one reciprocal on the SFU plus
// six FMAs, and the same e^(-u^2/2) serves the density term of the derivative.  |erf error| < 1.5e-7, far
// below the bf16 rounding of the result; the lower tail is formed directly (0.5 erfc), without cancellation.
__device__ __forceinline__ float sfu_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sfu_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// tail = 0.5 erfc(|u| / sqrt 2) = P(N(0,1) > |u|),  e = exp(-u^2 / 2)       (13 instructions, 2 of them SFU)
__device__ __forceinline__ void normal_tail(float u, float& tail, float& e) {
  const float t = sfu_rcp(fmaf(0.3275911f * 0.70710678118654752f, fabsf(u), 1.f));
  e = sfu_ex2(u * u * -0.72134752044448170f);         // -0.5 log2(e)
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  tail = poly * t * e;
}
__device__ __forceinline__ float gelu_f(float u) {    // u Phi(u) = max(u, 0) - |u| tail
  float tail, e;
  normal_tail(u, tail, e);
  return fmaf(-fabsf(u), tail, fmaxf(u, 0.f));
}
__device__ __forceinline__ float gelu_grad_f(float u) {   // Phi(u) + u phi(u)
  float tail, e;
  normal_tail(u, tail, e);
  const float cdf = u < 0.f ? tail : 1.f - tail;
  return fmaf(u * 0.3989422804014327f, e, cdf);
}

}  // namespace crv
