// Fused forward + backward of the VQA answer-head losses over [B, A] logits (A = 2274 .. 3129):
// BCE (hg_transformers/modeling_lxmert.py:248-253), LPF (hg_transformers/mask_trainer_VQA.py:111-129)
// and LMH / LearnedMixin (hg_transformers/vqa_debias_loss_functions.py:148-196), plus the batch VQA
// score (hg_transformers/data/metrics/__init__.py:90-104).  One CTA per sample row; the row lives in
// shared memory so HBM sees each input once and dlogits is written once.  Row partials go to a
// workspace and are summed in a fixed order by a one-block finalize kernel (deterministic).
#include <cfloat>

#include "common.cuh"

namespace crv {

constexpr int kLossThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < kLossThreads / 32; ++w) t += red[w];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = -FLT_MAX;
  for (int w = 0; w < kLossThreads / 32; ++w) t = fmaxf(t, red[w]);
  return t;
}
// first index of the row maximum (torch.max returns the first occurrence on CUDA for ties is
// unspecified; the lowest index is used here and in the oracle)
__device__ __forceinline__ int block_argmax(float v, int idx, float* redv, int* redi) {
  for (int o = 16; o; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { redv[wid] = v; redi[wid] = idx; }
  __syncthreads();
  float bv = redv[0];
  int bi = redi[0];
  for (int w = 1; w < kLossThreads / 32; ++w)
    if (redv[w] > bv || (redv[w] == bv && redi[w] < bi)) { bv = redv[w]; bi = redi[w]; }
  return bi;
}

__device__ __forceinline__ float softplusf(float x) {  // torch F.softplus, beta = 1, threshold = 20
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

// workspace rows: [0,B) loss partial, [B,2B) score partial, [2B,3B) aux partial (LMH entropy)
__global__ void loss_finalize_kernel(const float* __restrict__ ws, int B, float scale_loss, float scale_aux,
                                     float* __restrict__ out) {
  __shared__ float red[kLossThreads / 32];
  float l = 0.f, s = 0.f, a = 0.f;
  for (int i = threadIdx.x; i < B; i += kLossThreads) { l += ws[i]; s += ws[B + i]; a += ws[2 * B + i]; }
  l = block_sum(l, red);
  s = block_sum(s, red);
  a = block_sum(a, red);
  if (threadIdx.x == 0) { out[0] = l * scale_loss + a * scale_aux; out[1] = s; }
}

__global__ void __launch_bounds__(kLossThreads)
bce_kernel(const float* __restrict__ logits, const float* __restrict__ labels, float* __restrict__ dlogits,
           float* __restrict__ ws, int B, int A) {
  __shared__ float red[kLossThreads / 32];
  __shared__ int redi[kLossThreads / 32];
  const int b = blockIdx.x;
  const float* x = logits + static_cast<size_t>(b) * A;
  const float* y = labels + static_cast<size_t>(b) * A;
  float* dx = dlogits + static_cast<size_t>(b) * A;
  const float invB = 1.f / B;
  float acc = 0.f, bestv = -FLT_MAX;
  int besti = 0x7FFFFFFF;
  for (int a = threadIdx.x; a < A; a += kLossThreads) {
    const float xv = x[a], yv = y[a];
    // binary_cross_entropy_with_logits: (1 - y) x + softplus(-x), in the max/log1p stable form
    acc += fmaxf(xv, 0.f) - xv * yv + log1pf(expf(-fabsf(xv)));
    dx[a] = (sigmoidf(xv) - yv) * invB;
    if (xv > bestv) { bestv = xv; besti = a; }
  }
  acc = block_sum(acc, red);
  const int am = block_argmax(bestv, besti, red, redi);
  if (threadIdx.x == 0) { ws[b] = acc; ws[B + b] = y[am]; ws[2 * B + b] = 0.f; }
}

__global__ void __launch_bounds__(kLossThreads)
lpf_kernel(const float* __restrict__ logits, const float* __restrict__ bias, const long long* __restrict__ max_label,
           const float* __restrict__ labels, float gamma, float* __restrict__ dlogits, float* __restrict__ ws, int B,
           int A) {
  extern __shared__ float row[];
  __shared__ float red[kLossThreads / 32];
  __shared__ int redi[kLossThreads / 32];
  const int b = blockIdx.x;
  const float* x = logits + static_cast<size_t>(b) * A;
  float* dx = dlogits + static_cast<size_t>(b) * A;
  float mx = -FLT_MAX, bestv = -FLT_MAX;
  int besti = 0x7FFFFFFF;
  for (int a = threadIdx.x; a < A; a += kLossThreads) {
    const float v = x[a];
    row[a] = v;
    mx = fmaxf(mx, v);
    if (v > bestv) { bestv = v; besti = a; }
  }
  mx = block_max(mx, red);
  const int am = block_argmax(bestv, besti, red, redi);
  float se = 0.f;
  for (int a = threadIdx.x; a < A; a += kLossThreads) {
    const float e = expf(row[a] - mx);
    row[a] = e;
    se += e;
  }
  se = block_sum(se, red);
  const long long yb = max_label[b];
  const float inv = 1.f / se;
  const float py = row[yb] * inv;
  const float q = fmaxf(bias[static_cast<size_t>(b) * A + yb], 1.0e-7f);
  const float feedback = expf(logf(q));
  const float wgt = powf(1.f - feedback, gamma);
  const bool live = py > 1.0e-7f;  // torch.max(p, 1e-7): gradient flows to p only when it wins
  const float scale = live ? wgt / B : 0.f;
  for (int a = threadIdx.x; a < A; a += kLossThreads) dx[a] = scale * (row[a] * inv - (a == yb ? 1.f : 0.f));
  if (threadIdx.x == 0) {
    ws[b] = wgt * -logf(fmaxf(py, 1.0e-7f));
    ws[B + b] = labels ? labels[static_cast<size_t>(b) * A + am] : 0.f;
    ws[2 * B + b] = 0.f;
  }
}

// RUBI_loss (hg_transformers/mask_trainer_VQA.py:131-135): cross entropy of the bias-gated logits
// z = logits * sigmoid(bias) against the majority label y; d loss / d logits = (softmax(z) - onehot(y)) sigmoid(bias) / B.
__global__ void __launch_bounds__(kLossThreads)
rubi_kernel(const float* __restrict__ logits, const float* __restrict__ bias, const long long* __restrict__ max_label,
            const float* __restrict__ labels, float* __restrict__ dlogits, float* __restrict__ ws, int B, int A) {
  extern __shared__ float row[];
  __shared__ float red[kLossThreads / 32];
  __shared__ int redi[kLossThreads / 32];
  const int b = blockIdx.x;
  const float* x = logits + static_cast<size_t>(b) * A;
  const float* bi = bias + static_cast<size_t>(b) * A;
  float* dx = dlogits + static_cast<size_t>(b) * A;
  float mx = -FLT_MAX, bestv = -FLT_MAX;
  int besti = 0x7FFFFFFF;
  for (int a = threadIdx.x; a < A; a += kLossThreads) {
    const float xv = x[a];
    const float z = xv * sigmoidf(bi[a]);
    row[a] = z;
    mx = fmaxf(mx, z);
    if (xv > bestv) { bestv = xv; besti = a; }     // the VQA score is taken on the raw logits (reference :839-842)
  }
  mx = block_max(mx, red);
  const int am = block_argmax(bestv, besti, red, redi);
  float se = 0.f;
  for (int a = threadIdx.x; a < A; a += kLossThreads) se += expf(row[a] - mx);
  se = block_sum(se, red);
  const long long yb = max_label[b];
  const float lse = mx + logf(se);
  const float invB = 1.f / B;
  for (int a = threadIdx.x; a < A; a += kLossThreads)
    dx[a] = (expf(row[a] - lse) - (a == yb ? 1.f : 0.f)) * sigmoidf(bi[a]) * invB;
  if (threadIdx.x == 0) {
    ws[b] = lse - row[yb];
    ws[B + b] = labels ? labels[static_cast<size_t>(b) * A + am] : 0.f;
    ws[2 * B + b] = 0.f;
  }
}

__global__ void __launch_bounds__(kLossThreads)
lmh_kernel(const float* __restrict__ logits, const float* __restrict__ bias, const float* __restrict__ labels,
           const float* __restrict__ factor_pre, float smooth, float w_ent, float* __restrict__ dlogits,
           float* __restrict__ dfactor_pre, float* __restrict__ ws, int B, int A) {
  __shared__ float red[kLossThreads / 32];
  __shared__ int redi[kLossThreads / 32];
  const int b = blockIdx.x;
  const float* x = logits + static_cast<size_t>(b) * A;
  const float* bi = bias + static_cast<size_t>(b) * A;
  const float* y = labels + static_cast<size_t>(b) * A;
  float* dx = dlogits + static_cast<size_t>(b) * A;
  const float z = factor_pre[b];
  const float f = softplusf(z);
  const float invB = 1.f / B;
  const float ent_scale = w_ent / (static_cast<float>(B) * static_cast<float>(A));
  float sum_prob = 0.f, ent = 0.f, dfa = 0.f, dfe = 0.f, bestv = -FLT_MAX;
  int besti = 0x7FFFFFFF;
  for (int a = threadIdx.x; a < A; a += kLossThreads) {
    const float xv = x[a], yv = y[a], be = bi[a];
    const float c0 = logf(be + smooth), c1 = logf(1.f - be + smooth);
    const float b0 = c0 * f, b1 = c1 * f;
    const float lp = -softplusf(-xv);   // log sigmoid(x)
    const float l1p = -xv + lp;         // log (1 - sigmoid(x))
    const float u0 = b0 + lp, u1 = b1 + l1p;
    const float norm = fmaxf(u0, u1) + log1pf(expf(-fabsf(u0 - u1)));
    const float LP = u0 - norm, L1P = u1 - norm;
    sum_prob += LP * yv + (1.f - yv) * L1P;
    const float g0 = (expf(LP) - yv) * invB;  // d loss / d u0 = -(d loss / d u1) = d loss / d logit
    dx[a] = g0;
    dfa += g0 * (c0 - c1);
    // entropy of the re-normalised bias
    const float bn = fmaxf(b0, b1) + log1pf(expf(-fabsf(b0 - b1)));
    const float q0 = b0 - bn, q1 = b1 - bn;
    const float p0 = expf(q0), p1 = expf(q1);
    const float H = -(p0 * q0 + p1 * q1);
    ent += H;
    dfe += -(p0 * (q0 + H) * c0 + p1 * (q1 + H) * c1);
    if (xv > bestv) { bestv = xv; besti = a; }
  }
  sum_prob = block_sum(sum_prob, red);
  ent = block_sum(ent, red);
  dfa = block_sum(dfa, red);
  dfe = block_sum(dfe, red);
  const int am = block_argmax(bestv, besti, red, redi);
  const bool bad = isnan(sum_prob);  // reference :183 zeroes NaN rows (and with them their gradient)
  if (bad)
    for (int a = threadIdx.x; a < A; a += kLossThreads) dx[a] = 0.f;
  if (threadIdx.x == 0) {
    ws[b] = bad ? 0.f : -sum_prob;
    ws[B + b] = y[am];
    ws[2 * B + b] = ent;
    dfactor_pre[b] = ((bad ? 0.f : dfa) + ent_scale * dfe) * sigmoidf(z);
  }
}

}  // namespace crv

using namespace crv;

extern "C" size_t crv_vqa_loss_workspace_bytes(int B) { return B > 0 ? sizeof(float) * 3 * B : 0; }

extern "C" int crv_vqa_loss_bce(const float* logits, const float* labels, float* loss_out, float* dlogits, int B, int A,
                                void* workspace, void* stream) {
  if (!logits || !labels || !loss_out || !dlogits || !workspace || B <= 0 || A <= 0) return CRV_E_BADARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(workspace);
  bce_kernel<<<B, kLossThreads, 0, st>>>(logits, labels, dlogits, ws, B, A);
  int rc = launch_status();
  if (rc) return rc;
  loss_finalize_kernel<<<1, kLossThreads, 0, st>>>(ws, B, 1.f / B, 0.f, loss_out);
  return launch_status();
}

extern "C" int crv_vqa_loss_lpf(const float* logits, const float* bias, const long long* max_label, float gamma,
                                float* loss_out, const float* labels, float* dlogits, int B, int A, void* workspace,
                                void* stream) {
  if (!logits || !bias || !max_label || !loss_out || !dlogits || !workspace || B <= 0 || A <= 0) return CRV_E_BADARG;
  if (static_cast<size_t>(A) * sizeof(float) > 200 * 1024) return CRV_E_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(workspace);
  const size_t smem = static_cast<size_t>(A) * sizeof(float);
  if (smem > 48 * 1024) {
    static bool configured = false;
    if (!configured) {
      CRV_CUDA(cudaFuncSetAttribute(lpf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = true;
    }
  }
  lpf_kernel<<<B, kLossThreads, smem, st>>>(logits, bias, max_label, labels, gamma, dlogits, ws, B, A);
  int rc = launch_status();
  if (rc) return rc;
  loss_finalize_kernel<<<1, kLossThreads, 0, st>>>(ws, B, 1.f / B, 0.f, loss_out);
  return launch_status();
}

extern "C" int crv_vqa_loss_rubi(const float* logits, const float* bias, const long long* max_label, float* loss_out,
                                 const float* labels, float* dlogits, int B, int A, void* workspace, void* stream) {
  if (!logits || !bias || !max_label || !loss_out || !dlogits || !workspace || B <= 0 || A <= 0) return CRV_E_BADARG;
  if (static_cast<size_t>(A) * sizeof(float) > 200 * 1024) return CRV_E_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(workspace);
  const size_t smem = static_cast<size_t>(A) * sizeof(float);
  if (smem > 48 * 1024) {
    static bool configured = false;
    if (!configured) {
      CRV_CUDA(cudaFuncSetAttribute(rubi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = true;
    }
  }
  rubi_kernel<<<B, kLossThreads, smem, st>>>(logits, bias, max_label, labels, dlogits, ws, B, A);
  int rc = launch_status();
  if (rc) return rc;
  loss_finalize_kernel<<<1, kLossThreads, 0, st>>>(ws, B, 1.f / B, 0.f, loss_out);
  return launch_status();
}

extern "C" int crv_vqa_loss_lmh(const float* logits, const float* bias, const float* labels, const float* factor_pre,
                                float smooth, float w, float* loss_out, float* dlogits, float* dfactor_pre, int B,
                                int A, void* workspace, void* stream) {
  if (!logits || !bias || !labels || !factor_pre || !loss_out || !dlogits || !dfactor_pre || !workspace || B <= 0 ||
      A <= 0)
    return CRV_E_BADARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(workspace);
  lmh_kernel<<<B, kLossThreads, 0, st>>>(logits, bias, labels, factor_pre, smooth, w, dlogits, dfactor_pre, ws, B, A);
  int rc = launch_status();
  if (rc) return rc;
  loss_finalize_kernel<<<1, kLossThreads, 0, st>>>(ws, B, 1.f / B, w / (static_cast<float>(B) * A), loss_out);
  return launch_status();
}
