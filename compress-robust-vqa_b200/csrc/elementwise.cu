// Streaming (HBM-bound) kernels: operand casts, binariser, mask application, magnitude init,
// gradient-norm partial sums and the fused clip+AdamW update.  All are grid-stride loops over
// 16-byte vectors, launched with a multiple of the SM count.
#include <cmath>

#include "common.cuh"

namespace crv {

thread_local int g_last_cuda_error = 0;
unsigned long long g_launch_count = 0;

constexpr int kThreads = 256;

static inline int stream_grid(int64_t nvec, int per_sm = 8) {
  int64_t blocks = (nvec + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(num_sms()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t n) {
  const int64_t nvec = n >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    reinterpret_cast<uint4*>(dst)[i] =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const int64_t i = (nvec << 3) + threadIdx.x;
    __nv_bfloat16 t = __float2bfloat16_rn(src[i]);
    dst[i] = *reinterpret_cast<uint16_t*>(&t);
  }
}

// stage-3 pruned operand: dst = bf16(w * m), the product formed in fp32 exactly as torch.nn.utils.prune's
// forward pre-hook does (weight = weight_orig * weight_mask) before the operand is rounded for the MMA
__global__ void mul_cast_bf16_kernel(const float* __restrict__ w, const float* __restrict__ m, uint16_t* __restrict__ dst,
                                     int64_t n) {
  const int64_t nvec = n >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(w) + 2 * i + 1);
    const float4 ma = __ldg(reinterpret_cast<const float4*>(m) + 2 * i);
    const float4 mb = __ldg(reinterpret_cast<const float4*>(m) + 2 * i + 1);
    reinterpret_cast<uint4*>(dst)[i] = make_uint4(pack_bf16(a.x * ma.x, a.y * ma.y), pack_bf16(a.z * ma.z, a.w * ma.w),
                                                  pack_bf16(b.x * mb.x, b.y * mb.y), pack_bf16(b.z * mb.z, b.w * mb.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const int64_t i = (nvec << 3) + threadIdx.x;
    __nv_bfloat16 t = __float2bfloat16_rn(w[i] * m[i]);
    dst[i] = *reinterpret_cast<uint16_t*>(&t);
  }
}

__global__ void binarize_kernel(const float* __restrict__ s, const float* __restrict__ thr_p,
                                float* __restrict__ mf, uint8_t* __restrict__ mb, long long* __restrict__ kept,
                                int64_t n) {
  const float thr = __ldg(thr_p);
  const int64_t nvec = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  long long local = 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(s) + i);
    const bool k0 = v.x > thr, k1 = v.y > thr, k2 = v.z > thr, k3 = v.w > thr;
    if (mf) reinterpret_cast<float4*>(mf)[i] = make_float4(k0 ? 1.f : 0.f, k1 ? 1.f : 0.f, k2 ? 1.f : 0.f, k3 ? 1.f : 0.f);
    if (mb) reinterpret_cast<uchar4*>(mb)[i] = make_uchar4(k0, k1, k2, k3);
    local += int(k0) + int(k1) + int(k2) + int(k3);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (nvec << 2) + threadIdx.x;
    const bool k = s[i] > thr;
    if (mf) mf[i] = k ? 1.f : 0.f;
    if (mb) mb[i] = k;
    local += int(k);
  }
  if (kept) {
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ long long wsum[kThreads / 32];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      long long t = 0;
      for (int w = 0; w < kThreads / 32; ++w) t += wsum[w];
      if (t) atomicAdd(reinterpret_cast<unsigned long long*>(kept), static_cast<unsigned long long>(t));
    }
  }
}

__global__ void apply_mask_kernel(const uint16_t* __restrict__ w, const float* __restrict__ s,
                                  const float* __restrict__ thr_p, uint16_t* __restrict__ wm, int64_t n) {
  const float thr = __ldg(thr_p);
  const int64_t nvec = n >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(s) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(s) + 2 * i + 1);
    uint4 v = __ldg(reinterpret_cast<const uint4*>(w) + i);
    v.x &= (a.x > thr ? 0x0000FFFFu : 0u) | (a.y > thr ? 0xFFFF0000u : 0u);
    v.y &= (a.z > thr ? 0x0000FFFFu : 0u) | (a.w > thr ? 0xFFFF0000u : 0u);
    v.z &= (b.x > thr ? 0x0000FFFFu : 0u) | (b.y > thr ? 0xFFFF0000u : 0u);
    v.w &= (b.z > thr ? 0x0000FFFFu : 0u) | (b.w > thr ? 0xFFFF0000u : 0u);
    reinterpret_cast<uint4*>(wm)[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const int64_t i = (nvec << 3) + threadIdx.x;
    wm[i] = s[i] > thr ? w[i] : uint16_t(0);
  }
}

// Wm = W (.) (S > thr[seg]) over a whole score arena in one launch.  `chunks` holds (start, len, seg)
// triples (element units, start % 8 == 0, no chunk straddles a module) built once by the host.
__global__ void apply_mask_segmented_kernel(const uint16_t* __restrict__ w, const float* __restrict__ s,
                                            const float* __restrict__ thr_vec, const int4* __restrict__ chunks,
                                            int nchunks, uint16_t* __restrict__ wm) {
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int4 ch = __ldg(chunks + c);
    const float thr = __ldg(thr_vec + ch.z);
    const int64_t base = static_cast<int64_t>(ch.x) * 8;  // start is stored in units of 8 elements
    const int nvec = ch.y >> 3;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      const int64_t e = base + static_cast<int64_t>(i) * 8;
      const float4 a = __ldg(reinterpret_cast<const float4*>(s + e));
      const float4 b = __ldg(reinterpret_cast<const float4*>(s + e) + 1);
      uint4 v = __ldg(reinterpret_cast<const uint4*>(w + e));
      v.x &= (a.x > thr ? 0x0000FFFFu : 0u) | (a.y > thr ? 0xFFFF0000u : 0u);
      v.y &= (a.z > thr ? 0x0000FFFFu : 0u) | (a.w > thr ? 0xFFFF0000u : 0u);
      v.z &= (b.x > thr ? 0x0000FFFFu : 0u) | (b.y > thr ? 0xFFFF0000u : 0u);
      v.w &= (b.z > thr ? 0x0000FFFFu : 0u) | (b.w > thr ? 0xFFFF0000u : 0u);
      *reinterpret_cast<uint4*>(wm + e) = v;
    }
    for (int i = (nvec << 3) + threadIdx.x; i < ch.y; i += blockDim.x)
      wm[base + i] = s[base + i] > thr ? w[base + i] : uint16_t(0);
  }
}

__global__ void magnitude_init_kernel(const float* __restrict__ w, const float* __restrict__ thr_p, float hi,
                                      float lo, float* __restrict__ s, int64_t n) {
  const float thr = __ldg(thr_p);
  const int64_t nvec = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(w) + i);
    reinterpret_cast<float4*>(s)[i] = make_float4(fabsf(v.x) > thr ? hi : lo, fabsf(v.y) > thr ? hi : lo,
                                                  fabsf(v.z) > thr ? hi : lo, fabsf(v.w) > thr ? hi : lo);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (nvec << 2) + threadIdx.x;
    s[i] = fabsf(w[i]) > thr ? hi : lo;
  }
}

// Block partials -> *out in a FIXED order: every rank of a data-parallel job must get the same bits (the clip
// coefficient feeds every update; a float atomicAdd per block would make it depend on block scheduling).  Each block
// publishes its partial in `ws`; the block that arrives last (ticket in ws[0]) adds them up in index order.
__device__ __forceinline__ void sumsq_finish(float acc, float* __restrict__ ws, float* __restrict__ out) {
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float wsum[kThreads / 32];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += wsum[w];
    ws[1 + blockIdx.x] = t;
    __threadfence();
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(ws), 1u);
    last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    // fixed-shape tree over the partials: thread i sums i, i + 256, ... ; then the block tree above
    float t = 0.f;
    for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) t += __ldcg(ws + 1 + i);
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float total = 0.f;
      for (int w = 0; w < kThreads / 32; ++w) total += wsum[w];
      *out += total;
      *reinterpret_cast<unsigned*>(ws) = 0u;      // ready for the next launch
    }
  }
}

__global__ void sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out, float* __restrict__ ws) {
  const int64_t nvec = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = x[(nvec << 2) + threadIdx.x];
    acc += v * v;
  }
  sumsq_finish(acc, ws, out);
}

// sum of squares over the chunks of a table (the gradient shard a rank owns under the sharded optimiser)
__global__ void sumsq_segmented_kernel(const float* __restrict__ x, const int4* __restrict__ chunks, int nchunks,
                                       float* __restrict__ out, float* __restrict__ ws) {
  float acc = 0.f;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int4 ch = __ldg(chunks + c);
    const int64_t base = static_cast<int64_t>(ch.x) * 8;
    const int nvec = ch.y >> 2;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + base) + i);
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int i = (nvec << 2) + threadIdx.x; i < ch.y; i += blockDim.x) acc += x[base + i] * x[base + i];
  }
  sumsq_finish(acc, ws, out);
}

// mode 0: the reference's root optimization.AdamW (stage 2): step_size = lr sqrt(1 - b2^t) / (1 - b1^t) computed by the
//         caller, denom = sqrt(v) + eps, decoupled weight decay, running sum of |g|.
// mode 1: torch.optim.Adam as the stage-3 driver builds it (run_vqa_stage3.py:577-598): L2 weight decay folded into g,
//         step_size = lr / (1 - b1^t), denom = sqrt(v) / sqrt(1 - b2^t) + eps  (inv_bc2_sqrt = 1 / sqrt(1 - b2^t)).
// mode 2: torch.optim.AdamW as mPLUG's driver builds it (mPLUG/optim/optim_factory.py:60-89): p *= 1 - lr * wd first
//         (`decay`, rounded once from double as torch rounds the Python scalar), then mode 1's moments and step.
struct AdamArgs {
  float lr, step_size, beta1, beta2, eps, weight_decay, max_norm, inv_bc2_sqrt;
  int mode;
  float decay;
  // 1 - beta rounded ONCE from double, as torch rounds the Python scalars `1.0 - beta` of add_(alpha=) / lerp_ /
  // addcmul_(value=): float(1 - 0.999) = 0.001f, whereas 1.0f - 0.999f = 0.00100005f (4.7e-5 off in exp_avg_sq)
  float omb1, omb2;
};
static AdamArgs make_adam_args(float lr, float step_size, double beta1, double beta2, float eps, float weight_decay,
                               float max_norm, float inv_bc2_sqrt, int mode) {
  return AdamArgs{lr, step_size, static_cast<float>(beta1), static_cast<float>(beta2), eps, weight_decay, max_norm,
                  inv_bc2_sqrt, mode, 1.0f, static_cast<float>(1.0 - beta1), static_cast<float>(1.0 - beta2)};
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float* sum, const AdamArgs& a,
                                         float clip) {
  g *= clip;
  if (a.mode == 0) {
    if (sum) *sum += fabsf(g);
    m = m * a.beta1 + a.omb1 * g;
    v = v * a.beta2 + a.omb2 * g * g;
    const float denom = sqrtf(v) + a.eps;
    p = p - a.step_size * (m / denom);
    if (a.weight_decay > 0.f) p = p - a.lr * a.weight_decay * p;
  } else {
    if (a.mode == 2) p = p * a.decay;
    else if (a.weight_decay > 0.f) g = fmaf(a.weight_decay, p, g);
    m = m + (g - m) * a.omb1;
    v = v * a.beta2 + a.omb2 * g * g;
    const float denom = sqrtf(v) * a.inv_bc2_sqrt + a.eps;
    p = p - a.step_size * (m / denom);
  }
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, float* __restrict__ sum, int64_t n, AdamArgs a,
                             const float* __restrict__ total_sumsq, const float* __restrict__ hyper) {
  if (hyper) {  // {lr, step_size, 1/sqrt(1-b2^t)} live in device memory so a captured CUDA graph follows the schedule
    a.lr = __ldg(hyper);
    a.step_size = __ldg(hyper + 1);
    a.inv_bc2_sqrt = __ldg(hyper + 2);
  }
  float clip = 1.0f;
  if (total_sumsq) {
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float c = a.max_norm / (sqrtf(__ldg(total_sumsq)) + 1e-6f);
    clip = c < 1.0f ? c : 1.0f;
  }
  const int64_t nvec = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float4 ss = sum ? reinterpret_cast<float4*>(sum)[i] : make_float4(0, 0, 0, 0);
    adam_one(pp.x, gg.x, mm.x, vv.x, sum ? &ss.x : nullptr, a, clip);
    adam_one(pp.y, gg.y, mm.y, vv.y, sum ? &ss.y : nullptr, a, clip);
    adam_one(pp.z, gg.z, mm.z, vv.z, sum ? &ss.z : nullptr, a, clip);
    adam_one(pp.w, gg.w, mm.w, vv.w, sum ? &ss.w : nullptr, a, clip);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (sum) reinterpret_cast<float4*>(sum)[i] = ss;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (nvec << 2) + threadIdx.x;
    adam_one(p[i], g[i], m[i], v[i], sum ? sum + i : nullptr, a, clip);
  }
}


// Fused clip + AdamW over a whole score arena, segment aware: the same pass refreshes the masked bf16 operand
// Wm = W16 (.) (S_new > thr[segment]) (the scores are in registers here; a separate pass would re-read 4 B per score)
// and, optionally, clears the gradient after consuming it (model.zero_grad() of the reference loop,
// hg_transformers/mask_trainer_VQA.py:659) so that next step's split-K score-gradient GEMMs reduce-add into zeros
// without a memset per module.  chunks = {start / 8, length, segment, flags}; flags bit 0 = segment has a bf16
// operand to refresh, bit 1 = clear-only chunk (a slice owned by another data-parallel rank).  40 B per score: p, g, m, v, sum read; p, m, v, sum (+ g) written; W16 read, Wm written.
__global__ void __launch_bounds__(kThreads)
adamw_segmented_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       float* __restrict__ sum, const int4* __restrict__ chunks, int nchunks,
                       const float* __restrict__ thr_vec, const uint16_t* __restrict__ w16, uint16_t* __restrict__ wm,
                       AdamArgs a, const float* __restrict__ total_sumsq, const float* __restrict__ hyper,
                       int zero_grad) {
  if (hyper) {
    a.lr = __ldg(hyper);
    a.step_size = __ldg(hyper + 1);
    a.inv_bc2_sqrt = __ldg(hyper + 2);
  }
  float clip = 1.0f;
  if (total_sumsq) {
    const float c = a.max_norm / (sqrtf(__ldg(total_sumsq)) + 1e-6f);
    clip = c < 1.0f ? c : 1.0f;
  }
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int4 ch = __ldg(chunks + c);
    const float thr = a.mode == 0 ? __ldg(thr_vec + ch.z) : 0.f;
    const bool has_wm = (ch.w & 1) != 0;
    const int64_t base = static_cast<int64_t>(ch.x) * 8;
    const int nvec = ch.y >> 3;
    if (ch.w & 2) {      // a slice another rank owns (sharded optimiser): nothing to update, only the gradient to clear
      if (zero_grad) {
        for (int i = threadIdx.x; i < 2 * nvec; i += blockDim.x)
          reinterpret_cast<float4*>(g + base)[i] = make_float4(0, 0, 0, 0);
        for (int i = (nvec << 3) + threadIdx.x; i < ch.y; i += blockDim.x) g[base + i] = 0.f;
      }
      continue;
    }
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      const int64_t e = base + static_cast<int64_t>(i) * 8;
      float4 pp[2], gg[2], mm[2], vv[2], ss[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        pp[h] = reinterpret_cast<const float4*>(p + e)[h];
        gg[h] = reinterpret_cast<const float4*>(g + e)[h];
        mm[h] = reinterpret_cast<const float4*>(m + e)[h];
        vv[h] = reinterpret_cast<const float4*>(v + e)[h];
        ss[h] = sum ? reinterpret_cast<const float4*>(sum + e)[h] : make_float4(0, 0, 0, 0);
      }
      uint4 wv = has_wm ? __ldg(reinterpret_cast<const uint4*>(w16 + e)) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        adam_one(pp[h].x, gg[h].x, mm[h].x, vv[h].x, sum ? &ss[h].x : nullptr, a, clip);
        adam_one(pp[h].y, gg[h].y, mm[h].y, vv[h].y, sum ? &ss[h].y : nullptr, a, clip);
        adam_one(pp[h].z, gg[h].z, mm[h].z, vv[h].z, sum ? &ss[h].z : nullptr, a, clip);
        adam_one(pp[h].w, gg[h].w, mm[h].w, vv[h].w, sum ? &ss[h].w : nullptr, a, clip);
        reinterpret_cast<float4*>(p + e)[h] = pp[h];
        reinterpret_cast<float4*>(m + e)[h] = mm[h];
        reinterpret_cast<float4*>(v + e)[h] = vv[h];
        if (sum) reinterpret_cast<float4*>(sum + e)[h] = ss[h];
        if (zero_grad) reinterpret_cast<float4*>(g + e)[h] = make_float4(0, 0, 0, 0);
      }
      if (has_wm) {
        if (a.mode == 0) {       // stage 2: frozen bf16 weight AND (new score > threshold)
          wv.x &= (pp[0].x > thr ? 0x0000FFFFu : 0u) | (pp[0].y > thr ? 0xFFFF0000u : 0u);
          wv.y &= (pp[0].z > thr ? 0x0000FFFFu : 0u) | (pp[0].w > thr ? 0xFFFF0000u : 0u);
          wv.z &= (pp[1].x > thr ? 0x0000FFFFu : 0u) | (pp[1].y > thr ? 0xFFFF0000u : 0u);
          wv.w &= (pp[1].z > thr ? 0x0000FFFFu : 0u) | (pp[1].w > thr ? 0xFFFF0000u : 0u);
        } else {                 // stage 3: bf16(new weight) where the frozen 0/1 mask (bf16 in `w16`) is set
          wv.x = pack_bf16(pp[0].x, pp[0].y) & ((wv.x & 0xFFFFu ? 0x0000FFFFu : 0u) | (wv.x >> 16 ? 0xFFFF0000u : 0u));
          wv.y = pack_bf16(pp[0].z, pp[0].w) & ((wv.y & 0xFFFFu ? 0x0000FFFFu : 0u) | (wv.y >> 16 ? 0xFFFF0000u : 0u));
          wv.z = pack_bf16(pp[1].x, pp[1].y) & ((wv.z & 0xFFFFu ? 0x0000FFFFu : 0u) | (wv.z >> 16 ? 0xFFFF0000u : 0u));
          wv.w = pack_bf16(pp[1].z, pp[1].w) & ((wv.w & 0xFFFFu ? 0x0000FFFFu : 0u) | (wv.w >> 16 ? 0xFFFF0000u : 0u));
        }
        *reinterpret_cast<uint4*>(wm + e) = wv;
      }
    }
    for (int i = (nvec << 3) + threadIdx.x; i < ch.y; i += blockDim.x) {
      const int64_t e = base + i;
      adam_one(p[e], g[e], m[e], v[e], sum ? sum + e : nullptr, a, clip);
      if (zero_grad) g[e] = 0.f;
      if (has_wm) {
        if (a.mode == 0) {
          wm[e] = p[e] > thr ? w16[e] : uint16_t(0);
        } else {
          __nv_bfloat16 t = __float2bfloat16_rn(p[e]);
          wm[e] = w16[e] ? *reinterpret_cast<uint16_t*>(&t) : uint16_t(0);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ bias gradients: db[n] = sum_m dY[m, n]
// Stage 3 trains the biases too (run_vqa_stage3.py:577-598).  dY is the bf16 [M, N] operand the GEMMs already read;
// one more streaming pass: a CTA owns 256 columns x one slice of the rows (8 warps stride the rows, a lane holds 8
// columns = one 16-byte load), row-slice partials go to a workspace and a second launch adds them in index order
// (deterministic; gridDim.y <= 64 slices).
__global__ void colsum_partial_kernel(const uint16_t* __restrict__ x, int M, int N, float* __restrict__ part) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r_begin = blockIdx.y * rows_per, r_end = min(M, r_begin + rows_per);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < N) {
    for (int r = r_begin + warp; r < r_end; r += 8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * N + c0));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] += __uint_as_float(w[j] << 16);
        acc[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    part[static_cast<size_t>(blockIdx.y) * N + c] = t;
  }
}

// out[j] (+)= sum_i part[i][j], i in index order (also the second stage of the LayerNorm parameter gradients)
__global__ void partial_reduce_kernel(const float* __restrict__ part, int nparts, int n, float* __restrict__ out,
                                      int accumulate) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float t = 0.f;
  for (int i = 0; i < nparts; ++i) t += __ldg(part + static_cast<size_t>(i) * n + j);
  out[j] = accumulate ? out[j] + t : t;
}

// ------------------------------------------------------------------ tiny-K masked linear (box_fc, K = 4)
__global__ void small_k_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                   const float* __restrict__ s, const float* __restrict__ thr_p,
                                   const float* __restrict__ bias, float* __restrict__ y, int M, int N, int K) {
  const float thr = s ? __ldg(thr_p) : 0.f;
  const int64_t total = static_cast<int64_t>(M) * N;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(idx / N), n = static_cast<int>(idx % N);
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      const float wv = (!s || s[static_cast<int64_t>(n) * K + k] > thr) ? w[static_cast<int64_t>(n) * K + k] : 0.f;
      acc = fmaf(x[static_cast<int64_t>(m) * K + k], wv, acc);
    }
    y[idx] = acc + (bias ? bias[n] : 0.f);
  }
}

// dS[n,k] (+)= (sum_m dY[m,n] X[m,k]) * W[n,k]; one block per n, threads stride over m
__global__ void small_k_bwd_ds_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                      const float* __restrict__ w, float* __restrict__ ds, int accumulate, int M,
                                      int N, int K) {
  extern __shared__ float red[];  // [blockDim.x / 32][K]
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k0 = 0; k0 < K; k0 += 8) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int kc = min(8, K - k0);
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
      const float d = dy[static_cast<int64_t>(m) * N + n];
      for (int k = 0; k < kc; ++k) acc[k] = fmaf(d, x[static_cast<int64_t>(m) * K + k0 + k], acc[k]);
    }
    for (int k = 0; k < kc; ++k) {
      float v = acc[k];
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[wid * 8 + k] = v;
    }
    __syncthreads();
    if (threadIdx.x < kc) {
      float t = 0.f;
      for (int i = 0; i < nw; ++i) t += red[i * 8 + threadIdx.x];
      const int64_t o = static_cast<int64_t>(n) * K + k0 + threadIdx.x;
      const float val = t * w[o];
      ds[o] = accumulate ? ds[o] + val : val;
    }
    __syncthreads();
  }
}

__global__ void small_k_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                      const float* __restrict__ s, const float* __restrict__ thr_p,
                                      float* __restrict__ dx, int M, int N, int K) {
  const float thr = s ? __ldg(thr_p) : 0.f;
  const int64_t total = static_cast<int64_t>(M) * K;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(idx / K), k = static_cast<int>(idx % K);
    float acc = 0.f;
    for (int n = 0; n < N; ++n) {
      const int64_t o = static_cast<int64_t>(n) * K + k;
      const float wv = (!s || s[o] > thr) ? w[o] : 0.f;
      acc = fmaf(dy[static_cast<int64_t>(m) * N + n], wv, acc);
    }
    dx[idx] = acc;
  }
}

// ------------------------------------------------------------------ masked embedding
__global__ void embedding_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ w,
                                     const float* __restrict__ s, const float* __restrict__ thr_p,
                                     float* __restrict__ out, int64_t n_tokens, int64_t vocab, int dim) {
  const float thr = __ldg(thr_p);
  const int vec = dim >> 2;
  for (int64_t t = blockIdx.x; t < n_tokens; t += gridDim.x) {
    const long long id = ids[t];
    const bool ok = id >= 0 && id < vocab;
    const float4* wr = reinterpret_cast<const float4*>(w + (ok ? id : 0) * dim);
    const float4* sr = reinterpret_cast<const float4*>(s + (ok ? id : 0) * dim);
    float4* o = reinterpret_cast<float4*>(out + t * dim);
    for (int i = threadIdx.x; i < vec; i += blockDim.x) {
      float4 wv = __ldg(wr + i);
      const float4 sv = __ldg(sr + i);
      wv.x = (ok && sv.x > thr) ? wv.x : 0.f;
      wv.y = (ok && sv.y > thr) ? wv.y : 0.f;
      wv.z = (ok && sv.z > thr) ? wv.z : 0.f;
      wv.w = (ok && sv.w > thr) ? wv.w : 0.f;
      o[i] = wv;
    }
  }
}

__global__ void embedding_bwd_kernel(const long long* __restrict__ ids, const float* __restrict__ dout,
                                     const float* __restrict__ w, float* __restrict__ ds, int64_t n_tokens,
                                     int64_t vocab, int dim, long long padding_idx) {
  for (int64_t t = blockIdx.x; t < n_tokens; t += gridDim.x) {
    const long long id = ids[t];
    if (id < 0 || id >= vocab || id == padding_idx) continue;
    const float* wr = w + id * dim;
    const float* gr = dout + t * dim;
    float* dr = ds + id * dim;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) atomicAdd(dr + i, gr[i] * wr[i]);
  }
}

}  // namespace crv

using namespace crv;

extern "C" int crv_version(void) { return 1; }

extern "C" const char* crv_error_string(int code) {
  switch (code) {
    case CRV_OK: return "ok";
    case CRV_E_BADARG: return "bad argument (null pointer or non-positive size)";
    case CRV_E_ALIGN: return "pointer or leading dimension not 16-byte aligned";
    case CRV_E_SHAPE: return "shape not supported by this entry point";
    case CRV_E_WORKSPACE: return "workspace too small";
    case CRV_E_DRIVER: return "cuTensorMapEncodeTiled driver entry point unavailable";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown error";
  }
}

extern "C" int crv_last_cuda_error(void) { return g_last_cuda_error; }

extern "C" unsigned long long crv_launch_count(void) { return g_launch_count; }

extern "C" int crv_cast_f32_to_bf16(const float* src, uint16_t* dst, int64_t n, void* stream) {
  if (!src || !dst || n < 0) return CRV_E_BADARG;
  if (n == 0) return CRV_OK;
  if (!aligned16(src) || !aligned16(dst)) return CRV_E_ALIGN;
  cast_f32_bf16_kernel<<<stream_grid(n >> 3), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n);
  return launch_status();
}

extern "C" int crv_mul_cast_bf16(const float* w, const float* mask, uint16_t* dst, int64_t n, void* stream) {
  if (!w || !mask || !dst || n < 0) return CRV_E_BADARG;
  if (n == 0) return CRV_OK;
  if (!aligned16(w) || !aligned16(mask) || !aligned16(dst)) return CRV_E_ALIGN;
  mul_cast_bf16_kernel<<<stream_grid(n >> 3), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(w, mask, dst, n);
  return launch_status();
}

extern "C" int crv_binarize(const float* scores, const float* thr, float* mask_f32, uint8_t* mask_u8,
                            long long* kept_count, int64_t n, void* stream) {
  if (!scores || !thr || n < 0) return CRV_E_BADARG;
  if (n == 0) return CRV_OK;
  if (!aligned16(scores) || (mask_f32 && !aligned16(mask_f32)) || (mask_u8 && (reinterpret_cast<uintptr_t>(mask_u8) & 3)))
    return CRV_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kept_count) CRV_CUDA(cudaMemsetAsync(kept_count, 0, sizeof(long long), st));
  binarize_kernel<<<stream_grid(n >> 2), kThreads, 0, st>>>(scores, thr, mask_f32, mask_u8, kept_count, n);
  return launch_status();
}

extern "C" int crv_apply_mask_bf16(const uint16_t* w, const float* scores, const float* thr, uint16_t* wm,
                                   int64_t n, void* stream) {
  if (!w || !scores || !thr || !wm || n < 0) return CRV_E_BADARG;
  if (n == 0) return CRV_OK;
  if (!aligned16(w) || !aligned16(scores) || !aligned16(wm)) return CRV_E_ALIGN;
  apply_mask_kernel<<<stream_grid(n >> 3), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(w, scores, thr, wm, n);
  return launch_status();
}

extern "C" int crv_apply_mask_segmented(const uint16_t* w, const float* scores, const float* thr_vec,
                                        const int* chunks, int nchunks, uint16_t* wm, void* stream) {
  if (!w || !scores || !thr_vec || !chunks || !wm || nchunks < 0) return CRV_E_BADARG;
  if (nchunks == 0) return CRV_OK;
  if (!aligned16(w) || !aligned16(scores) || !aligned16(wm) || !aligned16(chunks)) return CRV_E_ALIGN;
  const int grid = nchunks < num_sms() * 8 ? nchunks : num_sms() * 8;
  apply_mask_segmented_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      w, scores, thr_vec, reinterpret_cast<const int4*>(chunks), nchunks, wm);
  return launch_status();
}

extern "C" int crv_magnitude_init(const float* w, const float* w_thr, float hi, float lo, float* scores, int64_t n,
                                  void* stream) {
  if (!w || !w_thr || !scores || n < 0) return CRV_E_BADARG;
  if (n == 0) return CRV_OK;
  if (!aligned16(w) || !aligned16(scores)) return CRV_E_ALIGN;
  magnitude_init_kernel<<<stream_grid(n >> 2), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(w, w_thr, hi, lo,
                                                                                                  scores, n);
  return launch_status();
}

extern "C" size_t crv_sumsq_workspace_bytes(void) { return (1 + 8 * 148 * 2) * sizeof(float); }

extern "C" int crv_sumsq(const float* x, int64_t n, float* out, void* workspace, void* stream) {
  if (!x || !out || !workspace || n < 0) return CRV_E_BADARG;
  if (n == 0) return CRV_OK;
  if (!aligned16(x)) return CRV_E_ALIGN;
  sumsq_kernel<<<stream_grid(n >> 2, 4), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x, n, out, static_cast<float*>(workspace));
  return launch_status();
}

extern "C" int crv_sumsq_segmented(const float* x, const int* chunks, int nchunks, float* out, void* workspace,
                                   void* stream) {
  if (!x || !chunks || !out || !workspace || nchunks < 0) return CRV_E_BADARG;
  if (nchunks == 0) return CRV_OK;
  if (!aligned16(x) || !aligned16(chunks)) return CRV_E_ALIGN;
  const int grid = nchunks < num_sms() * 4 ? nchunks : num_sms() * 4;
  sumsq_segmented_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<const int4*>(chunks), nchunks, out, static_cast<float*>(workspace));
  return launch_status();
}

extern "C" int crv_adamw_step(float* p, const float* g, float* m, float* v, float* sum, int64_t n, float lr,
                              float step_size, double beta1, double beta2, float eps, float weight_decay,
                              const float* total_sumsq, float max_norm, const float* hyper_dev, int mode,
                              float inv_bc2_sqrt, void* stream) {
  if (!p || !g || !m || !v || n < 0) return CRV_E_BADARG;
  if (n == 0) return CRV_OK;
  if (!aligned16(p) || !aligned16(g) || !aligned16(m) || !aligned16(v) || (sum && !aligned16(sum))) return CRV_E_ALIGN;
  if (mode != 0 && mode != 1) return CRV_E_BADARG;
  const AdamArgs a = make_adam_args(lr, step_size, beta1, beta2, eps, weight_decay, max_norm, inv_bc2_sqrt, mode);
  adamw_kernel<<<stream_grid(n >> 2), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, sum, n, a,
                                                                                         total_sumsq, hyper_dev);
  return launch_status();
}

extern "C" size_t crv_colsum_workspace_bytes(int N) { return N > 0 ? static_cast<size_t>(64) * N * sizeof(float) : 0; }

extern "C" int crv_colsum_bf16(const uint16_t* x, int M, int N, float* out, int accumulate, void* workspace,
                               void* stream) {
  if (!x || !out || !workspace || M <= 0 || N <= 0) return CRV_E_BADARG;
  if (N % 8) return CRV_E_SHAPE;
  if (!aligned16(x)) return CRV_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int col_blocks = (N + 255) / 256;
  int slices = (num_sms() * 4 + col_blocks - 1) / col_blocks;
  if (slices > 64) slices = 64;
  if (slices > (M + 63) / 64) slices = (M + 63) / 64;
  if (slices < 1) slices = 1;
  float* part = static_cast<float*>(workspace);
  colsum_partial_kernel<<<dim3(col_blocks, slices), 256, 0, st>>>(x, M, N, part);
  int rc = launch_status();
  if (rc) return rc;
  partial_reduce_kernel<<<(N + 255) / 256, 256, 0, st>>>(part, slices, N, out, accumulate);
  return launch_status();
}

extern "C" int crv_partial_reduce(const float* part, int nparts, int n, float* out, int accumulate, void* stream) {
  if (!part || !out || nparts <= 0 || n <= 0) return CRV_E_BADARG;
  partial_reduce_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(part, nparts, n, out,
                                                                                         accumulate);
  return launch_status();
}

extern "C" int crv_adamw_segmented(float* p, float* g, float* m, float* v, float* sum, const int* chunks, int nchunks,
                                   const float* thr_vec, const uint16_t* w_bf16, uint16_t* wm_bf16, float lr,
                                   float step_size, double beta1, double beta2, float eps, float weight_decay,
                                   const float* total_sumsq, float max_norm, const float* hyper_dev, int zero_grad,
                                   int mode, float inv_bc2_sqrt, void* stream) {
  if (!p || !g || !m || !v || !chunks || (!thr_vec && mode == 0) || nchunks < 0) return CRV_E_BADARG;
  if ((w_bf16 == nullptr) != (wm_bf16 == nullptr)) return CRV_E_BADARG;
  if (nchunks == 0) return CRV_OK;
  if (!aligned16(p) || !aligned16(g) || !aligned16(m) || !aligned16(v) || (sum && !aligned16(sum)) ||
      !aligned16(chunks) || (w_bf16 && (!aligned16(w_bf16) || !aligned16(wm_bf16))))
    return CRV_E_ALIGN;
  if (mode != 0 && mode != 1) return CRV_E_BADARG;
  const AdamArgs a = make_adam_args(lr, step_size, beta1, beta2, eps, weight_decay, max_norm, inv_bc2_sqrt, mode);
  const int grid = nchunks < num_sms() * 8 ? nchunks : num_sms() * 8;
  adamw_segmented_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, sum, reinterpret_cast<const int4*>(chunks), nchunks, thr_vec, w_bf16, wm_bf16, a, total_sumsq,
      hyper_dev, zero_grad);
  return launch_status();
}

// K <= 8 (box_fc: K = 4) variants with the lanes along N: coalesced 16-byte stores of Y / coalesced reads of dY.
// The general kernels above walk dY with a stride of N floats (one 32-byte sector per 4 useful bytes: 49 us for the
// 28 MB of the box_fc dY) and recompute the masked weight per output element.
namespace crv {
__global__ void small_k8_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                    const float* __restrict__ s, const float* __restrict__ thr_p,
                                    const float* __restrict__ bias, float* __restrict__ y, int M, int N, int K) {
  // thread = 4 consecutive columns n of a row block; its masked weights (4 x K) and biases stay in registers
  const float thr = s ? __ldg(thr_p) : 0.f;
  const int n0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (n0 >= N) return;
  float wm[4][8], b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + j;
    b[j] = (bias && n < N) ? bias[n] : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const bool ok = n < N && k < K;
      const int64_t o = static_cast<int64_t>(n) * K + k;
      wm[j][k] = (ok && (!s || s[o] > thr)) ? w[o] : 0.f;
    }
  }
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m_end = min(M, static_cast<int>((blockIdx.y + 1) * rows_per));
  for (int m = blockIdx.y * rows_per; m < m_end; ++m) {
    float xv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) xv[k] = k < K ? __ldg(x + static_cast<int64_t>(m) * K + k) : 0.f;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(xv[k], wm[j][k], acc);      // k ascending, as the general kernel
      o[j] = acc + b[j];
    }
    float* yr = y + static_cast<int64_t>(m) * N + n0;
    if (n0 + 3 < N && (N & 3) == 0) *reinterpret_cast<float4*>(yr) = make_float4(o[0], o[1], o[2], o[3]);
    else
      for (int j = 0; j < 4 && n0 + j < N; ++j) yr[j] = o[j];
  }
}

// dS[n][k] += w[n][k] * sum_m dy[m][n] x[m][k]: lane = column n (coalesced dY), warps and blockIdx.y split the rows,
// CTA partials go out with one atomicAdd per (n, k) (ds pre-zeroed by the caller unless accumulating)
__global__ void small_k8_bwd_ds_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                       const float* __restrict__ w, float* __restrict__ ds, int M, int N, int K) {
  __shared__ float red[8][8][33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m_end = min(M, static_cast<int>((blockIdx.y + 1) * rows_per));
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int m = blockIdx.y * rows_per + wid; m < m_end; m += 8) {
    const float d = n < N ? __ldg(dy + static_cast<int64_t>(m) * N + n) : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < K) acc[k] = fmaf(d, __ldg(x + static_cast<int64_t>(m) * K + k), acc[k]);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[wid][k][lane] = acc[k];
  __syncthreads();
  const int k = wid;                       // warp k finishes column block x score column k
  if (k < K && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][k][lane];
    const int64_t o = static_cast<int64_t>(n) * K + k;
    atomicAdd(ds + o, t * w[o]);
  }
}
}  // namespace crv

extern "C" int crv_masked_linear_small_k_fwd(const float* x, const float* w, const float* scores, const float* thr,
                                             const float* bias, float* y, int M, int N, int K, void* stream) {
  if (!x || !w || !y || M <= 0 || N <= 0 || K <= 0) return CRV_E_BADARG;
  if (scores && !thr) return CRV_E_BADARG;
  if (K > 64) return CRV_E_SHAPE;
  if (K <= 8 && aligned16(y)) {
    const int col_blocks = (N + 4 * 64 - 1) / (4 * 64);            // 64 threads x 4 columns
    int slices = (num_sms() * 8 + col_blocks - 1) / col_blocks;
    if (slices > M) slices = M;
    small_k8_fwd_kernel<<<dim3(col_blocks, slices), 64, 0, static_cast<cudaStream_t>(stream)>>>(x, w, scores, thr, bias,
                                                                                               y, M, N, K);
    return launch_status();
  }
  const int64_t total = static_cast<int64_t>(M) * N;
  small_k_fwd_kernel<<<stream_grid(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, w, scores, thr, bias,
                                                                                              y, M, N, K);
  return launch_status();
}

extern "C" int crv_masked_linear_small_k_bwd(const float* dy, const float* x, const float* w, const float* scores,
                                             const float* thr, float* dx, float* dscores, int accumulate, int M,
                                             int N, int K, void* stream) {
  if (!dy || !x || !w || !dscores || M <= 0 || N <= 0 || K <= 0) return CRV_E_BADARG;
  if (scores && !thr) return CRV_E_BADARG;
  if (K > 64) return CRV_E_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if (K <= 8) {
    if (!accumulate) CRV_CUDA(cudaMemsetAsync(dscores, 0, static_cast<size_t>(N) * K * sizeof(float), st));
    const int col_blocks = (N + 31) / 32;
    int slices = (num_sms() * 4 + col_blocks - 1) / col_blocks;
    if (slices > (M + 63) / 64) slices = (M + 63) / 64;
    if (slices < 1) slices = 1;
    small_k8_bwd_ds_kernel<<<dim3(col_blocks, slices), 256, 0, st>>>(dy, x, w, dscores, M, N, K);
    rc = launch_status();
  } else {
    small_k_bwd_ds_kernel<<<N, 256, (256 / 32) * 8 * sizeof(float), st>>>(dy, x, w, dscores, accumulate, M, N, K);
    rc = launch_status();
  }
  if (rc) return rc;
  if (dx) {
    const int64_t total = static_cast<int64_t>(M) * K;
    small_k_bwd_dx_kernel<<<stream_grid(total), kThreads, 0, st>>>(dy, w, scores, thr, dx, M, N, K);
    rc = launch_status();
  }
  return rc;
}

extern "C" int crv_masked_embedding_fwd(const long long* ids, const float* w, const float* scores, const float* thr,
                                        float* out, int64_t n_tokens, int64_t vocab, int dim, void* stream) {
  if (!ids || !w || !scores || !thr || !out || n_tokens < 0 || vocab <= 0 || dim <= 0) return CRV_E_BADARG;
  if (n_tokens == 0) return CRV_OK;
  if (dim % 4) return CRV_E_SHAPE;
  if (!aligned16(w) || !aligned16(scores) || !aligned16(out)) return CRV_E_ALIGN;
  const int grid = static_cast<int>(n_tokens < 16 * num_sms() ? n_tokens : 16 * num_sms());
  embedding_fwd_kernel<<<grid, 192, 0, static_cast<cudaStream_t>(stream)>>>(ids, w, scores, thr, out, n_tokens, vocab,
                                                                             dim);
  return launch_status();
}

extern "C" int crv_masked_embedding_bwd(const long long* ids, const float* dout, const float* w, float* dscores,
                                        int64_t n_tokens, int64_t vocab, int dim, long long padding_idx,
                                        void* stream) {
  if (!ids || !dout || !w || !dscores || n_tokens < 0 || vocab <= 0 || dim <= 0) return CRV_E_BADARG;
  if (n_tokens == 0) return CRV_OK;
  const int grid = static_cast<int>(n_tokens < 16 * num_sms() ? n_tokens : 16 * num_sms());
  embedding_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(ids, dout, w, dscores, n_tokens, vocab, dim,
                                                                             padding_idx);
  return launch_status();
}

// Momentum (EMA) update of mPLUG's distillation twins over MANY separately allocated tensors in one launch
// (mPLUG/models/model_vqa_mplug.py:152-156: param_m = param_m * m + param * (1 - m) for every paired parameter).
// PyTorch needs three multi-tensor passes (28 B per element); here 12 B.  Same arithmetic: two rounded products, one
// rounded sum (no FMA contraction), so the result is bit-identical to the reference expression.
// rows = {tensor index, first element / 4, number of float4 (0: scalar tail of < 4 elements follows in .w), tail count}
namespace crv {
__global__ void momentum_update_kernel(const float* const* __restrict__ online, float* const* __restrict__ twins,
                                       const int4* __restrict__ rows, int nrows, float m, float one_minus_m) {
  for (int r = blockIdx.x; r < nrows; r += gridDim.x) {
    const int4 row = __ldg(rows + r);
    const float* p = online[row.x] + static_cast<int64_t>(row.y) * 4;
    float* q = twins[row.x] + static_cast<int64_t>(row.y) * 4;
    for (int i = threadIdx.x; i < row.z; i += blockDim.x) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p) + i);
      float4 b = reinterpret_cast<float4*>(q)[i];
      b.x = __fadd_rn(__fmul_rn(b.x, m), __fmul_rn(a.x, one_minus_m));
      b.y = __fadd_rn(__fmul_rn(b.y, m), __fmul_rn(a.y, one_minus_m));
      b.z = __fadd_rn(__fmul_rn(b.z, m), __fmul_rn(a.z, one_minus_m));
      b.w = __fadd_rn(__fmul_rn(b.w, m), __fmul_rn(a.w, one_minus_m));
      reinterpret_cast<float4*>(q)[i] = b;
    }
    const int64_t t0 = static_cast<int64_t>(row.z) * 4;
    for (int i = threadIdx.x; i < row.w; i += blockDim.x)
      q[t0 + i] = __fadd_rn(__fmul_rn(q[t0 + i], m), __fmul_rn(p[t0 + i], one_minus_m));
  }
}
}  // namespace crv

extern "C" int crv_momentum_update(const float* const* online_dev, float* const* twins_dev, const int* rows_dev,
                                   int nrows, float m, float one_minus_m, void* stream) {
  if (!online_dev || !twins_dev || !rows_dev || nrows < 0) return CRV_E_BADARG;
  if (nrows == 0) return CRV_OK;
  if (!aligned16(rows_dev)) return CRV_E_ALIGN;
  const int grid = nrows < num_sms() * 8 ? nrows : num_sms() * 8;
  momentum_update_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      online_dev, twins_dev, reinterpret_cast<const int4*>(rows_dev), nrows, m, one_minus_m);
  return launch_status();
}

// Global-norm clip + torch.optim.AdamW over MANY separately allocated tensors in two launches (mPLUG keeps its scores
// as the module parameters the reference's checkpoints name; there is no arena).  What the engine does per step with
// PyTorch -- clip_grad_norm_ (norms, scale pass) + the foreach AdamW passes + one apply-mask launch per module on the
// next forward -- is ~88 B per score; here 4 B (norm) + 28 B (p, g, m, v read; p, m, v written) + 4 B (bf16 W read,
// W (.) M written: the new score is in registers).
// rows = {tensor, first element / 8, elements in this row, flags (bit 0: refresh the tensor's masked bf16 operand)}
namespace crv {
__global__ void sumsq_multi_kernel(const float* const* __restrict__ xs, const int4* __restrict__ rows, int nrows,
                                   float* __restrict__ out, float* __restrict__ ws) {
  float acc = 0.f;
  for (int r = blockIdx.x; r < nrows; r += gridDim.x) {
    const int4 row = __ldg(rows + r);
    const float* x = xs[row.x] + static_cast<int64_t>(row.y) * 8;
    const int nvec = row.z >> 2;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int i = (nvec << 2) + threadIdx.x; i < row.z; i += blockDim.x) acc += x[i] * x[i];
  }
  sumsq_finish(acc, ws, out);
}

struct MultiAdamTables {
  float* const* p;
  const float* const* g;
  float* const* m;
  float* const* v;
  const uint16_t* const* w16;
  uint16_t* const* wm;
  const float* const* thr;
};

__global__ void __launch_bounds__(kThreads)
adamw_multi_kernel(MultiAdamTables t, const int4* __restrict__ rows, int nrows, AdamArgs a,
                   const float* __restrict__ total_sumsq) {
  float clip = 1.0f;
  if (total_sumsq) {
    const float c = a.max_norm / (sqrtf(__ldg(total_sumsq)) + 1e-6f);
    clip = c < 1.0f ? c : 1.0f;
  }
  for (int r = blockIdx.x; r < nrows; r += gridDim.x) {
    const int4 row = __ldg(rows + r);
    const int64_t base = static_cast<int64_t>(row.y) * 8;
    float* p = t.p[row.x] + base;
    const float* g = t.g[row.x] + base;
    float* m = t.m[row.x] + base;
    float* v = t.v[row.x] + base;
    const bool has_wm = (row.w & 1) != 0;
    const uint16_t* w16 = has_wm ? t.w16[row.x] + base : nullptr;
    uint16_t* wm = has_wm ? t.wm[row.x] + base : nullptr;
    const float thr = has_wm ? __ldg(t.thr[row.x]) : 0.f;
    const int nvec = row.z >> 3;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      const int e = i * 8;
      float4 pp[2], gg[2], mm[2], vv[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        pp[h] = reinterpret_cast<const float4*>(p + e)[h];
        gg[h] = __ldg(reinterpret_cast<const float4*>(g + e) + h);
        mm[h] = reinterpret_cast<const float4*>(m + e)[h];
        vv[h] = reinterpret_cast<const float4*>(v + e)[h];
      }
      uint4 wv = has_wm ? __ldg(reinterpret_cast<const uint4*>(w16 + e)) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        adam_one(pp[h].x, gg[h].x, mm[h].x, vv[h].x, nullptr, a, clip);
        adam_one(pp[h].y, gg[h].y, mm[h].y, vv[h].y, nullptr, a, clip);
        adam_one(pp[h].z, gg[h].z, mm[h].z, vv[h].z, nullptr, a, clip);
        adam_one(pp[h].w, gg[h].w, mm[h].w, vv[h].w, nullptr, a, clip);
        reinterpret_cast<float4*>(p + e)[h] = pp[h];
        reinterpret_cast<float4*>(m + e)[h] = mm[h];
        reinterpret_cast<float4*>(v + e)[h] = vv[h];
      }
      if (has_wm) {
        wv.x &= (pp[0].x > thr ? 0x0000FFFFu : 0u) | (pp[0].y > thr ? 0xFFFF0000u : 0u);
        wv.y &= (pp[0].z > thr ? 0x0000FFFFu : 0u) | (pp[0].w > thr ? 0xFFFF0000u : 0u);
        wv.z &= (pp[1].x > thr ? 0x0000FFFFu : 0u) | (pp[1].y > thr ? 0xFFFF0000u : 0u);
        wv.w &= (pp[1].z > thr ? 0x0000FFFFu : 0u) | (pp[1].w > thr ? 0xFFFF0000u : 0u);
        *reinterpret_cast<uint4*>(wm + e) = wv;
      }
    }
    for (int e = (nvec << 3) + threadIdx.x; e < row.z; e += blockDim.x) {
      adam_one(p[e], g[e], m[e], v[e], nullptr, a, clip);
      if (has_wm) wm[e] = p[e] > thr ? w16[e] : uint16_t(0);
    }
  }
}
}  // namespace crv

extern "C" int crv_sumsq_multi(const float* const* xs_dev, const int* rows_dev, int nrows, float* out, void* workspace,
                               void* stream) {
  if (!xs_dev || !rows_dev || !out || !workspace || nrows < 0) return CRV_E_BADARG;
  if (nrows == 0) return CRV_OK;
  if (!aligned16(rows_dev)) return CRV_E_ALIGN;
  const int grid = nrows < num_sms() * 8 ? nrows : num_sms() * 8;
  sumsq_multi_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      xs_dev, reinterpret_cast<const int4*>(rows_dev), nrows, out, static_cast<float*>(workspace));
  return launch_status();
}

extern "C" int crv_adamw_multi(float* const* p_dev, const float* const* g_dev, float* const* m_dev, float* const* v_dev,
                               const uint16_t* const* w16_dev, uint16_t* const* wm_dev, const float* const* thr_dev,
                               const int* rows_dev, int nrows, double lr, int step, double beta1, double beta2,
                               double eps, double weight_decay, const float* total_sumsq, float max_norm,
                               void* stream) {
  if (!p_dev || !g_dev || !m_dev || !v_dev || !rows_dev || nrows < 0 || step < 1) return CRV_E_BADARG;
  if ((w16_dev == nullptr) != (wm_dev == nullptr) || (w16_dev == nullptr) != (thr_dev == nullptr)) return CRV_E_BADARG;
  if (nrows == 0) return CRV_OK;
  if (!aligned16(rows_dev)) return CRV_E_ALIGN;
  // torch.optim.AdamW (single-tensor / foreach paths): step_size = lr / (1 - b1^t), denom = sqrt(v) / sqrt(1 - b2^t) + eps
  const double bc1 = 1.0 - std::pow(beta1, step), bc2 = 1.0 - std::pow(beta2, step);
  AdamArgs a = make_adam_args(static_cast<float>(lr), static_cast<float>(lr / bc1), beta1, beta2,
                              static_cast<float>(eps), static_cast<float>(weight_decay), max_norm,
                              static_cast<float>(1.0 / std::sqrt(bc2)), 2);
  a.decay = static_cast<float>(1.0 - lr * weight_decay);
  const MultiAdamTables t{p_dev, g_dev, m_dev, v_dev, w16_dev, wm_dev, thr_dev};
  const int grid = nrows < num_sms() * 8 ? nrows : num_sms() * 8;
  adamw_multi_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      t, reinterpret_cast<const int4*>(rows_dev), nrows, a, total_sumsq);
  return launch_status();
}
