// Few-query attention: softmax(Q K^T / sqrt(d) + mask) V with dropout for a HANDFUL of queries (Lq <= 16) against up
// to 1024 keys per (batch, head), head dim 64, forward and backward.  These are the text-side attention cores of
// mPLUG (mPLUG/models/modeling_mplug.py BertSelfAttention; reference mPLUG/models/modeling_mplug.py:205-300): 16
// question tokens or 6 answer tokens attending to themselves or to 577 / 593 image(+question) tokens.  Library flash
// kernels tile 128 queries: with 6 of 128 rows live and an additive mask they take 426 us forward / 677 us forward +
// backward for the decoder's cross attention (64 x 12 heads, 6 x 593; tests/sdpa_shapes_probe.py) although the work is
// one pass over K and V (116 MB).  Here: one CTA of 128 threads per (batch, head).
//   forward : (1) a thread owns a key: its 128-byte K row against all queries (Q in shared memory as fp32, pre-scaled)
//             (2) a warp owns a query row: max / sum / probabilities; the dropout decision is hashed per element and the
//                 probability is SAVED as bf16 with the decision in its sign (+P kept, -P dropped)
//             (3) a warp owns a subset of the queries, its lanes two of the 64 output columns: O = P~ V, V streamed once
//   backward: from the saved signed probabilities (no exp, no hash): dP~ = dO V^T per key, D_i = sum_j P dP,
//             dS = P (dP - D_i) / sqrt(d); per key dV_j = sum_i P~_ij dO_i and dK_j = sum_i dS_ij Q_i; dQ = dS K.
// Tensors are [B, L, heads * 64] bf16, contiguous (the projections' own outputs: no head split / merge copies).
#include <cmath>

#include "common.cuh"

namespace crv {

namespace {

constexpr int kFqThreads = 128;
constexpr int kFqD = 64;

__device__ __forceinline__ uint64_t fq_mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ float bf_to_f(uint16_t h) { return __uint_as_float(static_cast<uint32_t>(h) << 16); }
__device__ __forceinline__ uint16_t f_to_bf(float x) {
  __nv_bfloat16 t = __float2bfloat16_rn(x);
  return *reinterpret_cast<uint16_t*>(&t);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  f[0] = bf_lo(r.x); f[1] = bf_hi(r.x); f[2] = bf_lo(r.y); f[3] = bf_hi(r.y);
  f[4] = bf_lo(r.z); f[5] = bf_hi(r.z); f[6] = bf_lo(r.w); f[7] = bf_hi(r.w);
}
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_add(float v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct FqParams {
  const uint16_t *q, *k, *v;     // [B, Lq, H*64], [B, Lk, H*64], [B, Lk, H*64]
  const float* mask;             // additive; element (b, i, j) at mask[b * mask_sb + i * mask_sq + j]; null = none
  long long mask_sb, mask_sq;
  uint16_t* out;                 // [B, Lq, H*64]
  uint16_t* probs;               // [B, H, Lq, Lk] signed probabilities
  const uint16_t* dout;          // backward: [B, Lq, H*64]
  uint16_t *dq, *dk, *dv;        // backward: shapes of q, k, v
  int B, H, Lq, Lk, lkp;         // lkp = Lk rounded up to 4 (row pitch of the score arrays in shared memory)
  float scale, p_drop;
  const unsigned long long* rng_state;
  int site;
};

// rows x keys products against a [LQ][64] fp32 matrix in shared memory: acc[i] = sum_d A[i][d] * row[d]
template <int LQ>
__device__ __forceinline__ void dot_rows(const float* __restrict__ A, const uint16_t* __restrict__ row, float (&acc)[LQ]) {
  uint4 r[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) r[c] = __ldg(reinterpret_cast<const uint4*>(row) + c);
#pragma unroll
  for (int i = 0; i < LQ; ++i) acc[i] = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float f[8];
    unpack8(r[c], f);
#pragma unroll
    for (int i = 0; i < LQ; ++i) {
      const float4 a0 = *reinterpret_cast<const float4*>(A + i * kFqD + c * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(A + i * kFqD + c * 8 + 4);
      acc[i] += a0.x * f[0] + a0.y * f[1] + a0.z * f[2] + a0.w * f[3] + a1.x * f[4] + a1.y * f[5] + a1.z * f[6] +
                a1.w * f[7];
    }
  }
}

// dst[i][0..63] = sum_j C[i][j] * M[j][0..63] for all queries i < Lq; M is a [Lk, H*64] bf16 matrix in global memory.
// The keys are dealt to the four warps (j = warp, warp + 4, ...), eight 128-byte rows in flight per warp (the lanes
// cover the 64 columns, two each); the warps' partial sums meet in `red` ([2][LQ][64] floats of shared memory).
// Every thread of the CTA must call this (it synchronises).
template <int LQ>
__device__ __forceinline__ void rows_times_matrix(const float* __restrict__ C, int lkp, const uint16_t* __restrict__ M,
                                                  long long row_stride, int Lk, int warp, int lane,
                                                  float* __restrict__ red, uint16_t* __restrict__ dst,
                                                  long long dst_stride, int Lq) {
  float acc[LQ][2];
#pragma unroll
  for (int i = 0; i < LQ; ++i) acc[i][0] = acc[i][1] = 0.f;
  const uint32_t* col = reinterpret_cast<const uint32_t*>(M) + lane;
  const long long rs = row_stride / 2;      // in 32-bit words
  int j = warp;
  for (; j + 28 < Lk; j += 32) {
    uint32_t w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) w[u] = __ldg(col + (j + 4 * u) * rs);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float lo = bf_lo(w[u]), hi = bf_hi(w[u]);
#pragma unroll
      for (int i = 0; i < LQ; ++i) {
        const float c = C[i * lkp + j + 4 * u];
        acc[i][0] += c * lo;
        acc[i][1] += c * hi;
      }
    }
  }
  for (; j < Lk; j += 4) {
    const uint32_t w = __ldg(col + j * rs);
    const float lo = bf_lo(w), hi = bf_hi(w);
#pragma unroll
    for (int i = 0; i < LQ; ++i) {
      const float c = C[i * lkp + j];
      acc[i][0] += c * lo;
      acc[i][1] += c * hi;
    }
  }
  // 4 -> 2 -> 1 over the warps
  float2* r2 = reinterpret_cast<float2*>(red);
  if (warp >= 2) {
#pragma unroll
    for (int i = 0; i < LQ; ++i) r2[((warp - 2) * LQ + i) * 32 + lane] = make_float2(acc[i][0], acc[i][1]);
  }
  __syncthreads();
  if (warp < 2) {
#pragma unroll
    for (int i = 0; i < LQ; ++i) {
      const float2 t = r2[(warp * LQ + i) * 32 + lane];
      acc[i][0] += t.x;
      acc[i][1] += t.y;
    }
  }
  __syncthreads();
  if (warp == 1) {
#pragma unroll
    for (int i = 0; i < LQ; ++i) r2[i * 32 + lane] = make_float2(acc[i][0], acc[i][1]);
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < LQ; ++i)
      if (i < Lq) {
        const float2 t = r2[i * 32 + lane];
        reinterpret_cast<uint32_t*>(dst + i * dst_stride)[lane] = pack2(acc[i][0] + t.x, acc[i][1] + t.y);
      }
  }
}

// one key's gradient row: dst[0..63] = sum_i c[i] * A[i][0..63]  (A: [LQ][64] fp32 in shared memory)
template <int LQ>
__device__ __forceinline__ void key_row_gradient(const float (&c)[LQ], const float* __restrict__ A,
                                                 uint4* __restrict__ dst) {
#pragma unroll 1
  for (int ch = 0; ch < 8; ++ch) {
    float a[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = 0.f;
#pragma unroll
    for (int i = 0; i < LQ; ++i) {
      const float4 g0 = *reinterpret_cast<const float4*>(A + i * kFqD + ch * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(A + i * kFqD + ch * 8 + 4);
      a[0] += c[i] * g0.x; a[1] += c[i] * g0.y; a[2] += c[i] * g0.z; a[3] += c[i] * g0.w;
      a[4] += c[i] * g1.x; a[5] += c[i] * g1.y; a[6] += c[i] * g1.z; a[7] += c[i] * g1.w;
    }
    dst[ch] = make_uint4(pack2(a[0], a[1]), pack2(a[2], a[3]), pack2(a[4], a[5]), pack2(a[6], a[7]));
  }
}

template <int LQ>
__global__ void __launch_bounds__(kFqThreads) fq_attention_fwd_kernel(FqParams p) {
  extern __shared__ __align__(16) float fq_smem[];
  float* Qs = fq_smem;                 // [2][LQ][64]: first half Q (scaled; rows >= Lq zero), all of it the reduction
                                       // scratch of step (3)
  float* S = fq_smem + 2 * LQ * kFqD;  // [LQ][lkp]; rows >= Lq are zero
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const long long ld = static_cast<long long>(p.H) * kFqD;
  const uint16_t* q = p.q + (static_cast<long long>(b) * p.Lq) * ld + h * kFqD;
  const uint16_t* k = p.k + (static_cast<long long>(b) * p.Lk) * ld + h * kFqD;
  const uint16_t* v = p.v + (static_cast<long long>(b) * p.Lk) * ld + h * kFqD;
  for (int idx = tid; idx < LQ * kFqD; idx += kFqThreads) {
    const int i = idx >> 6, d = idx & 63;
    Qs[idx] = i < p.Lq ? bf_to_f(q[i * ld + d]) * p.scale : 0.f;
  }
  for (int idx = tid; idx < LQ * p.lkp; idx += kFqThreads) S[idx] = 0.f;
  __syncthreads();
  // (1) scores: a thread owns a key
#pragma unroll 1
  for (int j = tid; j < p.Lk; j += kFqThreads) {
    float acc[LQ];
    dot_rows<LQ>(Qs, k + j * ld, acc);
#pragma unroll
    for (int i = 0; i < LQ; ++i)
      if (i < p.Lq) {
        const float m = p.mask ? __ldg(p.mask + b * p.mask_sb + i * p.mask_sq + j) : 0.f;
        S[i * p.lkp + j] = acc[i] + m;
      }
  }
  __syncthreads();
  // (2) softmax + dropout per query row; the probability is saved with the dropout decision in its sign
  uint64_t key = 0;
  uint32_t thresh = 0;
  float keep_scale = 1.f;
  if (p.rng_state != nullptr && p.p_drop > 0.f) {
    const uint64_t seed = p.rng_state[0], ctr = p.rng_state[1];
    key = (seed * 0xD1342543DE82EF95ull) ^ (ctr * 0xA24BAED4963EE407ull) ^ (static_cast<uint64_t>(p.site) << 40);
    thresh = static_cast<uint32_t>(fminf(p.p_drop, 0.9999f) * 65536.0f);
    if (thresh == 0) thresh = 1;
    keep_scale = 1.f / (1.f - p.p_drop);
  }
  for (int i = warp; i < p.Lq; i += 4) {
    float* row = S + i * p.lkp;
    float mx = -INFINITY;
    for (int j = lane; j < p.Lk; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < p.Lk; j += 32) sum += __expf(row[j] - mx);
    sum = warp_add(sum);
    const float inv = 1.f / sum;
    const long long e0 = ((static_cast<long long>(b) * p.H + h) * p.Lq + i) * p.Lk;
    uint16_t* prow = p.probs + e0;
    for (int j = lane; j < p.Lk; j += 32) {
      const uint16_t pb = f_to_bf(__expf(row[j] - mx) * inv);   // what the backward will see
      bool keep = true;
      if (thresh) {
        const uint64_t e = static_cast<uint64_t>(e0 + j);
        const uint64_t hsh = fq_mix64(key + (e >> 2) * 0x9E3779B97F4A7C15ull);
        keep = ((hsh >> (16 * (e & 3))) & 0xFFFFu) >= thresh;
      }
      prow[j] = keep ? pb : static_cast<uint16_t>(pb | 0x8000u);
      row[j] = keep ? bf_to_f(pb) * keep_scale : 0.f;
    }
  }
  __syncthreads();
  // (3) O = P~ V  (Q is no longer needed: its shared memory is the reduction scratch)
  rows_times_matrix<LQ>(S, p.lkp, v, ld, p.Lk, warp, lane, Qs,
                        p.out + (static_cast<long long>(b) * p.Lq) * ld + h * kFqD, ld, p.Lq);
}

template <int LQ>
__global__ void __launch_bounds__(kFqThreads) fq_attention_bwd_kernel(FqParams p) {
  extern __shared__ __align__(16) float fq_smem[];
  float* dOs = fq_smem;                          // [LQ][64]; rows >= Lq are zero
  float* Qs = dOs + LQ * kFqD;                   // [LQ][64] (unscaled); rows >= Lq are zero
  float* dS = Qs + LQ * kFqD;                    // [LQ][lkp] dP, later dS / sqrt(d); rows >= Lq are zero
  uint16_t* Ps = reinterpret_cast<uint16_t*>(dS + LQ * p.lkp);   // [LQ][lkp] the saved signed probabilities (bf16 bits)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const long long ld = static_cast<long long>(p.H) * kFqD;
  const long long qoff = (static_cast<long long>(b) * p.Lq) * ld + h * kFqD;
  const long long koff = (static_cast<long long>(b) * p.Lk) * ld + h * kFqD;
  const uint16_t* q = p.q + qoff;
  const uint16_t* dout = p.dout + qoff;
  const uint16_t* k = p.k + koff;
  const uint16_t* v = p.v + koff;
  const float keep_scale = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
  for (int idx = tid; idx < LQ * kFqD; idx += kFqThreads) {
    const int i = idx >> 6, d = idx & 63;
    dOs[idx] = i < p.Lq ? bf_to_f(dout[i * ld + d]) : 0.f;
    Qs[idx] = i < p.Lq ? bf_to_f(q[i * ld + d]) : 0.f;
  }
  for (int idx = tid; idx < LQ * p.lkp; idx += kFqThreads) {
    dS[idx] = 0.f;
    Ps[idx] = 0;
  }
  __syncthreads();
  // (1) dP~ = dO V^T per key; dP = dP~ / (1 - p) where kept, 0 where dropped
  const uint16_t* probs = p.probs + (static_cast<long long>(b) * p.H + h) * p.Lq * p.Lk;
#pragma unroll 1
  for (int j = tid; j < p.Lk; j += kFqThreads) {
    float acc[LQ];
    dot_rows<LQ>(dOs, v + j * ld, acc);
#pragma unroll
    for (int i = 0; i < LQ; ++i)
      if (i < p.Lq) {
        const uint16_t sp = probs[i * p.Lk + j];
        Ps[i * p.lkp + j] = sp;
        dS[i * p.lkp + j] = (sp & 0x8000u) ? 0.f : acc[i] * keep_scale;
      }
  }
  __syncthreads();
  // (2) D_i = sum_j P dP; dS = P (dP - D_i) / sqrt(d)
  for (int i = warp; i < p.Lq; i += 4) {
    const uint16_t* pr = Ps + i * p.lkp;
    float* dr = dS + i * p.lkp;
    float dsum = 0.f;
    for (int j = lane; j < p.Lk; j += 32) dsum += bf_to_f(pr[j] & 0x7FFFu) * dr[j];
    dsum = warp_add(dsum);
    for (int j = lane; j < p.Lk; j += 32) dr[j] = bf_to_f(pr[j] & 0x7FFFu) * (dr[j] - dsum) * p.scale;
  }
  __syncthreads();
  // (3) per key: dV_j = sum_i P~_ij dO_i, dK_j = sum_i dS_ij Q_i  (a thread owns a key, eight columns at a time)
#pragma unroll 1
  for (int j = tid; j < p.Lk; j += kFqThreads) {
    float c[LQ];
#pragma unroll
    for (int i = 0; i < LQ; ++i) {
      const uint16_t sp = Ps[i * p.lkp + j];
      c[i] = (sp & 0x8000u) ? 0.f : bf_to_f(sp) * keep_scale;
    }
    key_row_gradient<LQ>(c, dOs, reinterpret_cast<uint4*>(p.dv + koff + j * ld));
#pragma unroll
    for (int i = 0; i < LQ; ++i) c[i] = dS[i * p.lkp + j];
    key_row_gradient<LQ>(c, Qs, reinterpret_cast<uint4*>(p.dk + koff + j * ld));
  }
  __syncthreads();      // dO / Q in shared memory are done with: their space is the reduction scratch of step (4)
  // (4) dQ = dS K (dS already carries 1 / sqrt(d))
  rows_times_matrix<LQ>(dS, p.lkp, k, ld, p.Lk, warp, lane, dOs, p.dq + qoff, ld, p.Lq);
}

int fq_check(const FqParams& p) {
  if (!p.q || !p.k || !p.v || !p.probs || p.B <= 0 || p.H <= 0 || p.Lq <= 0 || p.Lk <= 0) return CRV_E_BADARG;
  if (p.Lq > 16 || p.Lk > 1024) return CRV_E_SHAPE;
  if (!(p.p_drop >= 0.f) || p.p_drop >= 1.f) return CRV_E_BADARG;
  if (!aligned16(p.q) || !aligned16(p.k) || !aligned16(p.v)) return CRV_E_ALIGN;
  return CRV_OK;
}

}  // namespace

}  // namespace crv

using crv::FqParams;

extern "C" int crv_fq_attention_fwd(const uint16_t* q, const uint16_t* k, const uint16_t* v, const float* mask,
                                    long long mask_batch_stride, long long mask_query_stride, uint16_t* out,
                                    uint16_t* probs, int B, int heads, int Lq, int Lk, float scale, float p_drop,
                                    const unsigned long long* rng_state, int site, void* stream) {
  FqParams p{};
  p.q = q; p.k = k; p.v = v; p.mask = mask; p.mask_sb = mask_batch_stride; p.mask_sq = mask_query_stride;
  p.out = out; p.probs = probs; p.B = B; p.H = heads; p.Lq = Lq; p.Lk = Lk; p.lkp = (Lk + 3) & ~3;
  p.scale = scale; p.p_drop = p_drop; p.rng_state = rng_state; p.site = site;
  if (!out) return CRV_E_BADARG;
  const int rc = crv::fq_check(p);
  if (rc != CRV_OK) return rc;
  if (!crv::aligned16(out)) return CRV_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(B * heads), block(crv::kFqThreads);
  if (Lq <= 8) {
    const size_t smem = (2 * 8 * crv::kFqD + 8 * static_cast<size_t>(p.lkp)) * sizeof(float);
    crv::fq_attention_fwd_kernel<8><<<grid, block, smem, st>>>(p);
  } else {
    const size_t smem = (2 * 16 * crv::kFqD + 16 * static_cast<size_t>(p.lkp)) * sizeof(float);
    CRV_CUDA(cudaFuncSetAttribute(crv::fq_attention_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>((2 * 16 * crv::kFqD + 16 * 1024) * sizeof(float))));
    crv::fq_attention_fwd_kernel<16><<<grid, block, smem, st>>>(p);
  }
  return crv::launch_status();
}

extern "C" int crv_fq_attention_bwd(const uint16_t* dout, const uint16_t* q, const uint16_t* k, const uint16_t* v,
                                    const uint16_t* probs, uint16_t* dq, uint16_t* dk, uint16_t* dv, int B, int heads,
                                    int Lq, int Lk, float scale, float p_drop, void* stream) {
  FqParams p{};
  p.q = q; p.k = k; p.v = v; p.probs = const_cast<uint16_t*>(probs); p.dout = dout; p.dq = dq; p.dk = dk; p.dv = dv;
  p.B = B; p.H = heads; p.Lq = Lq; p.Lk = Lk; p.lkp = (Lk + 3) & ~3; p.scale = scale; p.p_drop = p_drop;
  if (!dout || !dq || !dk || !dv) return CRV_E_BADARG;
  const int rc = crv::fq_check(p);
  if (rc != CRV_OK) return rc;
  if (!crv::aligned16(dout) || !crv::aligned16(dq) || !crv::aligned16(dk) || !crv::aligned16(dv)) return CRV_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(B * heads), block(crv::kFqThreads);
  if (Lq <= 8) {
    const size_t smem = 2 * 8 * crv::kFqD * sizeof(float) + 8 * static_cast<size_t>(p.lkp) * 6;   // fp32 dS + bf16 P
    CRV_CUDA(cudaFuncSetAttribute(crv::fq_attention_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(2 * 8 * crv::kFqD * sizeof(float) + 8 * 1024 * 6)));
    crv::fq_attention_bwd_kernel<8><<<grid, block, smem, st>>>(p);
  } else {
    const size_t smem = 2 * 16 * crv::kFqD * sizeof(float) + 16 * static_cast<size_t>(p.lkp) * 6;
    CRV_CUDA(cudaFuncSetAttribute(crv::fq_attention_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(2 * 16 * crv::kFqD * sizeof(float) + 16 * 1024 * 6)));
    crv::fq_attention_bwd_kernel<16><<<grid, block, smem, st>>>(p);
  }
  return crv::launch_status();
}
