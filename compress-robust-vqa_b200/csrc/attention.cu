// Small-sequence multi-head attention (20 question tokens / 36 regions / 56 VisualBERT tokens, head dim
// 64): softmax(Q K^T / sqrt(d) + mask) V with dropout, forward and backward ("next" row f3 of SURVEY.md
// section 8; hg_transformers/modeling_lxmert.py:798-827).
//
// Library flash kernels tile 128 x 128 and spend > 90 % of their work on padding at these lengths (cuDNN
// SDPA measured 7.8 ms per training step, 28 % of it).  Here ONE WARP owns one (batch, head) pair: Q, K, V
// (and dO) sit in that warp's shared-memory slab, the four small matmuls run on the tensor cores through
// warp-level 16x16x16 bf16 MMAs with fp32 accumulation, softmax / dropout are fp32 in shared memory, and
// the backward recomputes P instead of storing it.  No block-level barrier is needed (only __syncwarp).
// Inputs are read in place from the fused QKV projection ([B, S, 3H] row stride) and gradients are written
// straight into the fused dQKV tensor, so no split / concat copies exist.
#include <mma.h>

#include <type_traits>

#include "common.cuh"

namespace crv {

using namespace nvcuda;

constexpr int kHeadDim = 64;
constexpr int kLdH = 72;  // bf16 row pitch of Q / K / V / dO tiles (64 + 8: breaks bank alignment, 16-byte multiple)

__device__ __forceinline__ uint64_t amix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

struct AttnParams {
  const __nv_bfloat16 *q, *k, *v;   // element (b, s, h, d) at base + b * bstride + s * sstride + h * 64 + d
  long long q_bs, q_ss, k_bs, k_ss, v_bs, v_ss;
  const float* mask;                // additive, [B, Sk] or null
  __nv_bfloat16* out;               // [B, Sq, heads * 64]
  // backward only
  const __nv_bfloat16* dout;        // [B, Sq, heads * 64]
  __nv_bfloat16 *dq, *dk, *dv;      // same addressing as q / k / v (their own strides)
  long long dq_bs, dq_ss, dk_bs, dk_ss, dv_bs, dv_ss;
  int B, heads, Sq, Sk;
  float scale, p_drop;
  const unsigned long long* rng_state;
  int site;
};

template <int SP>
struct AttnSmem {
  static constexpr int kLdS = SP + 4;   // fp32 pitch of score tiles
  static constexpr int kLdP = SP + 8;   // bf16 pitch of probability tiles
  static constexpr int kTile = SP * kLdH * 2;                 // one bf16 operand tile
  static constexpr int kF32 = SP * ((SP > 64 ? SP : 64) + 4) * 4;  // fp32 scratch: scores or a [SP][64] output
  static constexpr int kP = SP * kLdP * 2;
  static constexpr int kFwd = 3 * kTile + kF32 + kP;
  static constexpr int kBwd = 4 * kTile + 2 * kF32 + kP;
};

// rows [0, S) of a [S][64] bf16 global tile -> smem [SP][72]; rows >= S zero-filled
template <int SP>
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, long long row_stride, int S,
                                          int lane) {
  for (int i = lane; i < SP * 8; i += 32) {
    const int r = i >> 3, c = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < S) v = __ldg(reinterpret_cast<const uint4*>(src + r * row_stride) + c);
    *reinterpret_cast<uint4*>(dst + r * kLdH + c * 8) = v;
  }
}

// fp32 smem [rows][ld] (first 64 columns) -> bf16 global rows
__device__ __forceinline__ void store_tile(__nv_bfloat16* dst, long long row_stride, const float* src, int ld, int S,
                                           int lane) {
  for (int i = lane; i < S * 8; i += 32) {
    const int r = i >> 3, c = i & 7;
    const float* s = src + r * ld + c * 8;
    __nv_bfloat162 a = __floats2bfloat162_rn(s[0], s[1]), b = __floats2bfloat162_rn(s[2], s[3]);
    __nv_bfloat162 e = __floats2bfloat162_rn(s[4], s[5]), f = __floats2bfloat162_rn(s[6], s[7]);
    uint4 v = make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b),
                         *reinterpret_cast<uint32_t*>(&e), *reinterpret_cast<uint32_t*>(&f));
    *reinterpret_cast<uint4*>(dst + r * row_stride + c * 8) = v;
  }
}

// C[MT*16 x NT*16] (fp32, ldc) = A . B with KT k-steps of 16; A / B majors chosen by the caller
template <typename ALayout, typename BLayout>
__device__ __forceinline__ void warp_gemm(float* C, int ldc, const __nv_bfloat16* A, int lda, const __nv_bfloat16* Bm,
                                          int ldb, int MT, int NT, int KT) {
  for (int mt = 0; mt < MT; ++mt)
    for (int nt = 0; nt < NT; ++nt) {
      wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
      wmma::fill_fragment(acc, 0.f);
      for (int kt = 0; kt < KT; ++kt) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, ALayout> a;
        wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, BLayout> b;
        const __nv_bfloat16* ap = std::is_same<ALayout, wmma::row_major>::value ? A + mt * 16 * lda + kt * 16
                                                                                 : A + kt * 16 * lda + mt * 16;
        const __nv_bfloat16* bp = std::is_same<BLayout, wmma::row_major>::value ? Bm + kt * 16 * ldb + nt * 16
                                                                                 : Bm + nt * 16 * ldb + kt * 16;
        wmma::load_matrix_sync(a, ap, lda);
        wmma::load_matrix_sync(b, bp, ldb);
        wmma::mma_sync(acc, a, b, acc);
      }
      wmma::store_matrix_sync(C + mt * 16 * ldc + nt * 16, acc, ldc, wmma::mem_row_major);
    }
}

struct DropKey {
  uint64_t key;
  uint32_t thresh;
  float scale;
  __device__ __forceinline__ bool keep(uint64_t idx) const {
    return thresh == 0 || (amix64(key + idx * 0x9E3779B97F4A7C15ull) & 0xFFFFu) >= thresh;
  }
};
__device__ __forceinline__ DropKey make_key(const unsigned long long* state, int site, float p) {
  DropKey r{0, 0, 1.f};
  if (state != nullptr && p > 0.f) {
    r.key = (state[0] * 0xD1342543DE82EF95ull) ^ (state[1] * 0xA24BAED4963EE407ull) ^ (static_cast<uint64_t>(site) << 40);
    r.thresh = static_cast<uint32_t>(fminf(p, 0.9999f) * 65536.0f);
    if (r.thresh == 0) r.thresh = 1;
    r.scale = 1.f / (1.f - p);
  }
  return r;
}

// softmax over the first Sk columns of rows < Sq of S (fp32, in place -> P); also writes dropout(P) as bf16
template <int SP>
__device__ __forceinline__ void softmax_rows(float* S, __nv_bfloat16* Pd, const AttnParams& p, int b, int pair,
                                             const DropKey& dk, int lane) {
  using L = AttnSmem<SP>;
  for (int r = lane; r < SP; r += 32) {
    float* row = S + r * L::kLdS;
    __nv_bfloat16* prow = Pd + r * L::kLdP;
    if (r >= p.Sq) {
      for (int j = 0; j < SP; ++j) prow[j] = __float2bfloat16(0.f);
      continue;
    }
    float mx = -3.0e38f;
    for (int j = 0; j < p.Sk; ++j) {
      float x = row[j] * p.scale;
      if (p.mask) x += p.mask[static_cast<long long>(b) * p.Sk + j];
      row[j] = x;
      mx = fmaxf(mx, x);
    }
    float sum = 0.f;
    for (int j = 0; j < p.Sk; ++j) {
      const float e = __expf(row[j] - mx);
      row[j] = e;
      sum += e;
    }
    const float inv = 1.f / sum;
    const uint64_t base = (static_cast<uint64_t>(pair) * p.Sq + r) * p.Sk;
    for (int j = 0; j < p.Sk; ++j) {
      const float pr = row[j] * inv;
      row[j] = pr;
      prow[j] = __float2bfloat16(dk.keep(base + j) ? pr * dk.scale : 0.f);
    }
    for (int j = p.Sk; j < SP; ++j) { row[j] = 0.f; prow[j] = __float2bfloat16(0.f); }
  }
}

template <int SP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
attn_fwd_kernel(const AttnParams p) {
  using L = AttnSmem<SP>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * WARPS + warp;
  if (pair >= p.B * p.heads) return;
  const int b = pair / p.heads, h = pair % p.heads;
  uint8_t* base = smem + warp * L::kFwd;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(base + L::kTile);
  __nv_bfloat16* sV = reinterpret_cast<__nv_bfloat16*>(base + 2 * L::kTile);
  float* sS = reinterpret_cast<float*>(base + 3 * L::kTile);
  __nv_bfloat16* sP = reinterpret_cast<__nv_bfloat16*>(base + 3 * L::kTile + L::kF32);
  load_tile<SP>(sQ, p.q + b * p.q_bs + h * kHeadDim, p.q_ss, p.Sq, lane);
  load_tile<SP>(sK, p.k + b * p.k_bs + h * kHeadDim, p.k_ss, p.Sk, lane);
  load_tile<SP>(sV, p.v + b * p.v_bs + h * kHeadDim, p.v_ss, p.Sk, lane);
  __syncwarp();
  const int MT = (p.Sq + 15) / 16, NT = (p.Sk + 15) / 16;
  warp_gemm<wmma::row_major, wmma::col_major>(sS, L::kLdS, sQ, kLdH, sK, kLdH, MT, NT, kHeadDim / 16);
  __syncwarp();
  const DropKey dk = make_key(p.rng_state, p.site, p.p_drop);
  softmax_rows<SP>(sS, sP, p, b, pair, dk, lane);
  __syncwarp();
  float* sO = sS;  // scores are dead; reuse as [SP][68] output staging
  warp_gemm<wmma::row_major, wmma::row_major>(sO, 68, sP, L::kLdP, sV, kLdH, MT, kHeadDim / 16, NT);
  __syncwarp();
  store_tile(p.out + (static_cast<long long>(b) * p.Sq) * (p.heads * kHeadDim) + h * kHeadDim, p.heads * kHeadDim, sO,
             68, p.Sq, lane);
}

template <int SP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
attn_bwd_kernel(const AttnParams p) {
  using L = AttnSmem<SP>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * WARPS + warp;
  if (pair >= p.B * p.heads) return;
  const int b = pair / p.heads, h = pair % p.heads;
  uint8_t* base = smem + warp * L::kBwd;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(base + L::kTile);
  __nv_bfloat16* sV = reinterpret_cast<__nv_bfloat16*>(base + 2 * L::kTile);
  __nv_bfloat16* sdO = reinterpret_cast<__nv_bfloat16*>(base + 3 * L::kTile);
  float* sS = reinterpret_cast<float*>(base + 4 * L::kTile);
  float* sT = reinterpret_cast<float*>(base + 4 * L::kTile + L::kF32);
  __nv_bfloat16* sP = reinterpret_cast<__nv_bfloat16*>(base + 4 * L::kTile + 2 * L::kF32);
  const long long HD = static_cast<long long>(p.heads) * kHeadDim;
  load_tile<SP>(sQ, p.q + b * p.q_bs + h * kHeadDim, p.q_ss, p.Sq, lane);
  load_tile<SP>(sK, p.k + b * p.k_bs + h * kHeadDim, p.k_ss, p.Sk, lane);
  load_tile<SP>(sV, p.v + b * p.v_bs + h * kHeadDim, p.v_ss, p.Sk, lane);
  load_tile<SP>(sdO, p.dout + static_cast<long long>(b) * p.Sq * HD + h * kHeadDim, HD, p.Sq, lane);
  __syncwarp();
  const int MT = (p.Sq + 15) / 16, NT = (p.Sk + 15) / 16, DT = kHeadDim / 16;
  // 1-2. P = softmax(Q K^T * scale + mask) (fp32 in sS), Pd = dropout(P) (bf16 in sP)
  warp_gemm<wmma::row_major, wmma::col_major>(sS, L::kLdS, sQ, kLdH, sK, kLdH, MT, NT, DT);
  __syncwarp();
  const DropKey dk = make_key(p.rng_state, p.site, p.p_drop);
  softmax_rows<SP>(sS, sP, p, b, pair, dk, lane);
  __syncwarp();
  // 3. dV = Pd^T . dO
  warp_gemm<wmma::col_major, wmma::row_major>(sT, 68, sP, L::kLdP, sdO, kLdH, NT, DT, MT);
  __syncwarp();
  store_tile(p.dv + b * p.dv_bs + h * kHeadDim, p.dv_ss, sT, 68, p.Sk, lane);
  __syncwarp();
  // 4. dPd = dO . V^T
  warp_gemm<wmma::row_major, wmma::col_major>(sT, L::kLdS, sdO, kLdH, sV, kLdH, MT, NT, DT);
  __syncwarp();
  // 5. dS = P (.) (dP - sum_j dP_j P_j) * scale, dP = dropout'(dPd)   -> bf16 in sP
  for (int r = lane; r < SP; r += 32) {
    __nv_bfloat16* prow = sP + r * L::kLdP;
    if (r >= p.Sq) {
      for (int j = 0; j < SP; ++j) prow[j] = __float2bfloat16(0.f);
      continue;
    }
    const float* P = sS + r * L::kLdS;
    float* dP = sT + r * L::kLdS;
    const uint64_t ib = (static_cast<uint64_t>(pair) * p.Sq + r) * p.Sk;
    float t = 0.f;
    for (int j = 0; j < p.Sk; ++j) {
      const float d = dk.keep(ib + j) ? dP[j] * dk.scale : 0.f;
      dP[j] = d;
      t += d * P[j];
    }
    for (int j = 0; j < p.Sk; ++j) prow[j] = __float2bfloat16(P[j] * (dP[j] - t) * p.scale);
    for (int j = p.Sk; j < SP; ++j) prow[j] = __float2bfloat16(0.f);
  }
  __syncwarp();
  // 6. dQ = dS . K
  warp_gemm<wmma::row_major, wmma::row_major>(sT, 68, sP, L::kLdP, sK, kLdH, MT, DT, NT);
  __syncwarp();
  store_tile(p.dq + b * p.dq_bs + h * kHeadDim, p.dq_ss, sT, 68, p.Sq, lane);
  __syncwarp();
  // 7. dK = dS^T . Q
  warp_gemm<wmma::col_major, wmma::row_major>(sT, 68, sP, L::kLdP, sQ, kLdH, NT, DT, MT);
  __syncwarp();
  store_tile(p.dk + b * p.dk_bs + h * kHeadDim, p.dk_ss, sT, 68, p.Sk, lane);
}

template <int SP, int WARPS, bool BWD>
static int launch_attn(const AttnParams& p, cudaStream_t st) {
  using L = AttnSmem<SP>;
  constexpr int smem = WARPS * (BWD ? L::kBwd : L::kFwd);
  auto kern = BWD ? attn_bwd_kernel<SP, WARPS> : attn_fwd_kernel<SP, WARPS>;
  static bool configured = false;
  if (!configured) {
    CRV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int pairs = p.B * p.heads;
  kern<<<(pairs + WARPS - 1) / WARPS, WARPS * 32, smem, st>>>(p);
  return launch_status();
}

static int check_attn(const AttnParams& p) {
  if (!p.q || !p.k || !p.v || p.B <= 0 || p.heads <= 0 || p.Sq <= 0 || p.Sk <= 0) return CRV_E_BADARG;
  if (p.Sq > 64 || p.Sk > 64) return CRV_E_SHAPE;
  if ((p.q_ss | p.k_ss | p.v_ss | p.q_bs | p.k_bs | p.v_bs) & 7) return CRV_E_ALIGN;
  if (!aligned16(p.q) || !aligned16(p.k) || !aligned16(p.v)) return CRV_E_ALIGN;
  return CRV_OK;
}

}  // namespace crv

using namespace crv;

extern "C" int crv_attention_fwd(const uint16_t* q, long long q_bs, long long q_ss, const uint16_t* k, long long k_bs,
                                 long long k_ss, const uint16_t* v, long long v_bs, long long v_ss, const float* mask,
                                 uint16_t* out, int B, int heads, int Sq, int Sk, float scale, float p_drop,
                                 const unsigned long long* rng_state, int site, void* stream) {
  AttnParams p{};
  p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.k = reinterpret_cast<const __nv_bfloat16*>(k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(v);
  p.q_bs = q_bs; p.q_ss = q_ss; p.k_bs = k_bs; p.k_ss = k_ss; p.v_bs = v_bs; p.v_ss = v_ss;
  p.mask = mask; p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.B = B; p.heads = heads; p.Sq = Sq; p.Sk = Sk; p.scale = scale; p.p_drop = p_drop; p.rng_state = rng_state; p.site = site;
  int rc = check_attn(p);
  if (rc) return rc;
  if (!out || !aligned16(out)) return CRV_E_BADARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int S = Sq > Sk ? Sq : Sk;
  if (S <= 32) return launch_attn<32, 4, false>(p, st);
  if (S <= 48) return launch_attn<48, 4, false>(p, st);
  return launch_attn<64, 4, false>(p, st);
}

extern "C" int crv_attention_bwd(const uint16_t* q, long long q_bs, long long q_ss, const uint16_t* k, long long k_bs,
                                 long long k_ss, const uint16_t* v, long long v_bs, long long v_ss, const float* mask,
                                 const uint16_t* dout, uint16_t* dq, long long dq_bs, long long dq_ss, uint16_t* dk,
                                 long long dk_bs, long long dk_ss, uint16_t* dv, long long dv_bs, long long dv_ss, int B,
                                 int heads, int Sq, int Sk, float scale, float p_drop,
                                 const unsigned long long* rng_state, int site, void* stream) {
  AttnParams p{};
  p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.k = reinterpret_cast<const __nv_bfloat16*>(k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(v);
  p.q_bs = q_bs; p.q_ss = q_ss; p.k_bs = k_bs; p.k_ss = k_ss; p.v_bs = v_bs; p.v_ss = v_ss;
  p.mask = mask; p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq); p.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.dq_bs = dq_bs; p.dq_ss = dq_ss; p.dk_bs = dk_bs; p.dk_ss = dk_ss; p.dv_bs = dv_bs; p.dv_ss = dv_ss;
  p.B = B; p.heads = heads; p.Sq = Sq; p.Sk = Sk; p.scale = scale; p.p_drop = p_drop; p.rng_state = rng_state; p.site = site;
  int rc = check_attn(p);
  if (rc) return rc;
  if (!dout || !dq || !dk || !dv) return CRV_E_BADARG;
  if ((dq_ss | dk_ss | dv_ss | dq_bs | dk_bs | dv_bs) & 7) return CRV_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int S = Sq > Sk ? Sq : Sk;
  if (S <= 32) return launch_attn<32, 3, true>(p, st);
  if (S <= 48) return launch_attn<48, 3, true>(p, st);
  return launch_attn<64, 2, true>(p, st);
}
