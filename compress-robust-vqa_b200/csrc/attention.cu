// Small-sequence multi-head attention (20 question tokens / 36 regions / 56 VisualBERT tokens, head dim
// 64): softmax(Q K^T / sqrt(d) + mask) V with dropout, forward and backward ("next" row f3 of SURVEY.md
// section 8; hg_transformers/modeling_lxmert.py:798-827).
//
// Library flash kernels tile 128 x 128 and spend > 90 % of their work on padding at these lengths (cuDNN
// SDPA measured 7.8 ms per training step, 28 % of it).  Here one (batch, head) pair is owned by a GROUP of
// KT2 = ceil(S / 16) warps that share the pair's Q / K / V / dO tiles in shared memory; each warp owns one
// 16-row block (query rows in the forward and in pass A of the backward, key rows in pass B).  The small
// matmuls run on the tensor cores (mma.sync m16n8k16, bf16 in / fp32 accumulate), scores and probabilities
// stay in registers, softmax / dropout are fp32, and the backward recomputes P (both in row and in transposed
// orientation) instead of storing it.  One block barrier after the tile load, one between the passes.
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
  // saved-probability path (training): probs[pair][i][j] = +P_ij if the dropout kept (i, j), -P_ij if it dropped it
  // (P = softmax probability BEFORE dropout, bf16; row pitch skp = Sk rounded up to 8, columns Sk .. skp-1 hold 0).
  // The forward writes it, the backward reads it instead of recomputing Q K^T, the softmax and the dropout hash twice.
  __nv_bfloat16* probs;
  int skp;
};

// ---------------------------------------------------------------------------------------------------
// Register-resident formulation (flash-attention-2 style, specialised for S <= 64):
// one warp per 16-row block of a (batch, head) pair; scores / probabilities never leave registers; K, V (and
// Q, dO in the backward) are staged once per pair in shared memory with 16-byte loads and read back as MMA fragments
// with ldmatrix; mma.sync.m16n8k16 (bf16 x bf16 -> fp32).  Lane = 4 g + t:
//   A (16x16): a0 (row g, k 2t..2t+1)  a1 (row g+8, same k)  a2 (row g, k+8)  a3 (row g+8, k+8)
//   B (16x8) : b0 (k 2t..2t+1, n g)    b1 (k 2t+8.., n g)
//   C (16x8) : c0 c1 (row g, n 2t, 2t+1)   c2 c3 (row g+8, n 2t, 2t+1)
// so the C fragments of two adjacent 8-wide score tiles ARE the A fragment of the next GEMM (P . V).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t sptr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
// A fragment of rows [r0, r0+16), k [k0, k0+16) of a row-major [row][k] tile (pitch kLdH)
__device__ __forceinline__ void ld_a(uint32_t (&a)[4], const __nv_bfloat16* tile, int r0, int k0, int lane) {
  const __nv_bfloat16* p = tile + (r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * kLdH + k0 + (lane >> 4) * 8;
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(sptr(p)));
}
// B fragment (k [k0,k0+16), n [n0,n0+8)) from a tile stored [n][k]  (B[k][n] = tile[n][k])
__device__ __forceinline__ void ld_b_nk(uint32_t (&b)[2], const __nv_bfloat16* tile, int n0, int k0, int lane) {
  const __nv_bfloat16* p = tile + (n0 + (lane & 7)) * kLdH + k0 + ((lane >> 3) & 1) * 8;
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(b[0]), "=r"(b[1]) : "r"(sptr(p)));
}
// B fragment from a tile stored [k][n]  (B[k][n] = tile[k][n]) -- transposing load
__device__ __forceinline__ void ld_b_kn(uint32_t (&b)[2], const __nv_bfloat16* tile, int n0, int k0, int lane) {
  const __nv_bfloat16* p = tile + (k0 + (lane & 15)) * kLdH + n0;
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(b[0]), "=r"(b[1]) : "r"(sptr(p)));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

template <int KT2>  // padded length SP = 16 * KT2 covers max(Sq, Sk)
struct AttnSmem {
  static constexpr int SP = 16 * KT2;
  static constexpr int kTile = SP * kLdH * 2;
  static constexpr int kFwd = 2 * kTile;                  // K, V
  static constexpr int kBwd = 4 * kTile + 3 * SP * 4;     // Q, K, V, dO + row max / row sum / D
};

// rows [0, S) of a [S][64] bf16 global tile -> smem [SP][72]; rows >= S zero-filled.  cp.async (16 B, L2 only): all
// of a thread's chunks are in flight at once and nothing is staged through registers -- with plain loads the
// load -> st.shared pairs of the (not unrollable) loop ran one global round trip after the other.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const int bytes = valid ? 16 : 0;      // 0 source bytes: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sptr(smem_dst)), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
template <int SP>
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, long long row_stride, int S,
                                          int tid, int nthreads) {
  for (int i = tid; i < SP * 8; i += nthreads) {
    const int r = i >> 3, c = i & 7;
    const bool valid = r < S;
    cp_async16(dst + r * kLdH + c * 8, src + (valid ? r : 0) * row_stride + c * 8, valid);
  }
}

// Dropout decisions come four to a hash: the 64-bit mix of a 2 x 2 block (query pair, key pair) of the
// probability matrix carries one 16-bit field per element.  An MMA lane holds two elements of such a block in
// both orientations of the problem (same row, adjacent keys in the forward / pass A; adjacent queries, same key
// in pass B), so every hash serves two elements.
struct DropKey {
  uint64_t key;
  uint32_t thresh;
  float scale;
  int half_sk;               // ceil(Sk / 2)
  uint64_t pair_base;        // pair * ceil(Sq / 2)
  __device__ __forceinline__ uint64_t bits(int i, int j) const {      // hash of the block holding (i, j)
    const uint64_t grp = (pair_base + static_cast<uint64_t>(i >> 1)) * half_sk + (j >> 1);
    return amix64(key + grp * 0x9E3779B97F4A7C15ull);
  }
  __device__ __forceinline__ bool keep(uint64_t h, int i, int j) const {
    return ((h >> (16 * ((i & 1) * 2 + (j & 1)))) & 0xFFFFu) >= thresh;
  }
};
__device__ __forceinline__ DropKey make_key(const unsigned long long* state, int site, float p, int pair, int Sq,
                                            int Sk) {
  DropKey r{0, 0, 1.f, (Sk + 1) >> 1, static_cast<uint64_t>(pair) * ((Sq + 1) >> 1)};
  if (state != nullptr && p > 0.f) {
    r.key = (state[0] * 0xD1342543DE82EF95ull) ^ (state[1] * 0xA24BAED4963EE407ull) ^ (static_cast<uint64_t>(site) << 40);
    r.thresh = static_cast<uint32_t>(fminf(p, 0.9999f) * 65536.0f);
    if (r.thresh == 0) r.thresh = 1;
    r.scale = 1.f / (1.f - p);
  }
  return r;
}

// scores of one 16-row block against all keys: acc[nt] = A(16 x 64) . tile[nt*8 .. +8][64]^T
template <int KT2>
__device__ __forceinline__ void scores_16(float (&acc)[2 * KT2][4], const uint32_t (&a)[4][4], const __nv_bfloat16* sB,
                                          int lane) {
#pragma unroll
  for (int nt = 0; nt < 2 * KT2; ++nt) {
    acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      uint32_t b[2];
      ld_b_nk(b, sB, nt * 8, kt * 16, lane);
      mma16816(acc[nt], a[kt], b);
    }
  }
}

template <int KT2, int PAIRS>
__global__ void __launch_bounds__(PAIRS * KT2 * 32)
attn_fwd_kernel(const AttnParams p) {
  pdl_wait();
  pdl_launch_dependents();
  using L = AttnSmem<KT2>;
  constexpr int SP = L::SP, NT = 2 * KT2;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp / KT2, tile = warp % KT2;      // pair slot inside the CTA, 16-row block inside the pair
  const int pair = blockIdx.x * PAIRS + slot;
  const bool valid = pair < p.B * p.heads;
  const int b = valid ? pair / p.heads : 0, h = valid ? pair % p.heads : 0;
  const int g = lane >> 2, t = lane & 3;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem + slot * L::kFwd);
  __nv_bfloat16* sV = sK + SP * kLdH;
  if (valid) {
    load_tile<SP>(sK, p.k + b * p.k_bs + h * kHeadDim, p.k_ss, p.Sk, tile * 32 + lane, KT2 * 32);
    load_tile<SP>(sV, p.v + b * p.v_bs + h * kHeadDim, p.v_ss, p.Sk, tile * 32 + lane, KT2 * 32);
  }
  // this warp's Q fragments straight from global memory, issued BEFORE waiting for the K / V tiles: one global round
  // trip instead of two in a row
  const __nv_bfloat16* qb = p.q + b * p.q_bs + h * kHeadDim;
  const int mt = tile;
  const bool active = valid && mt * 16 < p.Sq;
  const int r0 = mt * 16 + g, r1 = r0 + 8;
  uint32_t a[4][4];
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    const int c = kt * 16 + 2 * t;
    a[kt][0] = (active && r0 < p.Sq) ? __ldg(reinterpret_cast<const uint32_t*>(qb + r0 * p.q_ss + c)) : 0u;
    a[kt][1] = (active && r1 < p.Sq) ? __ldg(reinterpret_cast<const uint32_t*>(qb + r1 * p.q_ss + c)) : 0u;
    a[kt][2] = (active && r0 < p.Sq) ? __ldg(reinterpret_cast<const uint32_t*>(qb + r0 * p.q_ss + c + 8)) : 0u;
    a[kt][3] = (active && r1 < p.Sq) ? __ldg(reinterpret_cast<const uint32_t*>(qb + r1 * p.q_ss + c + 8)) : 0u;
  }
  cp_async_wait_all();
  __syncthreads();
  const DropKey dk = make_key(p.rng_state, p.site, p.p_drop, pair, p.Sq, p.Sk);
  const float* mrow = p.mask ? p.mask + static_cast<long long>(b) * p.Sk : nullptr;
  const long long HD = static_cast<long long>(p.heads) * kHeadDim;
  if (active) {
    float s[NT][4];
    scores_16<KT2>(s, a, sK, lane);
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = nt * 8 + 2 * t + e;
        const float add = j < p.Sk ? (mrow ? mrow[j] : 0.f) : -INFINITY;
        s[nt][e] = s[nt][e] * p.scale + add;
        s[nt][2 + e] = s[nt][2 + e] * p.scale + add;
        m0 = fmaxf(m0, s[nt][e]);
        m1 = fmaxf(m1, s[nt][2 + e]);
      }
    }
    m0 = quad_max(m0);
    m1 = quad_max(m1);
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] = __expf(s[nt][e] - m0);
        s[nt][2 + e] = __expf(s[nt][2 + e] - m1);
        l0 += s[nt][e];
        l1 += s[nt][2 + e];
      }
    }
    const float i0 = 1.f / quad_sum(l0), i1 = 1.f / quad_sum(l1);
    uint32_t pa[KT2][4];
    __nv_bfloat16* pr = p.probs ? p.probs + static_cast<size_t>(pair) * p.Sq * p.skp : nullptr;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float v[4];
      const int jb = nt * 8 + 2 * t;
      const uint64_t h0 = dk.thresh ? dk.bits(r0, jb) : 0ull, h1 = dk.thresh ? dk.bits(r1, jb) : 0ull;
      bool kp[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        kp[e] = dk.keep(h0, r0, jb + e);
        kp[2 + e] = dk.keep(h1, r1, jb + e);
        s[nt][e] *= i0;                      // P
        s[nt][2 + e] *= i1;
        v[e] = kp[e] ? s[nt][e] * dk.scale : 0.f;
        v[2 + e] = kp[2 + e] ? s[nt][2 + e] * dk.scale : 0.f;
      }
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(v[0], v[1]);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(v[2], v[3]);
      if (pr != nullptr && jb < p.skp) {     // the sign carries the dropout decision (P >= 0)
        if (r0 < p.Sq)
          *reinterpret_cast<uint32_t*>(pr + r0 * p.skp + jb) =
              pack_bf16x2(kp[0] ? s[nt][0] : -s[nt][0], kp[1] ? s[nt][1] : -s[nt][1]);
        if (r1 < p.Sq)
          *reinterpret_cast<uint32_t*>(pr + r1 * p.skp + jb) =
              pack_bf16x2(kp[2] ? s[nt][2] : -s[nt][2], kp[3] ? s[nt][3] : -s[nt][3]);
      }
    }
    __nv_bfloat16* ob = p.out + (static_cast<long long>(b) * p.Sq) * HD + h * kHeadDim;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kt = 0; kt < KT2; ++kt) {
        uint32_t bv[2];
        ld_b_kn(bv, sV, dt * 8, kt * 16, lane);
        mma16816(o, pa[kt], bv);
      }
      const int c = dt * 8 + 2 * t;
      if (r0 < p.Sq) *reinterpret_cast<uint32_t*>(ob + r0 * HD + c) = pack_bf16x2(o[0], o[1]);
      if (r1 < p.Sq) *reinterpret_cast<uint32_t*>(ob + r1 * HD + c) = pack_bf16x2(o[2], o[3]);
    }
  }
}

template <int KT2, int PAIRS>
__global__ void __launch_bounds__(PAIRS * KT2 * 32)
attn_bwd_kernel(const AttnParams p) {
  pdl_wait();
  pdl_launch_dependents();
  using L = AttnSmem<KT2>;
  constexpr int SP = L::SP, NT = 2 * KT2;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp / KT2, tile = warp % KT2;
  const int pair = blockIdx.x * PAIRS + slot;
  const bool valid = pair < p.B * p.heads;
  const int b = valid ? pair / p.heads : 0, h = valid ? pair % p.heads : 0;
  const int g = lane >> 2, t = lane & 3;
  uint8_t* base = smem + slot * L::kBwd;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* sK = sQ + SP * kLdH;
  __nv_bfloat16* sV = sK + SP * kLdH;
  __nv_bfloat16* sdO = sV + SP * kLdH;
  float* sM = reinterpret_cast<float*>(base + 4 * L::kTile);
  float* sL = sM + SP;
  float* sD = sL + SP;
  const long long HD = static_cast<long long>(p.heads) * kHeadDim;
  if (valid) {
    const int gt = tile * 32 + lane;
    load_tile<SP>(sQ, p.q + b * p.q_bs + h * kHeadDim, p.q_ss, p.Sq, gt, KT2 * 32);
    load_tile<SP>(sK, p.k + b * p.k_bs + h * kHeadDim, p.k_ss, p.Sk, gt, KT2 * 32);
    load_tile<SP>(sV, p.v + b * p.v_bs + h * kHeadDim, p.v_ss, p.Sk, gt, KT2 * 32);
    load_tile<SP>(sdO, p.dout + static_cast<long long>(b) * p.Sq * HD + h * kHeadDim, HD, p.Sq, gt, KT2 * 32);
  }
  cp_async_wait_all();
  __syncthreads();
  const DropKey dk = make_key(p.rng_state, p.site, p.p_drop, pair, p.Sq, p.Sk);
  const float* mrow = p.mask ? p.mask + static_cast<long long>(b) * p.Sk : nullptr;

  // ---- pass A: this warp's query-row block.  P, dP, D_i = sum_j dP_ij P_ij, dS -> dQ; row statistics to smem
  const int mt = tile;
  if (valid && mt * 16 < p.Sq) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    uint32_t a[4][4];
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) ld_a(a[kt], sQ, mt * 16, kt * 16, lane);
    float s[NT][4];
    scores_16<KT2>(s, a, sK, lane);
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = nt * 8 + 2 * t + e;
        const float add = j < p.Sk ? (mrow ? mrow[j] : 0.f) : -INFINITY;
        s[nt][e] = s[nt][e] * p.scale + add;
        s[nt][2 + e] = s[nt][2 + e] * p.scale + add;
        m0 = fmaxf(m0, s[nt][e]);
        m1 = fmaxf(m1, s[nt][2 + e]);
      }
    }
    m0 = quad_max(m0);
    m1 = quad_max(m1);
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] = __expf(s[nt][e] - m0);
        s[nt][2 + e] = __expf(s[nt][2 + e] - m1);
        l0 += s[nt][e];
        l1 += s[nt][2 + e];
      }
    }
    l0 = quad_sum(l0);
    l1 = quad_sum(l1);
    if (t == 0) { sM[r0] = m0; sL[r0] = l0; sM[r1] = m1; sL[r1] = l1; }
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    // dPd = dO . V^T
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) ld_a(a[kt], sdO, mt * 16, kt * 16, lane);
    float dp[NT][4];
    scores_16<KT2>(dp, a, sV, lane);
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int jb = nt * 8 + 2 * t;
      const uint64_t h0 = dk.thresh ? dk.bits(r0, jb) : 0ull, h1 = dk.thresh ? dk.bits(r1, jb) : 0ull;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = jb + e;
        s[nt][e] *= i0;            // P
        s[nt][2 + e] *= i1;
        dp[nt][e] = (j < p.Sk && dk.keep(h0, r0, j)) ? dp[nt][e] * dk.scale : 0.f;
        dp[nt][2 + e] = (j < p.Sk && dk.keep(h1, r1, j)) ? dp[nt][2 + e] * dk.scale : 0.f;
        d0 += dp[nt][e] * s[nt][e];
        d1 += dp[nt][2 + e] * s[nt][2 + e];
      }
    }
    d0 = quad_sum(d0);
    d1 = quad_sum(d1);
    if (t == 0) { sD[r0] = d0; sD[r1] = d1; }
    uint32_t da[KT2][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      da[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(s[nt][0] * (dp[nt][0] - d0) * p.scale, s[nt][1] * (dp[nt][1] - d0) * p.scale);
      da[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(s[nt][2] * (dp[nt][2] - d1) * p.scale, s[nt][3] * (dp[nt][3] - d1) * p.scale);
    }
    __nv_bfloat16* qo = p.dq + b * p.dq_bs + h * kHeadDim;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kt = 0; kt < KT2; ++kt) {
        uint32_t bk[2];
        ld_b_kn(bk, sK, dt * 8, kt * 16, lane);   // B[k = j][n = d] = K[j][d]
        mma16816(o, da[kt], bk);
      }
      const int c = dt * 8 + 2 * t;
      if (r0 < p.Sq) *reinterpret_cast<uint32_t*>(qo + r0 * p.dq_ss + c) = pack_bf16x2(o[0], o[1]);
      if (r1 < p.Sq) *reinterpret_cast<uint32_t*>(qo + r1 * p.dq_ss + c) = pack_bf16x2(o[2], o[3]);
    }
  }
  __syncthreads();   // every row block's statistics are in shared memory

  // ---- pass B: this warp's key-row block on the transposed problem.  P^T, dP^T -> dV = Pd^T dO, dK = dS^T Q
  const int jt = tile;
  if (valid && jt * 16 < p.Sk) {
    const int j0 = jt * 16 + g, j1 = j0 + 8;
    uint32_t a[4][4];
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) ld_a(a[kt], sK, jt * 16, kt * 16, lane);
    float s[NT][4];
    scores_16<KT2>(s, a, sQ, lane);               // S^T[j][i] = K[j] . Q[i]
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) ld_a(a[kt], sV, jt * 16, kt * 16, lane);
    float dp[NT][4];
    scores_16<KT2>(dp, a, sdO, lane);             // dPd^T[j][i] = V[j] . dO[i]
    const float add0 = j0 < p.Sk ? (mrow ? mrow[j0] : 0.f) : 0.f, add1 = j1 < p.Sk ? (mrow ? mrow[j1] : 0.f) : 0.f;
    uint32_t pa[KT2][4], da[KT2][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float pv[4], dv[4];
      const int ib = nt * 8 + 2 * t;
      const uint64_t h0 = dk.thresh ? dk.bits(ib, j0) : 0ull, h1 = dk.thresh ? dk.bits(ib, j1) : 0ull;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = ib + e;
        const bool iv = i < p.Sq;
        const float mi = iv ? sM[i] : 0.f, li = iv ? 1.f / sL[i] : 0.f, Di = iv ? sD[i] : 0.f;
        const float p0 = (iv && j0 < p.Sk) ? __expf(s[nt][e] * p.scale + add0 - mi) * li : 0.f;
        const float p1 = (iv && j1 < p.Sk) ? __expf(s[nt][2 + e] * p.scale + add1 - mi) * li : 0.f;
        const bool k0 = iv && j0 < p.Sk && dk.keep(h0, i, j0), k1 = iv && j1 < p.Sk && dk.keep(h1, i, j1);
        pv[e] = k0 ? p0 * dk.scale : 0.f;
        pv[2 + e] = k1 ? p1 * dk.scale : 0.f;
        const float dp0 = k0 ? dp[nt][e] * dk.scale : 0.f, dp1 = k1 ? dp[nt][2 + e] * dk.scale : 0.f;
        dv[e] = p0 * (dp0 - Di) * p.scale;
        dv[2 + e] = p1 * (dp1 - Di) * p.scale;
      }
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(pv[0], pv[1]);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
      da[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(dv[0], dv[1]);
      da[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(dv[2], dv[3]);
    }
    __nv_bfloat16* vo = p.dv + b * p.dv_bs + h * kHeadDim;
    __nv_bfloat16* ko = p.dk + b * p.dk_bs + h * kHeadDim;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      float ov[4] = {0.f, 0.f, 0.f, 0.f}, ok[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kt = 0; kt < KT2; ++kt) {
        uint32_t bo[2], bq[2];
        ld_b_kn(bo, sdO, dt * 8, kt * 16, lane);   // B[k = i][n = d] = dO[i][d]
        ld_b_kn(bq, sQ, dt * 8, kt * 16, lane);    // B[k = i][n = d] = Q[i][d]
        mma16816(ov, pa[kt], bo);
        mma16816(ok, da[kt], bq);
      }
      const int c = dt * 8 + 2 * t;
      if (j0 < p.Sk) {
        *reinterpret_cast<uint32_t*>(vo + j0 * p.dv_ss + c) = pack_bf16x2(ov[0], ov[1]);
        *reinterpret_cast<uint32_t*>(ko + j0 * p.dk_ss + c) = pack_bf16x2(ok[0], ok[1]);
      }
      if (j1 < p.Sk) {
        *reinterpret_cast<uint32_t*>(vo + j1 * p.dv_ss + c) = pack_bf16x2(ov[2], ov[3]);
        *reinterpret_cast<uint32_t*>(ko + j1 * p.dk_ss + c) = pack_bf16x2(ok[2], ok[3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Backward from the SAVED probabilities (the training path).  The recomputing kernel above spends ~3800 warp
// instructions per 16-row block on Q K^T in both orientations, two softmaxes and two rounds of dropout hashing and is
// issue-bound at 23 % of HBM bandwidth; with P (and the dropout decision in its sign) read back, what is left is
//   pass A (query rows):  dPd = dO V^T,  D_i = sum_j dPd_ij Pd_ij,  dS = P (.) (keep dPd / (1-p) - D) scale  -> dQ = dS K,
//                         dS also goes to shared memory (bf16)
//   pass B (key rows):    P^T and dS^T arrive as MMA A fragments through transposing ldmatrix from the two [i][j]
//                         tiles:  dV = Pd^T dO,  dK = dS^T Q
// -- five small GEMMs instead of eight, no exp, no hash.
template <int KT2>
struct AttnSmemP {
  static constexpr int SP = 16 * KT2;
  static constexpr int SPP = SP + 8;                 // bf16 pitch of the P / dS tile: (SP + 8) / 8 is odd -> ldmatrix rows spread over the banks
  static constexpr int kTile = SP * kLdH * 2;
  static constexpr int kP = SP * SPP * 2;
  static constexpr int kTotal = 4 * kTile + kP;      // Q, K, V, dO, and ONE [i][j] tile that holds P, then dS in place
};

// A fragment (rows r0 .. r0+15, k k0 .. k0+15) of the TRANSPOSE of a tile stored [k][row] with pitch `ld`
__device__ __forceinline__ void ld_a_trans(uint32_t (&a)[4], const __nv_bfloat16* tile, int ld, int r0, int k0, int lane) {
  const __nv_bfloat16* p = tile + (k0 + (lane & 7) + (lane >> 4) * 8) * ld + r0 + ((lane >> 3) & 1) * 8;
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(sptr(p)));
}
__device__ __forceinline__ float bf16_lo_f(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi_f(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

// Order of work (three block barriers): tiles land -> dV = Pd^T dO for this warp's key rows (reads ALL rows of P)
// -> barrier -> pass A for this warp's query rows: dS overwrites P in place (a lane rewrites exactly the elements it
// read), dQ = dS K -> barrier -> dK = dS^T Q.  One shared [i][j] tile instead of two: 33 KB per (batch, head) at
// S = 36, six CTAs per SM instead of five.
template <int KT2>
__global__ void __launch_bounds__(KT2 * 32)
attn_bwd_p_kernel(const AttnParams p) {
  pdl_wait();
  pdl_launch_dependents();
  using L = AttnSmemP<KT2>;
  constexpr int SP = L::SP, SPP = L::SPP, NT = 2 * KT2;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tile = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x;
  const int b = pair / p.heads, h = pair % p.heads;
  const int g = lane >> 2, t = lane & 3;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sK = sQ + SP * kLdH;
  __nv_bfloat16* sV = sK + SP * kLdH;
  __nv_bfloat16* sdO = sV + SP * kLdH;
  __nv_bfloat16* sP = sdO + SP * kLdH;       // P (signed), later dS
  const long long HD = static_cast<long long>(p.heads) * kHeadDim;
  {
    const int tid = threadIdx.x;
    load_tile<SP>(sQ, p.q + b * p.q_bs + h * kHeadDim, p.q_ss, p.Sq, tid, KT2 * 32);
    load_tile<SP>(sK, p.k + b * p.k_bs + h * kHeadDim, p.k_ss, p.Sk, tid, KT2 * 32);
    load_tile<SP>(sV, p.v + b * p.v_bs + h * kHeadDim, p.v_ss, p.Sk, tid, KT2 * 32);
    load_tile<SP>(sdO, p.dout + static_cast<long long>(b) * p.Sq * HD + h * kHeadDim, HD, p.Sq, tid, KT2 * 32);
    const __nv_bfloat16* pr = p.probs + static_cast<size_t>(pair) * p.Sq * p.skp;
    for (int i = tid; i < SP * (SP / 8); i += KT2 * 32) {     // whole [SP][SP] tile: rows >= Sq, columns >= skp zero-filled
      const int r = i / (SP / 8), c = i % (SP / 8);
      const bool valid = r < p.Sq && c * 8 < p.skp;
      cp_async16(sP + r * SPP + c * 8, pr + (valid ? r * p.skp + c * 8 : 0), valid);
    }
  }
  const float dscale = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
  cp_async_wait_all();
  __syncthreads();

  // ---- dV = Pd^T dO for this warp's key rows (P^T arrives as MMA A fragments through a transposing ldmatrix)
  const int jt = tile;
  const bool keys = jt * 16 < p.Sk;
  const int j0 = jt * 16 + g, j1 = j0 + 8;
  if (keys) {
    uint32_t pa[KT2][4];
    const __nv_bfloat162 ds2 = __floats2bfloat162_rn(dscale, dscale);
#pragma unroll
    for (int kt = 0; kt < KT2; ++kt) {
      ld_a_trans(pa[kt], sP, SPP, jt * 16, kt * 16, lane);
#pragma unroll
      for (int e = 0; e < 4; ++e) {      // signed P -> Pd: dropped entries (sign set) to +0, kept ones times 1 / (1 - p)
        uint32_t x = pa[kt][e];
        x &= ((x & 0x8000u) ? 0u : 0x0000FFFFu) | ((x & 0x80000000u) ? 0u : 0xFFFF0000u);
        if (p.p_drop > 0.f) {
          __nv_bfloat162 v = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&x), ds2);
          x = *reinterpret_cast<uint32_t*>(&v);
        }
        pa[kt][e] = x;
      }
    }
    __nv_bfloat16* vo = p.dv + b * p.dv_bs + h * kHeadDim;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      float ov[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kt = 0; kt < KT2; ++kt) {
        uint32_t bo[2];
        ld_b_kn(bo, sdO, dt * 8, kt * 16, lane);   // B[k = i][n = d] = dO[i][d]
        mma16816(ov, pa[kt], bo);
      }
      const int c = dt * 8 + 2 * t;
      if (j0 < p.Sk) *reinterpret_cast<uint32_t*>(vo + j0 * p.dv_ss + c) = pack_bf16x2(ov[0], ov[1]);
      if (j1 < p.Sk) *reinterpret_cast<uint32_t*>(vo + j1 * p.dv_ss + c) = pack_bf16x2(ov[2], ov[3]);
    }
  }
  __syncthreads();   // every warp has read P^T: the tile may now turn into dS

  // ---- pass A: this warp's query-row block: dPd = dO V^T, D, dS (in place over P), dQ = dS K
  const int mt = tile;
  if (mt * 16 < p.Sq) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    uint32_t a[4][4];
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) ld_a(a[kt], sdO, mt * 16, kt * 16, lane);
    float dp[NT][4];
    scores_16<KT2>(dp, a, sV, lane);               // dPd = dO . V^T
    float pv[NT][4];
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int jb = nt * 8 + 2 * t;
      const uint32_t u0 = *reinterpret_cast<const uint32_t*>(sP + r0 * SPP + jb);
      const uint32_t u1 = *reinterpret_cast<const uint32_t*>(sP + r1 * SPP + jb);
      const float s0 = bf16_lo_f(u0), s1 = bf16_hi_f(u0), s2 = bf16_lo_f(u1), s3 = bf16_hi_f(u1);
      pv[nt][0] = fabsf(s0); pv[nt][1] = fabsf(s1); pv[nt][2] = fabsf(s2); pv[nt][3] = fabsf(s3);
      dp[nt][0] = (u0 & 0x8000u) ? 0.f : dp[nt][0] * dscale;          // dP = keep dPd / (1 - p)
      dp[nt][1] = (u0 & 0x80000000u) ? 0.f : dp[nt][1] * dscale;
      dp[nt][2] = (u1 & 0x8000u) ? 0.f : dp[nt][2] * dscale;
      dp[nt][3] = (u1 & 0x80000000u) ? 0.f : dp[nt][3] * dscale;
      d0 += dp[nt][0] * pv[nt][0] + dp[nt][1] * pv[nt][1];
      d1 += dp[nt][2] * pv[nt][2] + dp[nt][3] * pv[nt][3];
    }
    d0 = quad_sum(d0);
    d1 = quad_sum(d1);
    uint32_t da[KT2][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int jb = nt * 8 + 2 * t;
      const uint32_t w0 = pack_bf16x2(pv[nt][0] * (dp[nt][0] - d0) * p.scale, pv[nt][1] * (dp[nt][1] - d0) * p.scale);
      const uint32_t w1 = pack_bf16x2(pv[nt][2] * (dp[nt][2] - d1) * p.scale, pv[nt][3] * (dp[nt][3] - d1) * p.scale);
      da[nt >> 1][(nt & 1) * 2 + 0] = w0;
      da[nt >> 1][(nt & 1) * 2 + 1] = w1;
      *reinterpret_cast<uint32_t*>(sP + r0 * SPP + jb) = w0;        // dS over P: same lane, same elements
      *reinterpret_cast<uint32_t*>(sP + r1 * SPP + jb) = w1;
    }
    __nv_bfloat16* qo = p.dq + b * p.dq_bs + h * kHeadDim;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kt = 0; kt < KT2; ++kt) {
        uint32_t bk[2];
        ld_b_kn(bk, sK, dt * 8, kt * 16, lane);   // B[k = j][n = d] = K[j][d]
        mma16816(o, da[kt], bk);
      }
      const int c = dt * 8 + 2 * t;
      if (r0 < p.Sq) *reinterpret_cast<uint32_t*>(qo + r0 * p.dq_ss + c) = pack_bf16x2(o[0], o[1]);
      if (r1 < p.Sq) *reinterpret_cast<uint32_t*>(qo + r1 * p.dq_ss + c) = pack_bf16x2(o[2], o[3]);
    }
  }
  // row blocks without query rows keep their zero-filled P rows, which read as dS = 0 below
  __syncthreads();   // dS of every row block is in shared memory

  // ---- dK = dS^T Q for this warp's key rows
  if (keys) {
    uint32_t da[KT2][4];
#pragma unroll
    for (int kt = 0; kt < KT2; ++kt) ld_a_trans(da[kt], sP, SPP, jt * 16, kt * 16, lane);
    __nv_bfloat16* ko = p.dk + b * p.dk_bs + h * kHeadDim;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      float ok[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kt = 0; kt < KT2; ++kt) {
        uint32_t bq[2];
        ld_b_kn(bq, sQ, dt * 8, kt * 16, lane);    // B[k = i][n = d] = Q[i][d]
        mma16816(ok, da[kt], bq);
      }
      const int c = dt * 8 + 2 * t;
      if (j0 < p.Sk) *reinterpret_cast<uint32_t*>(ko + j0 * p.dk_ss + c) = pack_bf16x2(ok[0], ok[1]);
      if (j1 < p.Sk) *reinterpret_cast<uint32_t*>(ko + j1 * p.dk_ss + c) = pack_bf16x2(ok[2], ok[3]);
    }
  }
}

template <int KT2>
static int launch_attn_bwd_p(const AttnParams& p, cudaStream_t st) {
  using L = AttnSmemP<KT2>;
  auto kern = attn_bwd_p_kernel<KT2>;
  static bool configured = false;
  if (!configured) {
    CRV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  CRV_CUDA(launch_pdl(kern, dim3(p.B * p.heads), dim3(KT2 * 32), L::kTotal, st, p));
  return launch_status();
}

template <int KT2, int PAIRS, bool BWD>
static int launch_attn(const AttnParams& p, cudaStream_t st) {
  using L = AttnSmem<KT2>;
  constexpr int smem = PAIRS * (BWD ? L::kBwd : L::kFwd);
  auto kern = BWD ? attn_bwd_kernel<KT2, PAIRS> : attn_fwd_kernel<KT2, PAIRS>;
  static bool configured = false;
  if (!configured) {
    CRV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int pairs = p.B * p.heads;
  CRV_CUDA(launch_pdl(kern, dim3((pairs + PAIRS - 1) / PAIRS), dim3(PAIRS * KT2 * 32), smem, st, p));
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
  if (S <= 32) return launch_attn<2, 4, false>(p, st);
  if (S <= 48) return launch_attn<3, 2, false>(p, st);
  return launch_attn<4, 2, false>(p, st);
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
  if (S <= 32) return launch_attn<2, 1, true>(p, st);
  if (S <= 48) return launch_attn<3, 1, true>(p, st);
  return launch_attn<4, 1, true>(p, st);
}


// Training pair: the forward also writes the signed probabilities (probs [B * heads][Sq][skp] bf16, skp = Sk rounded up
// to a multiple of 8), the backward consumes them.  crv_attention_probs_pitch() gives skp.
extern "C" int crv_attention_probs_pitch(int Sk) { return Sk > 0 ? (Sk + 7) / 8 * 8 : CRV_E_BADARG; }

extern "C" int crv_attention_fwd_p(const uint16_t* q, long long q_bs, long long q_ss, const uint16_t* k, long long k_bs,
                                   long long k_ss, const uint16_t* v, long long v_bs, long long v_ss, const float* mask,
                                   uint16_t* out, uint16_t* probs, int B, int heads, int Sq, int Sk, float scale,
                                   float p_drop, const unsigned long long* rng_state, int site, void* stream) {
  AttnParams p{};
  p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.k = reinterpret_cast<const __nv_bfloat16*>(k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(v);
  p.q_bs = q_bs; p.q_ss = q_ss; p.k_bs = k_bs; p.k_ss = k_ss; p.v_bs = v_bs; p.v_ss = v_ss;
  p.mask = mask; p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.B = B; p.heads = heads; p.Sq = Sq; p.Sk = Sk; p.scale = scale; p.p_drop = p_drop; p.rng_state = rng_state; p.site = site;
  p.probs = reinterpret_cast<__nv_bfloat16*>(probs);
  p.skp = crv_attention_probs_pitch(Sk);
  int rc = check_attn(p);
  if (rc) return rc;
  if (!out || !aligned16(out) || !probs || !aligned16(probs)) return CRV_E_BADARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int S = Sq > Sk ? Sq : Sk;
  if (S <= 32) return launch_attn<2, 4, false>(p, st);
  if (S <= 48) return launch_attn<3, 2, false>(p, st);
  return launch_attn<4, 2, false>(p, st);
}

extern "C" int crv_attention_bwd_p(const uint16_t* q, long long q_bs, long long q_ss, const uint16_t* k, long long k_bs,
                                   long long k_ss, const uint16_t* v, long long v_bs, long long v_ss,
                                   const uint16_t* probs, const uint16_t* dout, uint16_t* dq, long long dq_bs,
                                   long long dq_ss, uint16_t* dk, long long dk_bs, long long dk_ss, uint16_t* dv,
                                   long long dv_bs, long long dv_ss, int B, int heads, int Sq, int Sk, float scale,
                                   float p_drop, void* stream) {
  AttnParams p{};
  p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.k = reinterpret_cast<const __nv_bfloat16*>(k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(v);
  p.q_bs = q_bs; p.q_ss = q_ss; p.k_bs = k_bs; p.k_ss = k_ss; p.v_bs = v_bs; p.v_ss = v_ss;
  p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq); p.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.dq_bs = dq_bs; p.dq_ss = dq_ss; p.dk_bs = dk_bs; p.dk_ss = dk_ss; p.dv_bs = dv_bs; p.dv_ss = dv_ss;
  p.B = B; p.heads = heads; p.Sq = Sq; p.Sk = Sk; p.scale = scale; p.p_drop = p_drop;
  p.probs = const_cast<__nv_bfloat16*>(reinterpret_cast<const __nv_bfloat16*>(probs));
  p.skp = crv_attention_probs_pitch(Sk);
  int rc = check_attn(p);
  if (rc) return rc;
  if (!dout || !dq || !dk || !dv || !probs || !aligned16(probs)) return CRV_E_BADARG;
  if (p_drop < 0.f || p_drop >= 1.f) return CRV_E_BADARG;
  if ((dq_ss | dk_ss | dv_ss | dq_bs | dk_bs | dv_bs) & 7) return CRV_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int S = Sq > Sk ? Sq : Sk;
  if (S <= 32) return launch_attn_bwd_p<2>(p, st);
  if (S <= 48) return launch_attn_bwd_p<3>(p, st);
  return launch_attn_bwd_p<4>(p, st);
}
