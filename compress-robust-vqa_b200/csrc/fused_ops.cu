// Fused elementwise kernels between the masked GEMMs of a transformer layer ("next" row f3 of SURVEY.md
// section 8: hg_transformers/modeling_lxmert.py:830-903).  They exist so that the activations flowing
// from one tcgen05 GEMM to the next are produced directly as bf16 operands, the residual stream stays
// fp32, and one pass over HBM replaces PyTorch's bias-add / dropout / add / LayerNorm / cast chain:
//
//   ln_fwd : z = dropout(g) + res ; y = LayerNorm(z) -> y (fp32) and y (bf16), per-row mean / rstd
//   ln_bwd : dz = LayerNorm'(dy32 + dy16) ; d_res = dz (fp32) ; d_g = dropout'(dz) (bf16)
//            (gamma / beta are frozen in stage 2, masking/maskers.py:564-569: no parameter gradients)
//   gelu   : y = gelu(u) (erf form, bf16 in / bf16 out) and du = dy * gelu'(u)
//
// Dropout is counter based: keep(i) = hash(seed, step counter, call-site id, element index) >= p * 2^32,
// with (seed, counter) read from device memory so a captured CUDA graph draws a fresh mask every replay
// and the backward pass regenerates the forward mask instead of storing it.
// One warp per row (H <= 1024, H % 128 == 0), 16-byte accesses, fp32 math.
#include <cstdlib>

#include "common.cuh"
#include "gelu.cuh"

namespace crv {

constexpr int kRowsPerBlock = 8;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

struct Rng {
  uint64_t key;
  uint32_t thresh;  // per 16-bit lane: drop when lane < thresh (p quantised to 1/65536)
  float scale;
  // one 64-bit hash decides the four elements of a float4 (element index e, e % 4 == 0)
  __device__ __forceinline__ void drop4(float4& v, int64_t e) const {
    const uint64_t h = mix64(key + static_cast<uint64_t>(e >> 2) * 0x9E3779B97F4A7C15ull);
    v.x = ((h)&0xFFFFu) >= thresh ? v.x * scale : 0.f;
    v.y = ((h >> 16) & 0xFFFFu) >= thresh ? v.y * scale : 0.f;
    v.z = ((h >> 32) & 0xFFFFu) >= thresh ? v.z * scale : 0.f;
    v.w = ((h >> 48) & 0xFFFFu) >= thresh ? v.w * scale : 0.f;
  }
};

__device__ __forceinline__ Rng make_rng(const unsigned long long* state, int site, float p) {
  Rng r;
  r.thresh = 0;
  r.scale = 1.f;
  r.key = 0;
  if (state != nullptr && p > 0.f) {
    const uint64_t seed = state[0], ctr = state[1];
    r.key = (seed * 0xD1342543DE82EF95ull) ^ (ctr * 0xA24BAED4963EE407ull) ^ (static_cast<uint64_t>(site) << 40);
    r.thresh = static_cast<uint32_t>(fminf(p, 0.9999f) * 65536.0f);
    if (r.thresh == 0) r.thresh = 1;
    r.scale = 1.f / (1.f - p);
  }
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 load4(const void* base, bool is_bf16, int64_t e) {
  if (is_bf16) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(base) + e));
    return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xFFFF0000u),
                       __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xFFFF0000u));
  }
  return __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(base) + e));
}
__device__ __forceinline__ uint2 pack4(float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

template <int VEC>  // VEC = H / 128 float4 chunks per lane
__global__ void __launch_bounds__(kRowsPerBlock * 32, VEC <= 6 ? 3 : 2)
ln_fwd_kernel(const void* __restrict__ g, int g_bf16, const float* __restrict__ res, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, float p, const unsigned long long* __restrict__ rng_state,
              int site, float* __restrict__ y32, uint16_t* __restrict__ y16, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, int M, int H) {
  pdl_wait();
  const int row = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const Rng rng = make_rng(rng_state, site, p);
  // every load of the row is issued before the first dependent instruction (the dropout hash between the loads
  // made ptxas serialise them: one global round trip per chunk)
  float4 z[VEC], rv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int64_t e = static_cast<int64_t>(row) * H + (i * 32 + lane) * 4;
    z[i] = load4(g, g_bf16, e);
    rv[i] = res ? __ldg(reinterpret_cast<const float4*>(res + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int64_t e = static_cast<int64_t>(row) * H + (i * 32 + lane) * 4;
    float4 gv = z[i];
    if (rng.thresh) rng.drop4(gv, e);
    gv.x += rv[i].x; gv.y += rv[i].y; gv.z += rv[i].z; gv.w += rv[i].w;
    z[i] = gv;
    s += gv.x + gv.y + gv.z + gv.w;
  }
  pdl_launch_dependents();     // inputs are in registers
  const float mean = warp_sum(s) / H;
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float a = z[i].x - mean, b = z[i].y - mean, c = z[i].z - mean, d = z[i].w - mean;
    v += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(v) / H + eps);
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    const int64_t e = static_cast<int64_t>(row) * H + col;
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + col));
    float4 o;
    o.x = (z[i].x - mean) * rstd * ga.x + be.x;
    o.y = (z[i].y - mean) * rstd * ga.y + be.y;
    o.z = (z[i].z - mean) * rstd * ga.z + be.z;
    o.w = (z[i].w - mean) * rstd * ga.w + be.w;
    if (y32) *reinterpret_cast<float4*>(y32 + e) = o;
    if (y16) *reinterpret_cast<uint2*>(y16 + e) = pack4(o);
  }
}

// PGRAD: also emit this CTA's partial sums of the LayerNorm parameter gradients (stage-3 fine-tune: gamma / beta are
// trained) -- part[blockIdx.x][0..H) = sum_rows d * xhat (dgamma), part[blockIdx.x][H..2H) = sum_rows d (dbeta), d =
// dy32 + dy16 -- which crv_partial_reduce adds up in index order.
template <int VEC, bool PGRAD>
__global__ void __launch_bounds__(kRowsPerBlock * 32, 2)
ln_bwd_kernel(const float* __restrict__ dy32, const uint16_t* __restrict__ dy16, const void* __restrict__ g, int g_bf16,
              const float* __restrict__ res, const float* __restrict__ gamma, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, float p, const unsigned long long* __restrict__ rng_state, int site,
              void* __restrict__ dg, int dg_bf16, float* __restrict__ dres, float* __restrict__ part, int M, int H) {
  pdl_wait();
  const int row = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const bool valid = row < M;
  if (!PGRAD && !valid) return;
  const int rrow = valid ? row : M - 1;
  const Rng rng = make_rng(rng_state, site, p);
  const float mean = mean_in[rrow], rstd = rstd_in[rrow];
  // all loads first (see ln_fwd_kernel)
  float4 xh[VEC], dyg[VEC], rv[VEC], d2[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int64_t e = static_cast<int64_t>(rrow) * H + (i * 32 + lane) * 4;
    xh[i] = load4(g, g_bf16, e);
    rv[i] = res ? __ldg(reinterpret_cast<const float4*>(res + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
    dyg[i] = dy32 ? __ldg(reinterpret_cast<const float4*>(dy32 + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
    d2[i] = dy16 ? load4(dy16, 1, e) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    const int64_t e = static_cast<int64_t>(rrow) * H + col;
    float4 gv = xh[i];
    if (rng.thresh) rng.drop4(gv, e);
    gv.x += rv[i].x; gv.y += rv[i].y; gv.z += rv[i].z; gv.w += rv[i].w;
    xh[i] = make_float4((gv.x - mean) * rstd, (gv.y - mean) * rstd, (gv.z - mean) * rstd, (gv.w - mean) * rstd);
    float4 d = dyg[i];
    d.x += d2[i].x; d.y += d2[i].y; d.z += d2[i].z; d.w += d2[i].w;
    if (PGRAD) d2[i] = valid ? d : make_float4(0.f, 0.f, 0.f, 0.f);      // the pre-gamma gradient, kept for dgamma / dbeta
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + col));
    d.x *= ga.x; d.y *= ga.y; d.z *= ga.z; d.w *= ga.w;
    dyg[i] = d;
    s1 += d.x + d.y + d.z + d.w;
    s2 += d.x * xh[i].x + d.y * xh[i].y + d.z * xh[i].z + d.w * xh[i].w;
  }
  pdl_launch_dependents();     // inputs are in registers
  const float m1 = warp_sum(s1) / H, m2 = warp_sum(s2) / H;
  if (valid) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int col = (i * 32 + lane) * 4;
      const int64_t e = static_cast<int64_t>(row) * H + col;
      float4 dz;
      dz.x = rstd * (dyg[i].x - m1 - xh[i].x * m2);
      dz.y = rstd * (dyg[i].y - m1 - xh[i].y * m2);
      dz.z = rstd * (dyg[i].z - m1 - xh[i].z * m2);
      dz.w = rstd * (dyg[i].w - m1 - xh[i].w * m2);
      if (dres) *reinterpret_cast<float4*>(dres + e) = dz;
      if (dg) {
        if (rng.thresh) rng.drop4(dz, e);
        if (dg_bf16) *reinterpret_cast<uint2*>(static_cast<uint16_t*>(dg) + e) = pack4(dz);
        else *reinterpret_cast<float4*>(static_cast<float*>(dg) + e) = dz;
      }
    }
  }
  if (PGRAD) {
    __shared__ float4 red[kRowsPerBlock][VEC * 32];
    const int w = threadIdx.x >> 5;
    float* out = part + static_cast<size_t>(blockIdx.x) * 2 * H;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
      for (int i = 0; i < VEC; ++i)
        red[w][i * 32 + lane] = pass == 0 ? make_float4(d2[i].x * xh[i].x, d2[i].y * xh[i].y, d2[i].z * xh[i].z,
                                                        d2[i].w * xh[i].w)
                                          : d2[i];
      __syncthreads();
      for (int c = threadIdx.x; c < VEC * 32; c += kRowsPerBlock * 32) {
        float4 t = red[0][c];
#pragma unroll
        for (int r = 1; r < kRowsPerBlock; ++r) {
          const float4 o = red[r][c];
          t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
        }
        *reinterpret_cast<float4*>(out + pass * H + c * 4) = t;
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm -> (average of two) -> dropout: the entry blocks of LXMERT, which normalise FIRST and drop after
//   visual feature encoder  y = dropout((LN_a(visn_fc(feats)) + LN_b(box_fc(pos))) / 2)   (modeling_lxmert.py:576-592)
//   embeddings              y = dropout(LN_a(word + position + token_type))               (modeling_lxmert.py:744-770)
// PyTorch runs them as 6 - 8 launches over the [M, H] activations; here one pass reads a (and b), writes y as fp32
// (residual stream) and bf16 (next GEMM operand) plus the row statistics; the backward regenerates the dropout mask
// and returns da (and db).  gamma / beta are frozen in stage 2, so there are no parameter gradients (stage 3 keeps
// the PyTorch path).  One warp per row, as ln_fwd / ln_bwd.
template <int VEC, bool DUAL>
__global__ void __launch_bounds__(kRowsPerBlock * 32, 2)
ln_avg_drop_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gamma_a,
                       const float* __restrict__ beta_a, const float* __restrict__ gamma_b,
                       const float* __restrict__ beta_b, float eps, float p,
                       const unsigned long long* __restrict__ rng_state, int site, float* __restrict__ y32,
                       uint16_t* __restrict__ y16, float* __restrict__ stats, int M, int H) {
  pdl_wait();
  const int row = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const Rng rng = make_rng(rng_state, site, p);
  float4 av[VEC], bv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int64_t e = static_cast<int64_t>(row) * H + (i * 32 + lane) * 4;
    av[i] = __ldg(reinterpret_cast<const float4*>(a + e));
    if (DUAL) bv[i] = __ldg(reinterpret_cast<const float4*>(b + e));
  }
  pdl_launch_dependents();
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    sa += av[i].x + av[i].y + av[i].z + av[i].w;
    if (DUAL) sb += bv[i].x + bv[i].y + bv[i].z + bv[i].w;
  }
  const float ma = warp_sum(sa) / H, mb = DUAL ? warp_sum(sb) / H : 0.f;
  float va = 0.f, vb = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    float x = av[i].x - ma, y = av[i].y - ma, z = av[i].z - ma, w = av[i].w - ma;
    va += x * x + y * y + z * z + w * w;
    if (DUAL) {
      x = bv[i].x - mb; y = bv[i].y - mb; z = bv[i].z - mb; w = bv[i].w - mb;
      vb += x * x + y * y + z * z + w * w;
    }
  }
  const float ra = rsqrtf(warp_sum(va) / H + eps), rb = DUAL ? rsqrtf(warp_sum(vb) / H + eps) : 0.f;
  if (lane == 0) *reinterpret_cast<float4*>(stats + static_cast<size_t>(row) * 4) = make_float4(ma, ra, mb, rb);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    const int64_t e = static_cast<int64_t>(row) * H + col;
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma_a + col));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta_a + col));
    float4 o;
    o.x = (av[i].x - ma) * ra * ga.x + be.x;
    o.y = (av[i].y - ma) * ra * ga.y + be.y;
    o.z = (av[i].z - ma) * ra * ga.z + be.z;
    o.w = (av[i].w - ma) * ra * ga.w + be.w;
    if (DUAL) {
      const float4 gb = __ldg(reinterpret_cast<const float4*>(gamma_b + col));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(beta_b + col));
      o.x = (o.x + (bv[i].x - mb) * rb * gb.x + bb.x) * 0.5f;     // (x + y) / 2 as the reference writes it
      o.y = (o.y + (bv[i].y - mb) * rb * gb.y + bb.y) * 0.5f;
      o.z = (o.z + (bv[i].z - mb) * rb * gb.z + bb.z) * 0.5f;
      o.w = (o.w + (bv[i].w - mb) * rb * gb.w + bb.w) * 0.5f;
    }
    if (rng.thresh) rng.drop4(o, e);
    *reinterpret_cast<float4*>(y32 + e) = o;
    *reinterpret_cast<uint2*>(y16 + e) = pack4(o);
  }
}

template <int VEC, bool DUAL>
__global__ void __launch_bounds__(kRowsPerBlock * 32, 2)
ln_avg_drop_bwd_kernel(const float* __restrict__ dy32, const uint16_t* __restrict__ dy16, const float* __restrict__ a,
                       const float* __restrict__ b, const float* __restrict__ gamma_a, const float* __restrict__ gamma_b,
                       const float* __restrict__ stats, float p, const unsigned long long* __restrict__ rng_state,
                       int site, float* __restrict__ da, float* __restrict__ db, int M, int H) {
  pdl_wait();
  const int row = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const Rng rng = make_rng(rng_state, site, p);
  const float4 st = __ldg(reinterpret_cast<const float4*>(stats + static_cast<size_t>(row) * 4));
  float4 d[VEC], xa[VEC], xb[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int64_t e = static_cast<int64_t>(row) * H + (i * 32 + lane) * 4;
    d[i] = dy32 ? __ldg(reinterpret_cast<const float4*>(dy32 + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (dy16) {
      const float4 t = load4(dy16, 1, e);
      d[i].x += t.x; d[i].y += t.y; d[i].z += t.z; d[i].w += t.w;
    }
    xa[i] = __ldg(reinterpret_cast<const float4*>(a + e));
    if (DUAL) xb[i] = __ldg(reinterpret_cast<const float4*>(b + e));
  }
  pdl_launch_dependents();
  float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    const int64_t e = static_cast<int64_t>(row) * H + col;
    if (rng.thresh) rng.drop4(d[i], e);          // dropout' = the same keep / scale pattern
    if (DUAL) { d[i].x *= 0.5f; d[i].y *= 0.5f; d[i].z *= 0.5f; d[i].w *= 0.5f; }
    xa[i] = make_float4((xa[i].x - st.x) * st.y, (xa[i].y - st.x) * st.y, (xa[i].z - st.x) * st.y, (xa[i].w - st.x) * st.y);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma_a + col));
    s1a += d[i].x * ga.x + d[i].y * ga.y + d[i].z * ga.z + d[i].w * ga.w;
    s2a += d[i].x * ga.x * xa[i].x + d[i].y * ga.y * xa[i].y + d[i].z * ga.z * xa[i].z + d[i].w * ga.w * xa[i].w;
    if (DUAL) {
      xb[i] = make_float4((xb[i].x - st.z) * st.w, (xb[i].y - st.z) * st.w, (xb[i].z - st.z) * st.w, (xb[i].w - st.z) * st.w);
      const float4 gb = __ldg(reinterpret_cast<const float4*>(gamma_b + col));
      s1b += d[i].x * gb.x + d[i].y * gb.y + d[i].z * gb.z + d[i].w * gb.w;
      s2b += d[i].x * gb.x * xb[i].x + d[i].y * gb.y * xb[i].y + d[i].z * gb.z * xb[i].z + d[i].w * gb.w * xb[i].w;
    }
  }
  const float m1a = warp_sum(s1a) / H, m2a = warp_sum(s2a) / H;
  const float m1b = DUAL ? warp_sum(s1b) / H : 0.f, m2b = DUAL ? warp_sum(s2b) / H : 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    const int64_t e = static_cast<int64_t>(row) * H + col;
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma_a + col));
    float4 o;
    o.x = st.y * (d[i].x * ga.x - m1a - xa[i].x * m2a);
    o.y = st.y * (d[i].y * ga.y - m1a - xa[i].y * m2a);
    o.z = st.y * (d[i].z * ga.z - m1a - xa[i].z * m2a);
    o.w = st.y * (d[i].w * ga.w - m1a - xa[i].w * m2a);
    if (da) *reinterpret_cast<float4*>(da + e) = o;
    if (DUAL && db) {
      const float4 gb = __ldg(reinterpret_cast<const float4*>(gamma_b + col));
      o.x = st.w * (d[i].x * gb.x - m1b - xb[i].x * m2b);
      o.y = st.w * (d[i].y * gb.y - m1b - xb[i].y * m2b);
      o.z = st.w * (d[i].z * gb.z - m1b - xb[i].z * m2b);
      o.w = st.w * (d[i].w * gb.w - m1b - xb[i].w * m2b);
      *reinterpret_cast<float4*>(db + e) = o;
    }
  }
}

__global__ void gelu_fwd_kernel(const uint16_t* __restrict__ u, uint16_t* __restrict__ y, int64_t n) {
  pdl_wait();
  const int64_t nvec = n >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec; i += stride) {
    const float4 a = load4(u, 1, i * 8), b = load4(u, 1, i * 8 + 4);
    const uint2 p0 = pack4(make_float4(gelu_f(a.x), gelu_f(a.y), gelu_f(a.z), gelu_f(a.w)));
    const uint2 p1 = pack4(make_float4(gelu_f(b.x), gelu_f(b.y), gelu_f(b.z), gelu_f(b.w)));
    reinterpret_cast<uint4*>(y)[i] = make_uint4(p0.x, p0.y, p1.x, p1.y);
  }
}

__global__ void gelu_bwd_kernel(const uint16_t* __restrict__ u, const uint16_t* __restrict__ dy,
                                uint16_t* __restrict__ du, int64_t n) {
  pdl_wait();
  const int64_t nvec = n >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec; i += stride) {
    const float4 a = load4(u, 1, i * 8), b = load4(u, 1, i * 8 + 4);
    const float4 c = load4(dy, 1, i * 8), d = load4(dy, 1, i * 8 + 4);
    const uint2 p0 = pack4(make_float4(c.x * gelu_grad_f(a.x), c.y * gelu_grad_f(a.y), c.z * gelu_grad_f(a.z),
                                       c.w * gelu_grad_f(a.w)));
    const uint2 p1 = pack4(make_float4(d.x * gelu_grad_f(b.x), d.y * gelu_grad_f(b.y), d.z * gelu_grad_f(b.z),
                                       d.w * gelu_grad_f(b.w)));
    reinterpret_cast<uint4*>(du)[i] = make_uint4(p0.x, p0.y, p1.x, p1.y);
  }
}

// QuickGELU of the CLIP vision tower (mPLUG/models/clip/model.py:25-27): y = x sigmoid(1.702 x) on bf16, and
// dx = dy * s (1 + 1.702 x (1 - s)), s = sigmoid(1.702 x).  PyTorch runs it as three elementwise kernels forward and
// five backward over the [tokens, 3072] MLP activation; one pass each way here.  exp through ex2 / rcp on the SFU.
__device__ __forceinline__ float sigmoid1702(float x) {
  return sfu_rcp(1.f + sfu_ex2(x * (-1.702f * 1.4426950408889634f)));
}
__device__ __forceinline__ float qgelu_f(float x) { return x * sigmoid1702(x); }
__device__ __forceinline__ float qgelu_grad_f(float x) {
  const float s = sigmoid1702(x);
  return s * fmaf(1.702f * x, 1.f - s, 1.f);
}
__global__ void qgelu_fwd_kernel(const uint16_t* __restrict__ u, uint16_t* __restrict__ y, int64_t n) {
  pdl_wait();
  const int64_t nvec = n >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec; i += stride) {
    const float4 a = load4(u, 1, i * 8), b = load4(u, 1, i * 8 + 4);
    const uint2 p0 = pack4(make_float4(qgelu_f(a.x), qgelu_f(a.y), qgelu_f(a.z), qgelu_f(a.w)));
    const uint2 p1 = pack4(make_float4(qgelu_f(b.x), qgelu_f(b.y), qgelu_f(b.z), qgelu_f(b.w)));
    reinterpret_cast<uint4*>(y)[i] = make_uint4(p0.x, p0.y, p1.x, p1.y);
  }
}
__global__ void qgelu_bwd_kernel(const uint16_t* __restrict__ u, const uint16_t* __restrict__ dy,
                                 uint16_t* __restrict__ du, int64_t n) {
  pdl_wait();
  const int64_t nvec = n >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec; i += stride) {
    const float4 a = load4(u, 1, i * 8), b = load4(u, 1, i * 8 + 4);
    const float4 c = load4(dy, 1, i * 8), d = load4(dy, 1, i * 8 + 4);
    const uint2 p0 = pack4(make_float4(c.x * qgelu_grad_f(a.x), c.y * qgelu_grad_f(a.y), c.z * qgelu_grad_f(a.z),
                                       c.w * qgelu_grad_f(a.w)));
    const uint2 p1 = pack4(make_float4(d.x * qgelu_grad_f(b.x), d.y * qgelu_grad_f(b.y), d.z * qgelu_grad_f(b.z),
                                       d.w * qgelu_grad_f(b.w)));
    reinterpret_cast<uint4*>(du)[i] = make_uint4(p0.x, p0.y, p1.x, p1.y);
  }
}

__global__ void counter_inc_kernel(unsigned long long* state) { state[1] += 1ull; }

template <typename F>
static int dispatch_vec(int H, F&& f) {
  switch (H / 128) {
    case 1: return f(std::integral_constant<int, 1>{});
    case 2: return f(std::integral_constant<int, 2>{});
    case 4: return f(std::integral_constant<int, 4>{});
    case 6: return f(std::integral_constant<int, 6>{});
    case 8: return f(std::integral_constant<int, 8>{});
    default: return CRV_E_SHAPE;
  }
}

}  // namespace crv

using namespace crv;

extern "C" int crv_ln_fwd(const void* g, int g_dtype, const float* res, const float* gamma, const float* beta,
                          float eps, float p_drop, const unsigned long long* rng_state, int site, float* y_f32,
                          uint16_t* y_bf16, float* mean, float* rstd, int M, int H, void* stream) {
  if (!g || !gamma || !beta || !mean || !rstd || M <= 0 || H <= 0) return CRV_E_BADARG;
  if (H % 128 || H > 1024) return CRV_E_SHAPE;
  if (!aligned16(g) || (res && !aligned16(res)) || (y_f32 && !aligned16(y_f32)) || (y_bf16 && !aligned16(y_bf16)))
    return CRV_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (M + kRowsPerBlock - 1) / kRowsPerBlock;
  return dispatch_vec(H, [&](auto v) {
    constexpr int VEC = decltype(v)::value;
    CRV_CUDA(launch_pdl(ln_fwd_kernel<VEC>, dim3(grid), dim3(kRowsPerBlock * 32), 0, st, g,
                        static_cast<int>(g_dtype == CRV_DTYPE_BF16), res, gamma, beta, eps, p_drop, rng_state, site, y_f32,
                        y_bf16, mean, rstd, M, H));
    return launch_status();
  });
}

extern "C" size_t crv_ln_bwd_partials_bytes(int M, int H) {
  return M > 0 && H > 0 ? static_cast<size_t>((M + kRowsPerBlock - 1) / kRowsPerBlock) * 2 * H * sizeof(float) : 0;
}

extern "C" int crv_ln_bwd(const float* dy_f32, const uint16_t* dy_bf16, const void* g, int g_dtype, const float* res,
                          const float* gamma, const float* mean, const float* rstd, float p_drop,
                          const unsigned long long* rng_state, int site, void* dg, int dg_dtype, float* dres,
                          float* param_partials, int M, int H, void* stream) {
  if ((!dy_f32 && !dy_bf16) || !g || !gamma || !mean || !rstd || M <= 0 || H <= 0) return CRV_E_BADARG;
  if (H % 128 || H > 1024) return CRV_E_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (M + kRowsPerBlock - 1) / kRowsPerBlock;
  return dispatch_vec(H, [&](auto v) {
    constexpr int VEC = decltype(v)::value;
    if (param_partials)
      CRV_CUDA(launch_pdl(ln_bwd_kernel<VEC, true>, dim3(grid), dim3(kRowsPerBlock * 32), 0, st, dy_f32, dy_bf16, g,
                          static_cast<int>(g_dtype == CRV_DTYPE_BF16), res, gamma, mean, rstd, p_drop, rng_state, site,
                          dg, static_cast<int>(dg_dtype == CRV_DTYPE_BF16), dres, param_partials, M, H));
    else
      CRV_CUDA(launch_pdl(ln_bwd_kernel<VEC, false>, dim3(grid), dim3(kRowsPerBlock * 32), 0, st, dy_f32, dy_bf16, g,
                          static_cast<int>(g_dtype == CRV_DTYPE_BF16), res, gamma, mean, rstd, p_drop, rng_state, site,
                          dg, static_cast<int>(dg_dtype == CRV_DTYPE_BF16), dres, param_partials, M, H));
    return launch_status();
  });
}

extern "C" int crv_gelu_fwd(const uint16_t* u, uint16_t* y, int64_t n, void* stream) {
  if (!u || !y || n < 0) return CRV_E_BADARG;
  if (n % 8) return CRV_E_SHAPE;
  if (n == 0) return CRV_OK;
  int64_t blocks = ((n >> 3) + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  CRV_CUDA(launch_pdl(gelu_fwd_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), u, y, n));
  return launch_status();
}

extern "C" int crv_gelu_bwd(const uint16_t* u, const uint16_t* dy, uint16_t* du, int64_t n, void* stream) {
  if (!u || !dy || !du || n < 0) return CRV_E_BADARG;
  if (n % 8) return CRV_E_SHAPE;
  if (n == 0) return CRV_OK;
  int64_t blocks = ((n >> 3) + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  CRV_CUDA(launch_pdl(gelu_bwd_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), u, dy, du, n));
  return launch_status();
}

extern "C" int crv_quick_gelu_fwd(const uint16_t* u, uint16_t* y, int64_t n, void* stream) {
  if (!u || !y || n < 0) return CRV_E_BADARG;
  if (n % 8) return CRV_E_SHAPE;
  if (n == 0) return CRV_OK;
  int64_t blocks = ((n >> 3) + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  CRV_CUDA(launch_pdl(qgelu_fwd_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), u, y, n));
  return launch_status();
}

extern "C" int crv_quick_gelu_bwd(const uint16_t* u, const uint16_t* dy, uint16_t* du, int64_t n, void* stream) {
  if (!u || !dy || !du || n < 0) return CRV_E_BADARG;
  if (n % 8) return CRV_E_SHAPE;
  if (n == 0) return CRV_OK;
  int64_t blocks = ((n >> 3) + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  CRV_CUDA(launch_pdl(qgelu_bwd_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), u, dy, du, n));
  return launch_status();
}

extern "C" int crv_rng_advance(unsigned long long* rng_state, void* stream) {
  if (!rng_state) return CRV_E_BADARG;
  counter_inc_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(rng_state);
  return launch_status();
}


extern "C" int crv_ln_avg_drop_fwd(const float* a, const float* b, const float* gamma_a, const float* beta_a,
                                   const float* gamma_b, const float* beta_b, float eps, float p_drop,
                                   const unsigned long long* rng_state, int site, float* y_f32, uint16_t* y_bf16,
                                   float* stats, int M, int H, void* stream) {
  if (!a || !gamma_a || !beta_a || !y_f32 || !y_bf16 || !stats || M <= 0 || H <= 0) return CRV_E_BADARG;
  if (b && (!gamma_b || !beta_b)) return CRV_E_BADARG;
  if (H % 128 || H > 1024) return CRV_E_SHAPE;
  if (!aligned16(a) || (b && !aligned16(b)) || !aligned16(y_f32) || !aligned16(y_bf16) || !aligned16(stats))
    return CRV_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (M + kRowsPerBlock - 1) / kRowsPerBlock;
  return dispatch_vec(H, [&](auto v) {
    constexpr int VEC = decltype(v)::value;
    if (b)
      CRV_CUDA(launch_pdl(ln_avg_drop_fwd_kernel<VEC, true>, dim3(grid), dim3(kRowsPerBlock * 32), 0, st, a, b, gamma_a,
                          beta_a, gamma_b, beta_b, eps, p_drop, rng_state, site, y_f32, y_bf16, stats, M, H));
    else
      CRV_CUDA(launch_pdl(ln_avg_drop_fwd_kernel<VEC, false>, dim3(grid), dim3(kRowsPerBlock * 32), 0, st, a, b, gamma_a,
                          beta_a, gamma_b, beta_b, eps, p_drop, rng_state, site, y_f32, y_bf16, stats, M, H));
    return launch_status();
  });
}

extern "C" int crv_ln_avg_drop_bwd(const float* dy_f32, const uint16_t* dy_bf16, const float* a, const float* b,
                                   const float* gamma_a, const float* gamma_b, const float* stats, float p_drop,
                                   const unsigned long long* rng_state, int site, float* da, float* db, int M, int H,
                                   void* stream) {
  if ((!dy_f32 && !dy_bf16) || !a || !gamma_a || !stats || (!da && !db) || M <= 0 || H <= 0) return CRV_E_BADARG;
  if (b && !gamma_b) return CRV_E_BADARG;
  if (db && !b) return CRV_E_BADARG;
  if (H % 128 || H > 1024) return CRV_E_SHAPE;
  if (!aligned16(a) || (b && !aligned16(b)) || (da && !aligned16(da)) || (db && !aligned16(db)) || !aligned16(stats))
    return CRV_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (M + kRowsPerBlock - 1) / kRowsPerBlock;
  return dispatch_vec(H, [&](auto v) {
    constexpr int VEC = decltype(v)::value;
    if (b)
      CRV_CUDA(launch_pdl(ln_avg_drop_bwd_kernel<VEC, true>, dim3(grid), dim3(kRowsPerBlock * 32), 0, st, dy_f32, dy_bf16,
                          a, b, gamma_a, gamma_b, stats, p_drop, rng_state, site, da, db, M, H));
    else
      CRV_CUDA(launch_pdl(ln_avg_drop_bwd_kernel<VEC, false>, dim3(grid), dim3(kRowsPerBlock * 32), 0, st, dy_f32, dy_bf16,
                          a, b, gamma_a, gamma_b, stats, p_drop, rng_state, site, da, db, M, H));
    return launch_status();
  });
}
