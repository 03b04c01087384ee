// Masked GEMM family for sm_100a: TMA -> (mask transform in shared memory) -> tcgen05.mma -> TMEM
// -> epilogue.  One kernel template covers the three GEMMs of a masked linear layer
// (reference: masking/maskers.py:359-366 and its autograd):
//
//   FWD : Y [M,N]  = X [M,K]  . (W (.) (S > thr))^T + b      A = X  K-major,  B = W  K-major  (+S)
//   DX  : dX[M,K]  = dY[M,N]  . (W (.) (S > thr))            A = dY K-major,  B = W  MN-major (+S)
//   DS  : dS[N,K] += (dY^T . X) (.) W                        A = dY MN-major, B = X  MN-major, epilogue (.)W
//
// Generic view used below:  D[MM,NN] = sum_kk A(mm,kk) * B(nn,kk).
//
// CTA = 128 x BN output tile, BK = 64 bf16 (one 128-byte swizzle row) per pipeline stage.
// Warp roles: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc), warps2-5 = epilogue
// (TMEM -> registers -> global), warps6-9 = mask transform (only when XFORM).
// Shared-memory operand tiles are in the canonical 128B-swizzled UMMA layouts written by TMA:
//   K-major  tile [R rows][64 k]   : R x 128 B, 8-row swizzle atoms, SBO = 1024 B
//   MN-major tile [64 kk][R mn]    : R/64 boxes of (64 kk rows x 128 B), LBO = 8192 B, SBO = 1024 B
// The score tile (fp32) is loaded by TMA as boxes of 32 floats (128 B) per row with the same
// swizzle, so a transform thread reads one 16-byte chunk of W (8 bf16) plus the two matching
// 16-byte chunks of S without bank conflicts, zeroes the masked-out bf16 lanes in place, then
// fences the generic-proxy writes towards the async proxy before the MMA warp is released.
#include "common.cuh"
#include "ptx.cuh"

namespace crv {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiStoreF32 = 0;
constexpr int kEpiStoreBF16 = 1;
constexpr int kEpiScoreGrad = 2;

struct GemmParams {
  int MM, NN, KK;          // generic problem extents
  int kb_per_split;        // k-blocks handled by one blockIdx.z
  const float* thr;        // device scalar (XFORM)
  const float* bias;       // [NN] or null (store epilogues)
  void* out;               // D, row-major [MM, NN]
  const __nv_bfloat16* w;  // [MM, NN] bf16 multiplier (score-grad epilogue)
  int atomic_out;          // score-grad: 1 = red.add into out, 0 = plain store
};

template <int BN, bool XFORM>
struct SmemLayout {
  static constexpr int kA = BM * BK * 2;                 // 16 KB
  static constexpr int kB = BN * BK * 2;                 // 16 / 32 KB
  static constexpr int kS = XFORM ? BN * BK * 4 : 0;     // 32 KB
  static constexpr int kStage = kA + kB + kS;
  static constexpr int kStages = XFORM ? 3 : (BN == 256 ? 4 : 6);
  static constexpr int kBarBytes = 1024;
  static constexpr int kTotal = kStages * kStage + kBarBytes + 1024;  // + alignment slack
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

template <bool A_MN, bool B_MN, int BN, bool XFORM, int EPI>
__global__ void __launch_bounds__(XFORM ? 320 : 192, 1)
masked_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmS, const GemmParams p) {
  using L = SmemLayout<BN, XFORM>;
  constexpr int STAGES = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + STAGES * L::kStage;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* xform_bar = empty_bar + STAGES;
  uint64_t* tmem_full_bar = xform_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int num_kb_total = (p.KK + BK - 1) / BK;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(kb_begin + p.kb_per_split, num_kb_total);
  const int num_kb = kb_end - kb_begin;
  if (num_kb <= 0) return;  // uniform per CTA: an empty split contributes nothing

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (XFORM) tma_prefetch_desc(&tmS);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&xform_bar[s], 128);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], L::kStage);
        uint8_t* sA = smem + s * L::kStage;
        uint8_t* sB = sA + L::kA;
        uint8_t* sS = sB + L::kB;
        const int kk0 = (kb_begin + it) * BK;
        if (!A_MN) {
          tma_load_2d(sA, &tmA, &full_bar[s], kk0, m0);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) tma_load_2d(sA + j * 8192, &tmA, &full_bar[s], m0 + 64 * j, kk0);
        }
        if (!B_MN) {
          tma_load_2d(sB, &tmB, &full_bar[s], kk0, n0);
          if (XFORM) {
#pragma unroll
            for (int h = 0; h < 2; ++h) tma_load_2d(sS + h * (BN * 128), &tmS, &full_bar[s], kk0 + 32 * h, n0);
          }
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + j * 8192, &tmB, &full_bar[s], n0 + 64 * j, kk0);
          if (XFORM) {
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) tma_load_2d(sS + j * 8192, &tmS, &full_bar[s], n0 + 32 * j, kk0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    for (int it = 0; it < num_kb; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      if (XFORM) mbar_wait(&xform_bar[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t aBase = smem_u32(smem + s * L::kStage);
        const uint32_t bBase = aBase + L::kA;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t da = A_MN ? make_sw128_desc(aBase + k * 2048, 8192, 1024)
                                   : make_sw128_desc(aBase + k * 32, 16, 1024);
          const uint64_t db = B_MN ? make_sw128_desc(bBase + k * 2048, 8192, 1024)
                                   : make_sw128_desc(bBase + k * 32, 16, 1024);
          umma_bf16(tmem_base, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);                       // frees the smem slot when these MMAs retire
        if (it == num_kb - 1) umma_commit(tmem_full_bar);  // accumulator complete
      }
      __syncwarp();
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int m = m0 + row;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      const int nb = n0 + c * 32;
      if (m < p.MM && nb < p.NN) {
        const bool full = (nb + 32 <= p.NN) && ((p.NN & 3) == 0);
        if (EPI == kEpiStoreF32) {
          float* o = static_cast<float*>(p.out) + static_cast<size_t>(m) * p.NN + nb;
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 v;
              v.x = __uint_as_float(r[j]);
              v.y = __uint_as_float(r[j + 1]);
              v.z = __uint_as_float(r[j + 2]);
              v.w = __uint_as_float(r[j + 3]);
              if (p.bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + nb + j));
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
              }
              *reinterpret_cast<float4*>(o + j) = v;
            }
          } else {
            for (int j = 0; j < 32 && nb + j < p.NN; ++j)
              o[j] = __uint_as_float(r[j]) + (p.bias ? p.bias[nb + j] : 0.f);
          }
        } else if (EPI == kEpiStoreBF16) {
          __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(m) * p.NN + nb;
          if (full && (p.NN & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float lo = __uint_as_float(r[j + 2 * e]);
                float hi = __uint_as_float(r[j + 2 * e + 1]);
                if (p.bias) { lo += __ldg(p.bias + nb + j + 2 * e); hi += __ldg(p.bias + nb + j + 2 * e + 1); }
                __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
                pk[e] = *reinterpret_cast<uint32_t*>(&t);
              }
              *reinterpret_cast<uint4*>(o + j) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          } else {
            for (int j = 0; j < 32 && nb + j < p.NN; ++j)
              o[j] = __float2bfloat16_rn(__uint_as_float(r[j]) + (p.bias ? p.bias[nb + j] : 0.f));
          }
        } else {  // kEpiScoreGrad: (.) W then accumulate / store
          float* o = static_cast<float*>(p.out) + static_cast<size_t>(m) * p.NN + nb;
          const __nv_bfloat16* wrow = p.w + static_cast<size_t>(m) * p.NN + nb;
          if (full && (p.NN & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wrow + j));
              const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
              float v[8];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[2 * e] = __uint_as_float(r[j + 2 * e]) * __uint_as_float(ww[e] << 16);
                v[2 * e + 1] = __uint_as_float(r[j + 2 * e + 1]) * __uint_as_float(ww[e] & 0xFFFF0000u);
              }
              if (p.atomic_out) {
                red_add_v4(o + j, v[0], v[1], v[2], v[3]);
                red_add_v4(o + j + 4, v[4], v[5], v[6], v[7]);
              } else {
                *reinterpret_cast<float4*>(o + j) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(o + j + 4) = make_float4(v[4], v[5], v[6], v[7]);
              }
            }
          } else {
            for (int j = 0; j < 32 && nb + j < p.NN; ++j) {
              const float v = __uint_as_float(r[j]) * __bfloat162float(wrow[j]);
              if (p.atomic_out) atomicAdd(o + j, v); else o[j] = v;
            }
          }
        }
      }
    }
  } else if (XFORM) {
    // ------------------------------------------------------------------ mask transform (128 threads)
    const int tid = threadIdx.x - 192;
    const float thr = __ldg(p.thr);
    for (int it = 0; it < num_kb; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      uint8_t* sB = smem + s * L::kStage + L::kA;
      uint8_t* sS = sB + L::kB;
      mbar_wait(&full_bar[s], ph);
#pragma unroll
      for (int i = 0; i < BN * 8 / 128; ++i) {
        const int qd = i * 128 + tid;
        const int c = qd & 7;     // 16-byte chunk of the 128-byte W row (8 bf16)
        const int rr = qd >> 3;   // row over the whole tile
        int r, wbox_off, sbox_off;
        if (!B_MN) {              // W: one box of BN rows; S: two boxes (k halves) of BN rows
          r = rr;
          wbox_off = 0;
          sbox_off = (c >> 2) * (BN * 128);
        } else {                  // W: BN/64 boxes of 64 rows; S: BN/32 boxes of 64 rows
          r = rr & 63;
          wbox_off = (rr >> 6) * 8192;
          sbox_off = ((rr >> 6) * 2 + (c >> 2)) * 8192;
        }
        const int sw = r & 7;
        uint4* wp = reinterpret_cast<uint4*>(sB + wbox_off + r * 128 + ((c ^ sw) << 4));
        const int j0 = (c & 3) * 2;
        // lanes with c >= 4 fetch the odd chunk first so a quarter-warp touches 8 distinct banksets
        const int first = j0 + (c >> 2);
        const int second = j0 + 1 - (c >> 2);
        const uint8_t* srow = sS + sbox_off + r * 128;
        const float4 f1 = *reinterpret_cast<const float4*>(srow + ((first ^ sw) << 4));
        const float4 f2 = *reinterpret_cast<const float4*>(srow + ((second ^ sw) << 4));
        const float4 lo = (c >> 2) ? f2 : f1;  // scores of elements 0..3
        const float4 hi = (c >> 2) ? f1 : f2;  // scores of elements 4..7
        uint4 w = *wp;
        w.x &= (lo.x > thr ? 0x0000FFFFu : 0u) | (lo.y > thr ? 0xFFFF0000u : 0u);
        w.y &= (lo.z > thr ? 0x0000FFFFu : 0u) | (lo.w > thr ? 0xFFFF0000u : 0u);
        w.z &= (hi.x > thr ? 0x0000FFFFu : 0u) | (hi.y > thr ? 0xFFFF0000u : 0u);
        w.w &= (hi.z > thr ? 0x0000FFFFu : 0u) | (hi.w > thr ? 0xFFFF0000u : 0u);
        *wp = w;
      }
      fence_proxy_async_smem();
      mbar_arrive(&xform_bar[s]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2-D row-major tensor [rows][cols] of elem_bytes elements; box = [box_rows][box_cols], 128B swizzle.
static int make_map(CUtensorMap* map, const void* base, int elem_bytes, int64_t rows, int64_t cols,
                    int box_rows, int box_cols) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return CRV_E_DRIVER;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_cuda_error = static_cast<int>(r);
    return static_cast<int>(r);
  }
  return CRV_OK;
}

template <bool A_MN, bool B_MN, int BN, bool XFORM, int EPI>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmS, const GemmParams& p,
                  int splits, cudaStream_t stream) {
  using L = SmemLayout<BN, XFORM>;
  auto kern = masked_gemm_kernel<A_MN, B_MN, BN, XFORM, EPI>;
  static bool configured = false;
  if (!configured) {
    CRV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  dim3 grid((p.NN + BN - 1) / BN, (p.MM + BM - 1) / BM, splits);
  kern<<<grid, XFORM ? 320 : 192, L::kTotal, stream>>>(tmA, tmB, tmS, p);
  return launch_status();
}

static int pick_bn(int MM, int NN) {
  // 256-wide tiles halve the A re-reads; use them when the grid still fills the machine
  const int tiles256 = ((NN + 255) / 256) * ((MM + BM - 1) / BM);
  return (NN % 256 == 0 && tiles256 >= num_sms()) ? 256 : 128;
}

}  // namespace crv

using namespace crv;

extern "C" int crv_masked_linear_fwd(const uint16_t* x, const uint16_t* w, const float* scores, const float* thr,
                                     const float* bias, void* y, int y_dtype, int M, int N, int K, void* stream) {
  if (!x || !w || !y || M <= 0 || N <= 0 || K <= 0) return CRV_E_BADARG;
  if (scores && !thr) return CRV_E_BADARG;
  if (K % 8) return CRV_E_SHAPE;
  if (!aligned16(x) || !aligned16(w) || !aligned16(y) || (scores && !aligned16(scores))) return CRV_E_ALIGN;
  if (y_dtype != CRV_DTYPE_F32 && y_dtype != CRV_DTYPE_BF16) return CRV_E_BADARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GemmParams p{};
  p.MM = M; p.NN = N; p.KK = K;
  p.kb_per_split = (K + BK - 1) / BK;
  p.thr = thr; p.bias = bias; p.out = y;
  CUtensorMap tmA, tmB, tmS;
  int rc;
  if ((rc = make_map(&tmA, x, 2, M, K, BM, BK))) return rc;
  if (scores) {
    if ((rc = make_map(&tmB, w, 2, N, K, 128, BK))) return rc;
    if ((rc = make_map(&tmS, scores, 4, N, K, 128, 32))) return rc;
    return y_dtype == CRV_DTYPE_F32 ? launch<false, false, 128, true, kEpiStoreF32>(tmA, tmB, tmS, p, 1, st)
                                    : launch<false, false, 128, true, kEpiStoreBF16>(tmA, tmB, tmS, p, 1, st);
  }
  const int bn = pick_bn(M, N);
  if ((rc = make_map(&tmB, w, 2, N, K, bn, BK))) return rc;
  tmS = tmB;
  if (bn == 256)
    return y_dtype == CRV_DTYPE_F32 ? launch<false, false, 256, false, kEpiStoreF32>(tmA, tmB, tmS, p, 1, st)
                                    : launch<false, false, 256, false, kEpiStoreBF16>(tmA, tmB, tmS, p, 1, st);
  return y_dtype == CRV_DTYPE_F32 ? launch<false, false, 128, false, kEpiStoreF32>(tmA, tmB, tmS, p, 1, st)
                                  : launch<false, false, 128, false, kEpiStoreBF16>(tmA, tmB, tmS, p, 1, st);
}

extern "C" int crv_masked_linear_bwd_dx(const uint16_t* dy, const uint16_t* w, const float* scores,
                                        const float* thr, void* dx, int dx_dtype, int M, int N, int K,
                                        void* stream) {
  if (!dy || !w || !dx || M <= 0 || N <= 0 || K <= 0) return CRV_E_BADARG;
  if (scores && !thr) return CRV_E_BADARG;
  if ((N % 8) || (K % 8)) return CRV_E_SHAPE;
  if (!aligned16(dy) || !aligned16(w) || !aligned16(dx) || (scores && !aligned16(scores))) return CRV_E_ALIGN;
  if (dx_dtype != CRV_DTYPE_F32 && dx_dtype != CRV_DTYPE_BF16) return CRV_E_BADARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // generic: MM = M, NN = K, KK = N;  A = dY [M][N] K-major;  B = W [N][K] = [KK][NN] MN-major
  GemmParams p{};
  p.MM = M; p.NN = K; p.KK = N;
  p.kb_per_split = (N + BK - 1) / BK;
  p.thr = thr; p.bias = nullptr; p.out = dx;
  CUtensorMap tmA, tmB, tmS;
  int rc;
  if ((rc = make_map(&tmA, dy, 2, M, N, BM, BK))) return rc;
  if ((rc = make_map(&tmB, w, 2, N, K, 64, 64))) return rc;
  if (scores) {
    if ((rc = make_map(&tmS, scores, 4, N, K, 64, 32))) return rc;
    return dx_dtype == CRV_DTYPE_F32 ? launch<false, true, 128, true, kEpiStoreF32>(tmA, tmB, tmS, p, 1, st)
                                     : launch<false, true, 128, true, kEpiStoreBF16>(tmA, tmB, tmS, p, 1, st);
  }
  tmS = tmB;
  if (pick_bn(M, K) == 256)
    return dx_dtype == CRV_DTYPE_F32 ? launch<false, true, 256, false, kEpiStoreF32>(tmA, tmB, tmS, p, 1, st)
                                     : launch<false, true, 256, false, kEpiStoreBF16>(tmA, tmB, tmS, p, 1, st);
  return dx_dtype == CRV_DTYPE_F32 ? launch<false, true, 128, false, kEpiStoreF32>(tmA, tmB, tmS, p, 1, st)
                                   : launch<false, true, 128, false, kEpiStoreBF16>(tmA, tmB, tmS, p, 1, st);
}

extern "C" int crv_masked_linear_bwd_ds(const uint16_t* dy, const uint16_t* x, const uint16_t* w, float* dscores,
                                        int accumulate, int M, int N, int K, void* stream) {
  if (!dy || !x || !w || !dscores || M <= 0 || N <= 0 || K <= 0) return CRV_E_BADARG;
  if ((N % 8) || (K % 8)) return CRV_E_SHAPE;
  if (!aligned16(dy) || !aligned16(x) || !aligned16(w) || !aligned16(dscores)) return CRV_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // generic: MM = N, NN = K, KK = M;  A = dY [M][N] = [KK][MM] MN-major;  B = X [M][K] = [KK][NN] MN-major
  const int num_kb = (M + BK - 1) / BK;
  const int tiles = ((N + BM - 1) / BM) * ((K + 127) / 128);
  int splits = num_sms() / tiles;  // one wave of CTAs; every extra split costs one more fp32 red pass
  if (splits < 1) splits = 1;
  const int max_splits = (num_kb + 7) / 8;  // keep >= 8 k-blocks per split so the atomics stay a minor cost
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  GemmParams p{};
  p.MM = N; p.NN = K; p.KK = M;
  p.kb_per_split = (num_kb + splits - 1) / splits;
  splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.out = dscores; p.w = reinterpret_cast<const __nv_bfloat16*>(w);
  p.atomic_out = (splits > 1 || accumulate) ? 1 : 0;
  if (splits > 1 && !accumulate)
    CRV_CUDA(cudaMemsetAsync(dscores, 0, static_cast<size_t>(N) * K * sizeof(float), st));
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_map(&tmA, dy, 2, M, N, 64, 64))) return rc;
  if ((rc = make_map(&tmB, x, 2, M, K, 64, 64))) return rc;
  return launch<true, true, 128, false, kEpiScoreGrad>(tmA, tmB, tmB, p, splits, st);
}
