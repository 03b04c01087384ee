// Masked GEMM family for sm_100a: TMA -> (mask transform in shared memory) -> tcgen05.mma -> TMEM
// -> epilogue -> TMA store.  One persistent, warp-specialised kernel template covers the three GEMMs
// of a masked linear layer (reference: masking/maskers.py:359-366 and its autograd):
//
//   FWD : Y [M,N]  = X [M,K]  . (W (.) (S > thr))^T + b      A = X  K-major,  B = W  K-major  (+S)
//   DX  : dX[M,K]  = dY[M,N]  . (W (.) (S > thr))            A = dY K-major,  B = W  MN-major (+S)
//   DS  : dS[N,K] += (dY^T . X) (.) W                        A = dY MN-major, B = X  MN-major, epilogue (.)W
//
// Generic view used below:  D[MM,NN] = sum_kk A(mm,kk) * B(nn,kk).
//
// Grid = min(#tiles, #SMs); every CTA walks tiles t = blockIdx.x, +gridDim.x, ... (n fastest).  Tile =
// 128 x BN outputs, BK = 64 bf16 (one 128-byte swizzle row) per pipeline stage.  Warp roles:
//   warp 0      TMA producer (runs ahead across tile boundaries through the smem ring)
//   warp 1      MMA issuer, one thread; accumulators double-buffered in TMEM (2 x BN columns) so the
//               next tile's mainloop overlaps this tile's epilogue
//   warps 2-5   epilogue: tcgen05.ld -> registers -> (+bias | (.)W) -> 128B-swizzled smem staging ->
//               TMA store (or TMA reduce-add for split / accumulating dS); TMA clips partial tiles
//   warps 6-9   mask transform (XFORM only): AND the bf16 lanes of the W tile with (S > thr) in place
// Shared-memory operand tiles are the canonical 128B-swizzled UMMA layouts written by TMA:
//   K-major  tile [R rows][64 k]   : R x 128 B, 8-row swizzle atoms, SBO = 1024 B
//   MN-major tile [64 kk][R mn]    : R/64 boxes of (64 kk rows x 128 B), LBO = 8192 B, SBO = 1024 B
// The score tile (fp32) arrives as boxes of 32 floats (128 B) per row with the same swizzle, so a
// transform thread reads one 16-byte chunk of W (8 bf16) plus the two matching 16-byte chunks of S
// without bank conflicts, then fences its generic-proxy writes towards the async proxy.
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "gelu.cuh"
#include "ptx.cuh"

namespace crv {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiStore = 0;      // D (+ bias) -> out
constexpr int kEpiScoreGrad = 2;  // D (.) W -> out (store or reduce-add)

struct GemmParams {
  int MM, NN, KK;          // generic problem extents
  int kb_per_split;        // k-blocks handled by one split
  int splits;              // reduction splits (score-grad only)
  int num_m, num_n;        // tile counts
  const float* thr;        // device scalar (XFORM)
  const float* bias;       // [NN] or null (store epilogue)
  const float* w;          // [MM, NN] fp32 multiplier (score-grad epilogue): the reference's dS = dM * W is fp32
  int reduce_out;          // score-grad: 1 = TMA reduce-add into out, 0 = plain TMA store
  long long* dbg;          // optional per-CTA timestamps (crv_gemm_debug_timestamps), else null
  const uint8_t* bits;     // 2-CTA mask transform: tile-contiguous mask bits (binarize_bits_tiled_kernel), 1 KB per half tile
  int bits_pitch;          // tiles per tile-row of that layout
};

template <int BN, bool XFORM>
struct SmemLayout {
  static constexpr int kA = BM * BK * 2;                 // 16 KB
  static constexpr int kB = BN * BK * 2;                 // 16 / 32 KB
  static constexpr int kS = XFORM ? BN * BK * 4 : 0;     // 32 KB
  static constexpr int kStage = kA + kB + kS;
  static constexpr int kStages = XFORM ? 3 : (BN == 256 ? 4 : 6);
  static constexpr int kEpi = 4 * 2 * 4096;              // 4 warps x 2 staging buffers x (32 rows x 128 B)
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kStages * kStage + kEpi + kBarBytes + 1024;  // + alignment slack
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// contiguous global -> shared bulk copy (16-byte aligned, size a multiple of 16) completing on an mbarrier of this CTA
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct TileCoord {
  int m0, n0, kb_begin, num_kb;
};

__device__ __forceinline__ TileCoord tile_coord(const GemmParams& p, int tile, int bn) {
  const int per_z = p.num_m * p.num_n;
  const int z = tile / per_z;
  const int r = tile - z * per_z;
  const int mt = r / p.num_n;
  const int nt = r - mt * p.num_n;
  const int total_kb = (p.KK + BK - 1) / BK;
  TileCoord t;
  t.m0 = mt * BM;
  t.n0 = nt * bn;
  t.kb_begin = z * p.kb_per_split;
  int e = t.kb_begin + p.kb_per_split;
  if (e > total_kb) e = total_kb;
  t.num_kb = e - t.kb_begin;  // host guarantees >= 1
  return t;
}

template <bool A_MN, bool B_MN, int BN, bool XFORM, int EPI, bool OUT_BF16>
__global__ void __launch_bounds__(XFORM ? 320 : 192, 1)
masked_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmOut,
                   const GemmParams p) {
  using L = SmemLayout<BN, XFORM>;
  constexpr int STAGES = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_base = smem + STAGES * L::kStage;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_base + L::kEpi);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* xform_bar = empty_bar + STAGES;
  uint64_t* tmem_full_bar = xform_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m * p.num_n * p.splits;
  if (p.dbg && threadIdx.x == 0) p.dbg[static_cast<size_t>(blockIdx.x) * 8 + 7] = gtime();  // kernel entry

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    if (XFORM) tma_prefetch_desc(&tmS);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&xform_bar[s], 4);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // barriers, TMEM and descriptors are set up; operands of a predecessor from here on
  long long* dbg = p.dbg ? p.dbg + static_cast<size_t>(blockIdx.x) * 8 : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = gtime();  // setup done

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord tc = tile_coord(p, tile, BN);
        for (int kb = 0; kb < tc.num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], L::kStage);
          uint8_t* sA = smem + s * L::kStage;
          uint8_t* sB = sA + L::kA;
          uint8_t* sS = sB + L::kB;
          const int kk0 = (tc.kb_begin + kb) * BK;
          if (dbg && it == 0) dbg[1] = gtime();  // first TMA issue
          if (!A_MN) {
            tma_load_2d(sA, &tmA, &full_bar[s], kk0, tc.m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sA + j * 8192, &tmA, &full_bar[s], tc.m0 + 64 * j, kk0);
          }
          if (!B_MN) {
            tma_load_2d(sB, &tmB, &full_bar[s], kk0, tc.n0);
            if (XFORM) {
#pragma unroll
              for (int h = 0; h < 2; ++h)
                tma_load_2d(sS + h * (BN * 128), &tmS, &full_bar[s], kk0 + 32 * h, tc.n0);
            }
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + j * 8192, &tmB, &full_bar[s], tc.n0 + 64 * j, kk0);
            if (XFORM) {
#pragma unroll
              for (int j = 0; j < BN / 32; ++j) tma_load_2d(sS + j * 8192, &tmS, &full_bar[s], tc.n0 + 32 * j, kk0);
            }
          }
        }
      }
      pdl_launch_dependents();   // every load of this CTA is in flight: let the next grid be scheduled
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
      const TileCoord tc = tile_coord(p, tile, BN);
      const int acc = tcount & 1;
      mbar_wait(&tmem_empty_bar[acc], ((tcount >> 1) & 1) ^ 1);  // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < tc.num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        if (XFORM) mbar_wait(&xform_bar[s], ph);
        tc_fence_after();
        if (dbg && it == 0 && lane == 0) dbg[2] = gtime();  // first stage landed
        if (lane == 0) {
          const uint32_t aBase = smem_u32(smem + s * L::kStage);
          const uint32_t bBase = aBase + L::kA;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? make_sw128_desc(aBase + k * 2048, 8192, 1024)
                                     : make_sw128_desc(aBase + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_sw128_desc(bBase + k * 2048, 8192, 1024)
                                     : make_sw128_desc(bBase + k * 32, 16, 1024);
            umma_bf16(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);                                // frees the smem slot when these MMAs retire
          if (kb == tc.num_kb - 1) umma_commit(&tmem_full_bar[acc]);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint8_t* stage_buf = epi_base + (warp - 2) * 8192;
    constexpr int COLS = OUT_BF16 ? 64 : 32;  // columns per 128-byte staging row
    int tcount = 0, chunk_no = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
      const TileCoord tc = tile_coord(p, tile, BN);
      const int acc = tcount & 1;
      const int row0 = tc.m0 + q * 32;
      const int m = row0 + lane;
      mbar_wait(&tmem_full_bar[acc], (tcount >> 1) & 1);
      tc_fence_after();
      if (dbg && tcount == 0 && warp == 2 && lane == 0) dbg[3] = gtime();  // first accumulator ready
      const uint32_t t_addr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / COLS; ++c, ++chunk_no) {
        const int nb = tc.n0 + c * COLS;
        uint32_t r[32];
        uint32_t pk[32];  // staging row as 32 words (fp32) or 32 packed bf16 pairs
        // bias of this chunk: lane l fetches columns nb + l (and nb + 32 + l); broadcast by shuffle below.
        // Issued before the TMEM wait so the two latencies overlap.
        float bias_lo = 0.f, bias_hi = 0.f;
        if (EPI == kEpiStore && p.bias != nullptr) {
          if (nb + lane < p.NN) bias_lo = __ldg(p.bias + nb + lane);
          if (OUT_BF16 && nb + 32 + lane < p.NN) bias_hi = __ldg(p.bias + nb + 32 + lane);
        }
        tmem_ld_32x32(t_addr + c * COLS, r);
        if (OUT_BF16) {
          uint32_t r2[32];
          tmem_ld_32x32(t_addr + c * COLS + 32, r2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a0 = __uint_as_float(r[2 * j]) + __shfl_sync(0xffffffffu, bias_lo, 2 * j);
            const float a1 = __uint_as_float(r[2 * j + 1]) + __shfl_sync(0xffffffffu, bias_lo, 2 * j + 1);
            const float b0 = __uint_as_float(r2[2 * j]) + __shfl_sync(0xffffffffu, bias_hi, 2 * j);
            const float b1 = __uint_as_float(r2[2 * j + 1]) + __shfl_sync(0xffffffffu, bias_hi, 2 * j + 1);
            __nv_bfloat162 ta = __floats2bfloat162_rn(a0, a1);
            __nv_bfloat162 tb = __floats2bfloat162_rn(b0, b1);
            pk[j] = *reinterpret_cast<uint32_t*>(&ta);
            pk[16 + j] = *reinterpret_cast<uint32_t*>(&tb);
          }
        } else if (EPI == kEpiStore) {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            pk[j] = __float_as_uint(__uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bias_lo, j));
        } else {  // score gradient: (.) W, W in fp32 as the reference multiplies (masking/maskers.py:337-339,365-366)
          tmem_ld_wait();
          const bool row_ok = m < p.MM;
          const float* wrow = p.w + static_cast<size_t>(row_ok ? m : 0) * p.NN + nb;
          if (row_ok && nb + 32 <= p.NN && (p.NN & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow) + j);
              pk[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) * wv.x);
              pk[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) * wv.y);
              pk[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) * wv.z);
              pk[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) * wv.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float wv = (row_ok && nb + j < p.NN) ? __ldg(wrow + j) : 0.f;
              pk[j] = __float_as_uint(__uint_as_float(r[j]) * wv);
            }
          }
        }
        // staging buffer (double-buffered per warp): wait until the TMA store issued two chunks ago has read it
        uint8_t* buf = stage_buf + (chunk_no & 1) * 4096;
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        const int sw = lane & 7;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(buf + lane * 128 + ((j ^ sw) << 4)) =
              make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && nb < p.NN && row0 < p.MM) {
          if (EPI == kEpiScoreGrad && p.reduce_out) tma_reduce_add_2d(&tmOut, buf, nb, row0);
          else tma_store_2d(&tmOut, buf, nb, row0);
        }
        if (lane == 0) bulk_commit();
      }
      // all TMEM reads of this accumulator are done: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (dbg && tcount == 0 && warp == 2 && lane == 0) dbg[4] = gtime();  // first epilogue drained TMEM
    }
    if (lane == 0) bulk_wait_all();
    if (dbg && warp == 2 && lane == 0) { dbg[5] = gtime(); dbg[6] = tcount; }  // all stores done
  } else if (XFORM) {
    // ------------------------------------------------------------------ mask transform (128 threads)
    const int tid = threadIdx.x - 192;
    const float thr = __ldg(p.thr);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const TileCoord tc = tile_coord(p, tile, BN);
      for (int kb = 0; kb < tc.num_kb; ++kb, ++it) {
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
        __syncwarp();
        if (lane == 0) mbar_arrive(&xform_bar[s]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile.  Each CTA
// stages its own 128 rows of A and HALF of B (128 of the 256 columns) per k-block -- 32 KB instead of
// 48 KB per 512 MMA cycles, which is what the L2 -> SM path sustains (measured ~62 B/cycle/SM: the
// 1-CTA 128 x 256 kernel above needs 94 and runs its mainloop at 67 %).  The leader CTA's single thread
// issues tcgen05.mma.cta_group::2 (M = 256); each CTA's TMEM holds its own 128 accumulator rows and its
// own eight epilogue warps (two per TMEM lane quadrant, 128 columns each) drain them.  Barriers: both producers signal the LEADER's full barrier (2SM TMA
// with the peer bit cleared); the MMA commit multicasts to both CTAs' empty / tmem-full barriers; both
// CTAs' epilogue warps arrive remotely on the leader's tmem-empty barrier.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                   "r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

// XFORM = the in-kernel mask transform (north-star kernel 1) at 2-CTA speed.  The scores are binarised ONCE per call
// into a bit mask (binarize_bits_kernel: 1 bit per weight, [rows][cols / 8] bytes -- 32 x less than the fp32 score tile
// the 1-CTA transform stages, which made that kernel shared-memory / L2-ingest bound at ~290 TFLOP/s).  Per k-block
// each CTA's producer TMA-loads its half of the W tile AND the matching 1-2 KB of mask bits onto a CTA-local barrier;
// four transform warps (one thread per 128-byte swizzled W row) AND the bf16 lanes with the expanded bits in place,
// fence towards the async proxy and arrive (cluster scope) on the leader's full barrier, which the MMA thread waits on.
template <bool XFORM>
struct Smem2T {
  static constexpr int kA = BM * BK * 2;        // 16 KB: this CTA's 128 rows of A
  static constexpr int kB = 128 * BK * 2;       // 16 KB: this CTA's half of the 256 B columns
  static constexpr int kBits = XFORM ? 1024 : 0;  // mask bits of the half tile: [128 rows][8 B] or [64 rows][16 B]
  static constexpr int kStage = kA + kB + kBits;
  static constexpr int kStages = XFORM ? 5 : 6;
  static constexpr int kEpiWarps = 8;            // two per TMEM lane quadrant: each drains 32 rows x 128 columns
  static constexpr int kXformWarps = XFORM ? 4 : 0;
  static constexpr int kThreads = 64 + 32 * (kEpiWarps + kXformWarps);
  static constexpr int kEpi = kEpiWarps * 4096;  // one 4 KB staging buffer per epilogue warp
  static constexpr int kTotal = kStages * kStage + kEpi + 256 + 1024;
};
using Smem2 = Smem2T<false>;

// expand two mask bits into the lane mask of a packed bf16 pair
__device__ __forceinline__ uint32_t bits2_to_lanes(uint32_t t) {
  return ((t & 1u) ? 0x0000FFFFu : 0u) | ((t & 2u) ? 0xFFFF0000u : 0u);
}

template <bool A_MN, bool B_MN, int EPI, bool OUT_BF16, bool XFORM = false>
__global__ void __launch_bounds__(Smem2T<XFORM>::kThreads, 1)
masked_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmOut, const GemmParams p) {
  using L = Smem2T<XFORM>;
  constexpr int STAGES = L::kStages;
  constexpr int BN = 256;
  constexpr uint32_t kBitsBox = 1024;   // mask bits of one half tile (128 x 64 weights), one contiguous bulk copy
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_base = smem + STAGES * L::kStage;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_base + L::kEpi);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint64_t* bfull_bar = tmem_empty_bar + 2;       // [STAGES] XFORM: this CTA's W half + mask bits have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfull_bar + STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = p.num_m * p.num_n * p.splits;   // num_m counts 256-row pair tiles here

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < STAGES; ++s) {
      // XFORM: + one arrival per transform warp of both CTAs (the operands are complete only once masked)
      mbar_init(&full_bar[s], 2 + 2 * L::kXformWarps);
      mbar_init(&empty_bar[s], 1);
      if (XFORM) mbar_init(&bfull_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 2 * L::kEpiWarps);   // every epilogue warp of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // barriers, TMEM and descriptors are set up; operands of a predecessor from here on

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int it = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
        TileCoord tc = tile_coord(p, tile, BN);
        const int m0 = tc.m0 * 2 + rank * BM;   // tile_coord used BM = 128 per m index; pair tiles are 256 rows
        const int nh = tc.n0 + rank * 128;      // this CTA's half of the B columns
        for (int kb = 0; kb < tc.num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          // XFORM: only A goes to the leader's barrier; B and its mask bits land on this CTA's own barrier
          if (leader) mbar_expect_tx(&full_bar[s], XFORM ? 2 * L::kA : 2 * L::kStage);
          else mbar_arrive_remote(&full_bar[s], 0);
          uint8_t* sA = smem + s * L::kStage;
          uint8_t* sB = sA + L::kA;
          const int kk0 = (tc.kb_begin + kb) * BK;
          if (!A_MN) {
            tma_load_2d_2sm(sA, &tmA, &full_bar[s], kk0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_2sm(sA + j * 8192, &tmA, &full_bar[s], m0 + 64 * j, kk0);
          }
          if (XFORM) {
            mbar_expect_tx(&bfull_bar[s], L::kB + kBitsBox);
            uint8_t* sBits = sB + L::kB;
            const int kbi = tc.kb_begin + kb;
            if (!B_MN) {     // W rows nh .. nh+127, W columns kk0 .. kk0+63: bit tile (nh / 128, kbi)
              tma_load_2d(sB, &tmB, &bfull_bar[s], kk0, nh);
              bulk_load_1d(sBits, p.bits + (static_cast<size_t>(nh >> 7) * p.bits_pitch + kbi) * 1024, 1024, &bfull_bar[s]);
            } else {         // W rows kk0 .. kk0+63, W columns nh .. nh+127: bit tile (kbi, nh / 128)
#pragma unroll
              for (int j = 0; j < 2; ++j) tma_load_2d(sB + j * 8192, &tmB, &bfull_bar[s], nh + 64 * j, kk0);
              bulk_load_1d(sBits, p.bits + (static_cast<size_t>(kbi) * p.bits_pitch + (nh >> 7)) * 1024, 1024, &bfull_bar[s]);
            }
          } else if (!B_MN) {
            tma_load_2d_2sm(sB, &tmB, &full_bar[s], kk0, nh);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_2sm(sB + j * 8192, &tmB, &full_bar[s], nh + 64 * j, kk0);
          }
        }
      }
      pdl_launch_dependents();   // every load of this CTA is in flight: let the next grid be scheduled
    }
  } else if (XFORM && warp >= 2 + L::kEpiWarps) {
    // ------------------------------------------------------------------ mask transform (both CTAs, 128 threads)
    // thread r owns one 128-byte (64 bf16) swizzled row of this CTA's W half; its 64 mask bits sit in the bits box
    const int r = threadIdx.x - 32 * (2 + L::kEpiWarps);
    const int sw = r & 7;
    const uint32_t row_off = B_MN ? (r >> 6) * 8192 + (r & 63) * 128 : r * 128;
    const uint32_t bits_off = B_MN ? (r & 63) * 16 + (r >> 6) * 8 : r * 8;
    int it = 0;
    for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
      const TileCoord tc = tile_coord(p, tile, BN);
      for (int kb = 0; kb < tc.num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        uint8_t* sB = smem + s * L::kStage + L::kA;
        mbar_wait(&bfull_bar[s], ph);
        const uint2 bw = *reinterpret_cast<const uint2*>(sB + L::kB + bits_off);
        uint4* row = reinterpret_cast<uint4*>(sB + row_off);
        uint4 w[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) w[c] = row[c ^ sw];      // logical 16-byte chunk c (elements 8c .. 8c+7)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t b = ((c < 4 ? bw.x : bw.y) >> (8 * (c & 3))) & 0xFFu;
          w[c].x &= bits2_to_lanes(b);
          w[c].y &= bits2_to_lanes(b >> 2);
          w[c].z &= bits2_to_lanes(b >> 4);
          w[c].w &= bits2_to_lanes(b >> 6);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) row[c ^ sw] = w[c];
        // generic-proxy writes -> async proxy (the tensor core reads this tile), then tell the leader's MMA thread; same
        // arrive as the epilogue's tmem-empty hand-off (a cluster-scope release compiles to MEMBAR.GPU per warp per
        // stage and measured no different in results)
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(&full_bar[s], 0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int it = 0, tcount = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++tcount) {
        const TileCoord tc = tile_coord(p, tile, BN);
        const int acc = tcount & 1;
        mbar_wait(&tmem_empty_bar[acc], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < tc.num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
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
              umma_bf16_2sm(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[s]);
            if (kb == tc.num_kb - 1) umma_commit_2sm(&tmem_full_bar[acc]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;       // which 128 of the tile's 256 columns
    uint8_t* buf = epi_base + (warp - 2) * 4096;
    constexpr int COLS = OUT_BF16 ? 64 : 32;
    constexpr int HALF_N = BN / 2;
    int tcount = 0;
    for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++tcount) {
      const TileCoord tc = tile_coord(p, tile, BN);
      const int acc = tcount & 1;
      const int row0 = tc.m0 * 2 + rank * BM + q * 32;
      const int m = row0 + lane;
      mbar_wait(&tmem_full_bar[acc], (tcount >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < HALF_N / COLS; ++c) {
        const int col0 = half * HALF_N + c * COLS;
        const int nb = tc.n0 + col0;
        uint32_t r[32];
        uint32_t pk[32];
        float bias_lo = 0.f, bias_hi = 0.f;
        if (EPI == kEpiStore && p.bias != nullptr) {
          if (nb + lane < p.NN) bias_lo = __ldg(p.bias + nb + lane);
          if (OUT_BF16 && nb + 32 + lane < p.NN) bias_hi = __ldg(p.bias + nb + 32 + lane);
        }
        tmem_ld_32x32(t_addr + col0, r);
        if (OUT_BF16) {
          uint32_t r2[32];
          tmem_ld_32x32(t_addr + col0 + 32, r2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a0 = __uint_as_float(r[2 * j]) + __shfl_sync(0xffffffffu, bias_lo, 2 * j);
            const float a1 = __uint_as_float(r[2 * j + 1]) + __shfl_sync(0xffffffffu, bias_lo, 2 * j + 1);
            const float b0 = __uint_as_float(r2[2 * j]) + __shfl_sync(0xffffffffu, bias_hi, 2 * j);
            const float b1 = __uint_as_float(r2[2 * j + 1]) + __shfl_sync(0xffffffffu, bias_hi, 2 * j + 1);
            __nv_bfloat162 ta = __floats2bfloat162_rn(a0, a1);
            __nv_bfloat162 tb = __floats2bfloat162_rn(b0, b1);
            pk[j] = *reinterpret_cast<uint32_t*>(&ta);
            pk[16 + j] = *reinterpret_cast<uint32_t*>(&tb);
          }
        } else if (EPI == kEpiStore) {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            pk[j] = __float_as_uint(__uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bias_lo, j));
        } else {
          const bool row_ok = m < p.MM;
          const float* wrow = p.w + static_cast<size_t>(row_ok ? m : 0) * p.NN + nb;
          const bool fast = row_ok && nb + 32 <= p.NN && (p.NN & 3) == 0;
          float4 wq[8];
          if (fast) {      // the fp32 multiplier row is fetched while the TMEM load is in flight, not after it
#pragma unroll
            for (int j = 0; j < 8; ++j) wq[j] = __ldg(reinterpret_cast<const float4*>(wrow) + j);
          }
          tmem_ld_wait();
          if (fast) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              pk[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) * wq[j].x);
              pk[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) * wq[j].y);
              pk[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) * wq[j].z);
              pk[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) * wq[j].w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float wv = (row_ok && nb + j < p.NN) ? __ldg(wrow + j) : 0.f;
              pk[j] = __float_as_uint(__uint_as_float(r[j]) * wv);
            }
          }
        }
        if (lane == 0) bulk_wait_read<0>();     // the previous chunk's TMA store has read the staging buffer
        __syncwarp();
        const int sw = lane & 7;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(buf + lane * 128 + ((j ^ sw) << 4)) =
              make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && nb < p.NN && row0 < p.MM) {
          if (EPI == kEpiScoreGrad && p.reduce_out) tma_reduce_add_2d(&tmOut, buf, nb, row0);
          else tma_store_2d(&tmOut, buf, nb, row0);
        }
        if (lane == 0) bulk_commit();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tmem_empty_bar[acc], 0);
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();   // the peer may still be signalling this CTA's barriers / reading its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 2 * BN);
  }
}


// ------------------------------------------------------------------------------------------------
// Grouped 2-CTA kernel: up to kMaxGroup independent GEMM problems of a layer phase in ONE persistent launch
// (dX and dS of a module -- both read the same dY; the two sides of a cross-modality layer; the language and the
// vision stack in lockstep).  The stage-2 step is 369 GEMMs of which a third are one-wave problems (N = K = 768:
// 60 - 108 tile pairs on 74 CTA pairs) whose pipeline fill, epilogue and drain cost more than their mainloop; in a
// group the tile loop runs across problem boundaries, so the epilogue of one problem's last tile overlaps the
// mainloop of the next problem's first tile, wave quantisation applies to the SUM of the tiles, and launch + set-up
// (TMEM allocation, barrier init, cluster sync, descriptor prefetch) is paid once per group.
//
// Same pipeline as masked_gemm2_kernel (2-SM TMA into 128B-swizzled tiles, tcgen05.mma.cta_group::2 from one
// thread, double-buffered 256-column TMEM accumulators, eight epilogue warps per CTA) with the per-problem
// properties (operand majors, epilogue kind, split of the reduction) read at run time from the problem table in
// kernel parameter space.  Two more epilogues fuse the GELU of the FFN into the GEMMs around it:
//   kGEpiGelu      u = D + bias -> bf16 ; out = bf16(gelu(u)) and aux = u (both TMA-stored): FF1 forward
//   kGEpiGeluGrad  out = bf16(D * gelu'(u)), u read from aux_in: the dX of FF2 becomes dU directly
// Staging is double-buffered per epilogue warp (2 x 4 KB), which is what the two-output epilogue needs and lets a
// chunk be staged while the previous chunk's TMA store still reads its buffer; five 32 KB operand stages.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxGroup = 4;
constexpr int kGEpiF32 = 0, kGEpiBf16 = 1, kGEpiScoreGrad = 2, kGEpiGelu = 3, kGEpiGeluGrad = 4;

struct alignas(64) GProblem {
  CUtensorMap tmA, tmB, tmOut, tmAux;
  const float* bias;         // [NN] or null (store epilogues)
  const float* w;            // [MM, NN] fp32 multiplier (score-grad epilogue)
  const uint16_t* aux_in;    // [MM, NN] bf16 pre-activation u (gelu-grad epilogue)
  int MM, NN, KK;
  int kb_per_split, splits, num_m, num_n;   // num_m counts 256-row pair tiles
  int tile_end;              // exclusive end of this problem's tiles in the group's tile numbering
  int a_mn, b_mn, epi, reduce_out;
};
struct GArgs {
  GProblem p[kMaxGroup];
  int count, total_tiles;
};

template <int STAGES_, int NBUF_>
struct SmemG {
  static constexpr int kStage = 32768;           // A: this CTA's 128 rows; B: this CTA's half of the 256 columns
  static constexpr int kStages = STAGES_;        // 6 stages + single staging, or 5 stages + double staging (227 KB)
  static constexpr int kBufs = NBUF_;
  static constexpr int kEpiWarps = 8;
  static constexpr int kEpi = kEpiWarps * NBUF_ * 4096;
  static constexpr int kTotal = kStages * kStage + kEpi + 256 + 1024;
};

struct GTile {
  int pi, m0, n0, kb_begin, num_kb;
};

__device__ __forceinline__ GTile gtile(const GArgs& a, int tile) {
  int pi = 0, begin = 0;
#pragma unroll
  for (int i = 0; i < kMaxGroup - 1; ++i) {
    if (pi == i && tile >= a.p[i].tile_end) {
      begin = a.p[i].tile_end;
      pi = i + 1;
    }
  }
  const GProblem& p = a.p[pi];
  const int local = tile - begin;
  const int per_z = p.num_m * p.num_n;
  const int z = local / per_z;
  const int r = local - z * per_z;
  const int mt = r / p.num_n;
  const int nt = r - mt * p.num_n;
  const int total_kb = (p.KK + BK - 1) / BK;
  GTile t;
  t.pi = pi;
  t.m0 = mt * 256;
  t.n0 = nt * 256;
  t.kb_begin = z * p.kb_per_split;
  int e = t.kb_begin + p.kb_per_split;
  if (e > total_kb) e = total_kb;
  t.num_kb = e - t.kb_begin;
  return t;
}

__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// stage 32 rows x 128 bytes (one register row per lane) into a swizzled buffer and hand it to TMA
template <int NBUF>
__device__ __forceinline__ void stage_and_store(uint8_t* bufs, int& nstore, const uint32_t (&pk)[32], int lane,
                                                const CUtensorMap* map, int c0, int r0, bool in_range, bool reduce) {
  uint8_t* buf = bufs + (NBUF == 2 ? (nstore & 1) * 4096 : 0);
  ++nstore;
  if (lane == 0) bulk_wait_read<NBUF - 1>();   // the previous store that used this buffer has read its data
  __syncwarp();
  const int sw = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<uint4*>(buf + lane * 128 + ((j ^ sw) << 4)) =
        make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    if (in_range) {
      if (reduce) tma_reduce_add_2d(map, buf, c0, r0);
      else tma_store_2d(map, buf, c0, r0);
    }
    bulk_commit();
  }
}


// One epilogue warp's share of a tile: 32 accumulator rows (lane = row) x 128 columns, in chunks of one 128-byte
// staging row (32 fp32 or 64 bf16 columns).  The grouped kernel only takes problems whose output width is a multiple
// of 256, so every chunk has all its columns; rows past MM exist only in the last row tile: their loads are clamped
// to row MM - 1 and their results never leave the SM (the TMA store clips them).  No per-element guards: the kernel
// has to stay small -- past ~40 KB of code the single MMA-issuing thread starts to wait on instruction fetches.
template <int EPI, int NBUF>
__device__ __forceinline__ void epilogue_tile(const GProblem& p, int n0, int row0, uint32_t t_addr, int half, int lane,
                                              uint8_t* bufs, int& nstore) {
  const int MM = p.MM, NN = p.NN;
  const int m_ld = min(row0 + lane, MM - 1);
  const bool in_range = row0 < MM;
  const CUtensorMap* tm_out = &p.tmOut;
  if (EPI == kGEpiF32 || EPI == kGEpiScoreGrad) {
    const float* bias = p.bias;
    const float* wrow0 = EPI == kGEpiScoreGrad ? p.w + static_cast<size_t>(m_ld) * NN : nullptr;
    const bool reduce = EPI == kGEpiScoreGrad && p.reduce_out;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int col0 = half * 128 + c * 32;
      const int nb = n0 + col0;
      uint32_t r[32];
      float bias_v = 0.f;
      float4 wq[8];
      if (EPI == kGEpiF32) {
        if (bias != nullptr) bias_v = __ldg(bias + nb + lane);
      } else {                 // the fp32 multiplier row is fetched while the TMEM load is in flight
#pragma unroll
        for (int j = 0; j < 8; ++j) wq[j] = __ldg(reinterpret_cast<const float4*>(wrow0 + nb) + j);
      }
      tmem_ld_32x32(t_addr + col0, r);
      tmem_ld_wait();
      if (EPI == kGEpiF32) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          r[j] = __float_as_uint(__uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bias_v, j));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) * wq[j].x);
          r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) * wq[j].y);
          r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) * wq[j].z);
          r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) * wq[j].w);
        }
      }
      stage_and_store<NBUF>(bufs, nstore, r, lane, tm_out, nb, row0, in_range, reduce);
    }
  } else {
    const uint16_t* urow0 = EPI == kGEpiGeluGrad ? p.aux_in + static_cast<size_t>(m_ld) * NN : nullptr;
    const float* bias = p.bias;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int col0 = half * 128 + c * 64;
      const int nb = n0 + col0;
      uint32_t r[32], r2[32], pk[32];
      if (EPI == kGEpiGeluGrad) {
        uint4 uq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) uq[j] = __ldg(reinterpret_cast<const uint4*>(urow0 + nb) + j);
        tmem_ld_32x32(t_addr + col0, r);
        tmem_ld_32x32(t_addr + col0 + 32, r2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {        // uq[j]: u of columns 8j .. 8j+7 ; uq[4 + j]: columns 32 + 8j ..
          const uint32_t ua[4] = {uq[j].x, uq[j].y, uq[j].z, uq[j].w};
          const uint32_t ub[4] = {uq[4 + j].x, uq[4 + j].y, uq[4 + j].z, uq[4 + j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            pk[4 * j + e] = pack2_bf16(__uint_as_float(r[8 * j + 2 * e]) * gelu_grad_f(bf16_lo(ua[e])),
                                       __uint_as_float(r[8 * j + 2 * e + 1]) * gelu_grad_f(bf16_hi(ua[e])));
            pk[16 + 4 * j + e] = pack2_bf16(__uint_as_float(r2[8 * j + 2 * e]) * gelu_grad_f(bf16_lo(ub[e])),
                                            __uint_as_float(r2[8 * j + 2 * e + 1]) * gelu_grad_f(bf16_hi(ub[e])));
          }
        }
        stage_and_store<NBUF>(bufs, nstore, pk, lane, tm_out, nb, row0, in_range, false);
      } else {
        float bias_lo = 0.f, bias_hi = 0.f;
        if (bias != nullptr) {
          bias_lo = __ldg(bias + nb + lane);
          bias_hi = __ldg(bias + nb + 32 + lane);
        }
        tmem_ld_32x32(t_addr + col0, r);
        tmem_ld_32x32(t_addr + col0 + 32, r2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          pk[j] = pack2_bf16(__uint_as_float(r[2 * j]) + __shfl_sync(0xffffffffu, bias_lo, 2 * j),
                             __uint_as_float(r[2 * j + 1]) + __shfl_sync(0xffffffffu, bias_lo, 2 * j + 1));
          pk[16 + j] = pack2_bf16(__uint_as_float(r2[2 * j]) + __shfl_sync(0xffffffffu, bias_hi, 2 * j),
                                  __uint_as_float(r2[2 * j + 1]) + __shfl_sync(0xffffffffu, bias_hi, 2 * j + 1));
        }
        if (EPI == kGEpiGelu) {
          // the pre-activation goes out as it is (bf16), then the same registers become gelu(u) -- computed from
          // the ROUNDED u, which is what the backward multiplies gelu'() of
          stage_and_store<NBUF>(bufs, nstore, pk, lane, &p.tmAux, nb, row0, in_range, false);
#pragma unroll
          for (int j = 0; j < 32; ++j) pk[j] = pack2_bf16(gelu_f(bf16_lo(pk[j])), gelu_f(bf16_hi(pk[j])));
        }
        stage_and_store<NBUF>(bufs, nstore, pk, lane, tm_out, nb, row0, in_range, false);
      }
    }
  }
}

// EPIMASK: bit k set = the group may contain epilogue kind k (only those loops are compiled in)
template <int STAGES_, int NBUF, int EPIMASK>
__global__ void __launch_bounds__(64 + 32 * 8, 1)
grouped_gemm2_kernel(const __grid_constant__ GArgs args) {
  using L = SmemG<STAGES_, NBUF>;
  constexpr int STAGES = L::kStages;
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_base = smem + STAGES * L::kStage;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_base + L::kEpi);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = args.total_tiles;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < args.count; ++i) {
      tma_prefetch_desc(&args.p[i].tmA);
      tma_prefetch_desc(&args.p[i].tmB);
      tma_prefetch_desc(&args.p[i].tmOut);
      if (args.p[i].epi == kGEpiGelu) tma_prefetch_desc(&args.p[i].tmAux);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 2 * L::kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int it = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
        const GTile tc = gtile(args, tile);
        const GProblem& p = args.p[tc.pi];
        const int m0 = tc.m0 + rank * BM;
        const int nh = tc.n0 + rank * 128;
        const bool a_mn = p.a_mn != 0, b_mn = p.b_mn != 0;
        for (int kb = 0; kb < tc.num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (leader) mbar_expect_tx(&full_bar[s], 2 * L::kStage);
          else mbar_arrive_remote(&full_bar[s], 0);
          uint8_t* sA = smem + s * L::kStage;
          uint8_t* sB = sA + 16384;
          const int kk0 = (tc.kb_begin + kb) * BK;
          if (!a_mn) {
            tma_load_2d_2sm(sA, &p.tmA, &full_bar[s], kk0, m0);
          } else {
            tma_load_2d_2sm(sA, &p.tmA, &full_bar[s], m0, kk0);
            tma_load_2d_2sm(sA + 8192, &p.tmA, &full_bar[s], m0 + 64, kk0);
          }
          if (!b_mn) {
            tma_load_2d_2sm(sB, &p.tmB, &full_bar[s], kk0, nh);
          } else {
            tma_load_2d_2sm(sB, &p.tmB, &full_bar[s], nh, kk0);
            tma_load_2d_2sm(sB + 8192, &p.tmB, &full_bar[s], nh + 64, kk0);
          }
        }
      }
      pdl_launch_dependents();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (leader) {
      int it = 0, tcount = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++tcount) {
        const GTile tc = gtile(args, tile);
        const GProblem& p = args.p[tc.pi];
        const bool a_mn = p.a_mn != 0, b_mn = p.b_mn != 0;
        const uint32_t idesc = make_idesc_bf16(256, BN, a_mn ? 1 : 0, b_mn ? 1 : 0);
        const int acc = tcount & 1;
        mbar_wait(&tmem_empty_bar[acc], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < tc.num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t aBase = smem_u32(smem + s * L::kStage);
            const uint32_t bBase = aBase + 16384;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = a_mn ? make_sw128_desc(aBase + k * 2048, 8192, 1024)
                                       : make_sw128_desc(aBase + k * 32, 16, 1024);
              const uint64_t db = b_mn ? make_sw128_desc(bBase + k * 2048, 8192, 1024)
                                       : make_sw128_desc(bBase + k * 32, 16, 1024);
              umma_bf16_2sm(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[s]);
            if (kb == tc.num_kb - 1) umma_commit_2sm(&tmem_full_bar[acc]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;       // which 128 of the tile's 256 columns
    uint8_t* bufs = epi_base + (warp - 2) * (NBUF * 4096);
    int tcount = 0, nstore = 0;
    for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++tcount) {
      const GTile tc = gtile(args, tile);
      const GProblem& p = args.p[tc.pi];
      const int epi = p.epi;
      const int acc = tcount & 1;
      const int row0 = tc.m0 + rank * BM + q * 32;
      mbar_wait(&tmem_full_bar[acc], (tcount >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
      // one specialised loop per epilogue kind the instantiation admits
      if (((EPIMASK >> kGEpiF32) & 1) && epi == kGEpiF32)
        epilogue_tile<kGEpiF32, NBUF>(p, tc.n0, row0, t_addr, half, lane, bufs, nstore);
      if (((EPIMASK >> kGEpiBf16) & 1) && epi == kGEpiBf16)
        epilogue_tile<kGEpiBf16, NBUF>(p, tc.n0, row0, t_addr, half, lane, bufs, nstore);
      if (((EPIMASK >> kGEpiScoreGrad) & 1) && epi == kGEpiScoreGrad)
        epilogue_tile<kGEpiScoreGrad, NBUF>(p, tc.n0, row0, t_addr, half, lane, bufs, nstore);
      if (((EPIMASK >> kGEpiGelu) & 1) && epi == kGEpiGelu)
        epilogue_tile<kGEpiGelu, NBUF>(p, tc.n0, row0, t_addr, half, lane, bufs, nstore);
      if (((EPIMASK >> kGEpiGeluGrad) & 1) && epi == kGEpiGeluGrad)
        epilogue_tile<kGEpiGeluGrad, NBUF>(p, tc.n0, row0, t_addr, half, lane, bufs, nstore);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tmem_empty_bar[acc], 0);
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 2 * BN);
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

// SMs the persistent GEMM grids may occupy.  Default: all.  CRVQA_GEMM_SMS=n (even) leaves the rest to whatever runs
// beside the GEMMs -- in data-parallel runs the NCCL kernels of the gradient exchange, whose CTAs cannot share an SM
// with a 227 KB GEMM CTA and would otherwise delay the tiles statically assigned to the SMs they grab.
static int gemm_sms() {
  static const int n = [] {
    const char* e = getenv("CRVQA_GEMM_SMS");
    int v = e ? atoi(e) : 0;
    const int all = num_sms();
    if (v <= 0 || v > all) v = all;
    return v & ~1;
  }();
  return n;
}

// 2-D row-major tensor [rows][cols] of elem_bytes elements; box = [box_rows][box_cols], 128B swizzle.
static int make_map(CUtensorMap* map, const void* base, int elem_bytes, bool is_float, int64_t rows, int64_t cols,
                    int box_rows, int box_cols) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return CRV_E_DRIVER;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = is_float ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = enc(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_cuda_error = static_cast<int>(r);
    return static_cast<int>(r);
  }
  return CRV_OK;
}

static int make_out_map(CUtensorMap* map, void* out, bool bf16, int64_t rows, int64_t cols) {
  return bf16 ? make_map(map, out, 2, false, rows, cols, 32, 64) : make_map(map, out, 4, true, rows, cols, 32, 32);
}

static long long* g_dbg = nullptr;

template <bool A_MN, bool B_MN, int BN, bool XFORM, int EPI, bool OUT_BF16>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmS, const CUtensorMap& tmOut,
                  GemmParams p, cudaStream_t stream) {
  using L = SmemLayout<BN, XFORM>;
  auto kern = masked_gemm_kernel<A_MN, B_MN, BN, XFORM, EPI, OUT_BF16>;
  static bool configured = false;
  if (!configured) {
    CRV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  p.num_m = (p.MM + BM - 1) / BM;
  p.num_n = (p.NN + BN - 1) / BN;
  p.dbg = g_dbg;
  const int tiles = p.num_m * p.num_n * p.splits;
  const int grid = tiles < gemm_sms() ? tiles : gemm_sms();
  CRV_CUDA(launch_pdl(kern, dim3(grid), dim3(XFORM ? 320 : 192), L::kTotal, stream, tmA, tmB, tmS, tmOut, p));
  return launch_status();
}


static bool use_2cta(int MM, int NN) {
  static const int mode = [] { const char* e = getenv("CRVQA_2CTA"); return e ? atoi(e) : 1; }();
  return mode != 0 && NN % 256 == 0 && MM >= 256;
}

// ---- mask bits of the 2-CTA in-kernel transform ------------------------------------------------------------------
// bit j of a byte = S[row][8 c + j] > thr, stored TILE-CONTIGUOUS so that the bits of one half tile (128 x 64 weights =
// 1 KB) arrive with one bulk copy:
//   layout 0 (forward, tile = 128 W rows x 64 W cols):  tile (row / 128, col / 64),  byte (row % 128) * 8  + (col % 64) / 8
//   layout 1 (dX,      tile = 64 W rows x 128 W cols):  tile (row / 64,  col / 128), byte (row % 64)  * 16 + (col % 128) / 8
// Rows / columns past the matrix get zero bits (the W tile is zero-filled there anyway).
__global__ void binarize_bits_tiled_kernel(const float* __restrict__ s, const float* __restrict__ thr,
                                           uint8_t* __restrict__ bits, int rows, int cols, int layout, int pitch,
                                           int64_t nbytes) {
  pdl_wait();
  const float t = __ldg(thr);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nbytes; i += stride) {
    const int64_t tile = i >> 10;
    const int in = static_cast<int>(i & 1023);
    const int tr = static_cast<int>(tile / pitch), tcn = static_cast<int>(tile % pitch);
    int row, col;
    if (layout == 0) { row = tr * 128 + (in >> 3); col = tcn * 64 + (in & 7) * 8; }
    else { row = tr * 64 + (in >> 4); col = tcn * 128 + (in & 15) * 8; }
    uint32_t v = 0;
    if (row < rows && col < cols) {       // cols % 8 == 0: the eight scores of a byte are all inside
      const float4* src = reinterpret_cast<const float4*>(s + static_cast<size_t>(row) * cols + col);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      v = (a.x > t ? 1u : 0u) | (a.y > t ? 2u : 0u) | (a.z > t ? 4u : 0u) | (a.w > t ? 8u : 0u) |
          (b.x > t ? 16u : 0u) | (b.y > t ? 32u : 0u) | (b.z > t ? 64u : 0u) | (b.w > t ? 128u : 0u);
    }
    bits[i] = static_cast<uint8_t>(v);
  }
}

// Scratch for the mask bits, one buffer per (device, stream): calls on one stream are ordered, so the next call's
// pre-pass cannot overwrite bits a running GEMM still reads; another stream gets its own buffer.  Grown (never
// shrunk) outside the data path; the allocation runs in relaxed capture mode so a first use during CUDA-graph
// capture is legal (the buffer outlives the graph: it is never freed while the library is loaded).
static uint8_t* xform_bits_scratch(cudaStream_t stream, size_t bytes) {
  struct Buf { uint8_t* p; size_t n; };
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, Buf> bufs;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  Buf& b = bufs[{dev, stream}];
  if (b.n < bytes) {
    size_t want = bytes < (size_t(4) << 20) ? (size_t(4) << 20) : bytes;   // 4 MB covers every LXMERT / VisualBERT module
    cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
    cudaThreadExchangeStreamCaptureMode(&mode);
    uint8_t* np = nullptr;
    const cudaError_t e = cudaMalloc(&np, want);
    cudaThreadExchangeStreamCaptureMode(&mode);
    if (e != cudaSuccess) { g_last_cuda_error = static_cast<int>(e); return nullptr; }
    // the old buffer (if any) may still be read by a kernel in flight: it is left allocated (a few MB, once)
    b.p = np;
    b.n = want;
  }
  return b.p;
}

// the 2-CTA transform needs whole 256-wide tiles; forward k-blocks must be whole too (w_cols % 64 == 0)
static bool use_2cta_xform(int MM, int NN, int w_cols) {
  static const int mode = [] { const char* e = getenv("CRVQA_XFORM_2CTA"); return e ? atoi(e) : 1; }();
  return mode != 0 && use_2cta(MM, NN) && w_cols % 64 == 0;
}

// binarise `scores` [rows][cols] against *thr into the stream's scratch in tile layout `layout`; fills p.bits / p.bits_pitch
static int prepare_mask_bits(const float* scores, const float* thr, int rows, int cols, int layout, GemmParams* p,
                             cudaStream_t st) {
  const int tile_rows = layout == 0 ? 128 : 64, tile_cols = layout == 0 ? 64 : 128;
  const int tr = (rows + tile_rows - 1) / tile_rows, tcn = (cols + tile_cols - 1) / tile_cols;
  const int64_t nbytes = static_cast<int64_t>(tr) * tcn * 1024;
  uint8_t* bits = xform_bits_scratch(st, static_cast<size_t>(nbytes));
  if (!bits) return CRV_E_DRIVER;
  int64_t blocks = (nbytes + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  CRV_CUDA(launch_pdl(binarize_bits_tiled_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, scores, thr,
                      bits, rows, cols, layout, tcn, nbytes));
  p->bits = bits;
  p->bits_pitch = tcn;
  return launch_status();
}

template <bool A_MN, bool B_MN, int EPI, bool OUT_BF16, bool XFORM = false>
static int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, GemmParams p,
                   cudaStream_t stream) {
  using L2 = Smem2T<XFORM>;
  auto kern = masked_gemm2_kernel<A_MN, B_MN, EPI, OUT_BF16, XFORM>;
  static bool configured = false;
  if (!configured) {
    CRV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L2::kTotal));
    configured = true;
  }
  p.num_m = (p.MM + 255) / 256;
  p.num_n = (p.NN + 255) / 256;
  p.dbg = nullptr;
  const int tiles = p.num_m * p.num_n * p.splits;
  const int max_pairs = gemm_sms() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(L2::kThreads);
  cfg.dynamicSmemBytes = L2::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  CRV_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, p));
  return launch_status();
}

static int pick_bn(int MM, int NN) {
  // wave quantisation on a persistent grid: cost ~ ceil(tiles / SMs) * BN; 256-wide tiles halve the A
  // re-reads and win ties
  if (NN % 256) return 128;
  static const int forced = [] { const char* e = getenv("CRVQA_BN"); return e ? atoi(e) : 0; }();
  if (forced == 128 || forced == 256) return forced;
  const int sms = gemm_sms();
  const int mt = (MM + BM - 1) / BM;
  const int t256 = mt * (NN / 256), t128 = mt * ((NN + 127) / 128);
  const int c256 = ((t256 + sms - 1) / sms) * 256, c128 = ((t128 + sms - 1) / sms) * 128;
  return c256 <= c128 ? 256 : 128;
}

}  // namespace crv

using namespace crv;

// Debug aid: when set, every GEMM CTA writes 8 x int64 globaltimer stamps {setup, first TMA, first stage
// landed, first accumulator ready, first epilogue drained, all stores done, tiles, -} at buf[8 * blockIdx.x].
extern "C" int crv_gemm_debug_timestamps(long long* buf) {
  g_dbg = buf;
  return CRV_OK;
}

extern "C" int crv_masked_linear_fwd(const uint16_t* x, const uint16_t* w, const float* scores, const float* thr,
                                     const float* bias, void* y, int y_dtype, int M, int N, int K, void* stream) {
  if (!x || !w || !y || M <= 0 || N <= 0 || K <= 0) return CRV_E_BADARG;
  if (scores && !thr) return CRV_E_BADARG;
  if (y_dtype != CRV_DTYPE_F32 && y_dtype != CRV_DTYPE_BF16) return CRV_E_BADARG;
  const bool obf = y_dtype == CRV_DTYPE_BF16;
  if ((K % 8) || (N % (obf ? 8 : 4))) return CRV_E_SHAPE;
  if (!aligned16(x) || !aligned16(w) || !aligned16(y) || (scores && !aligned16(scores))) return CRV_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GemmParams p{};
  p.MM = M; p.NN = N; p.KK = K;
  p.kb_per_split = (K + BK - 1) / BK;
  p.splits = 1;
  p.thr = thr; p.bias = bias;
  CUtensorMap tmA, tmB, tmS, tmO;
  int rc;
  if ((rc = make_map(&tmA, x, 2, false, M, K, BM, BK))) return rc;
  if ((rc = make_out_map(&tmO, y, obf, M, N))) return rc;
  if (scores && use_2cta_xform(M, N, K)) {
    // in-kernel mask at 2-CTA speed: bits pre-pass (one read of S) + transform warps between TMA and tcgen05.mma
    if ((rc = prepare_mask_bits(scores, thr, N, K, 0, &p, st))) return rc;
    if ((rc = make_map(&tmB, w, 2, false, N, K, 128, BK))) return rc;
    return obf ? launch2<false, false, kEpiStore, true, true>(tmA, tmB, tmO, p, st)
               : launch2<false, false, kEpiStore, false, true>(tmA, tmB, tmO, p, st);
  }
  if (scores) {
    if ((rc = make_map(&tmB, w, 2, false, N, K, 128, BK))) return rc;
    if ((rc = make_map(&tmS, scores, 4, true, N, K, 128, 32))) return rc;
    return obf ? launch<false, false, 128, true, kEpiStore, true>(tmA, tmB, tmS, tmO, p, st)
               : launch<false, false, 128, true, kEpiStore, false>(tmA, tmB, tmS, tmO, p, st);
  }
  if (use_2cta(M, N)) {
    if ((rc = make_map(&tmB, w, 2, false, N, K, 128, BK))) return rc;
    return obf ? launch2<false, false, kEpiStore, true>(tmA, tmB, tmO, p, st)
               : launch2<false, false, kEpiStore, false>(tmA, tmB, tmO, p, st);
  }
  const int bn = pick_bn(M, N);
  if ((rc = make_map(&tmB, w, 2, false, N, K, bn, BK))) return rc;
  tmS = tmB;
  if (bn == 256)
    return obf ? launch<false, false, 256, false, kEpiStore, true>(tmA, tmB, tmS, tmO, p, st)
               : launch<false, false, 256, false, kEpiStore, false>(tmA, tmB, tmS, tmO, p, st);
  return obf ? launch<false, false, 128, false, kEpiStore, true>(tmA, tmB, tmS, tmO, p, st)
             : launch<false, false, 128, false, kEpiStore, false>(tmA, tmB, tmS, tmO, p, st);
}

extern "C" int crv_masked_linear_bwd_dx(const uint16_t* dy, const uint16_t* w, const float* scores,
                                        const float* thr, void* dx, int dx_dtype, int M, int N, int K,
                                        void* stream) {
  if (!dy || !w || !dx || M <= 0 || N <= 0 || K <= 0) return CRV_E_BADARG;
  if (scores && !thr) return CRV_E_BADARG;
  if ((N % 8) || (K % 8)) return CRV_E_SHAPE;
  if (!aligned16(dy) || !aligned16(w) || !aligned16(dx) || (scores && !aligned16(scores))) return CRV_E_ALIGN;
  if (dx_dtype != CRV_DTYPE_F32 && dx_dtype != CRV_DTYPE_BF16) return CRV_E_BADARG;
  const bool obf = dx_dtype == CRV_DTYPE_BF16;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // generic: MM = M, NN = K, KK = N;  A = dY [M][N] K-major;  B = W [N][K] = [KK][NN] MN-major
  GemmParams p{};
  p.MM = M; p.NN = K; p.KK = N;
  p.kb_per_split = (N + BK - 1) / BK;
  p.splits = 1;
  p.thr = thr;
  CUtensorMap tmA, tmB, tmS, tmO;
  int rc;
  if ((rc = make_map(&tmA, dy, 2, false, M, N, BM, BK))) return rc;
  if ((rc = make_map(&tmB, w, 2, false, N, K, 64, 64))) return rc;
  if ((rc = make_out_map(&tmO, dx, obf, M, K))) return rc;
  if (scores && use_2cta_xform(M, K, K)) {
    if ((rc = prepare_mask_bits(scores, thr, N, K, 1, &p, st))) return rc;
    return obf ? launch2<false, true, kEpiStore, true, true>(tmA, tmB, tmO, p, st)
               : launch2<false, true, kEpiStore, false, true>(tmA, tmB, tmO, p, st);
  }
  if (scores) {
    if ((rc = make_map(&tmS, scores, 4, true, N, K, 64, 32))) return rc;
    return obf ? launch<false, true, 128, true, kEpiStore, true>(tmA, tmB, tmS, tmO, p, st)
               : launch<false, true, 128, true, kEpiStore, false>(tmA, tmB, tmS, tmO, p, st);
  }
  tmS = tmB;
  if (use_2cta(M, K))
    return obf ? launch2<false, true, kEpiStore, true>(tmA, tmB, tmO, p, st)
               : launch2<false, true, kEpiStore, false>(tmA, tmB, tmO, p, st);
  if (pick_bn(M, K) == 256)
    return obf ? launch<false, true, 256, false, kEpiStore, true>(tmA, tmB, tmS, tmO, p, st)
               : launch<false, true, 256, false, kEpiStore, false>(tmA, tmB, tmS, tmO, p, st);
  return obf ? launch<false, true, 128, false, kEpiStore, true>(tmA, tmB, tmS, tmO, p, st)
             : launch<false, true, 128, false, kEpiStore, false>(tmA, tmB, tmS, tmO, p, st);
}

extern "C" int crv_masked_linear_bwd_ds(const uint16_t* dy, const uint16_t* x, const float* w, float* dscores,
                                        int accumulate, int M, int N, int K, void* stream) {
  if (!dy || !x || !w || !dscores || M <= 0 || N <= 0 || K <= 0) return CRV_E_BADARG;
  if ((N % 8) || (K % 8)) return CRV_E_SHAPE;
  if (!aligned16(dy) || !aligned16(x) || !aligned16(w) || !aligned16(dscores)) return CRV_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // generic: MM = N, NN = K, KK = M;  A = dY [M][N] = [KK][MM] MN-major;  B = X [M][K] = [KK][NN] MN-major
  const int num_kb = (M + BK - 1) / BK;
  // 256-wide tiles need 62 instead of 94 operand bytes per MMA cycle from L2; the reduction over M (144 k-blocks at
  // batch 256) is split across blockIdx.z so that one wave of CTAs covers the machine even for 768 x 768 modules
  const bool two = use_2cta(N, K);
  const int bn = (K % 256 == 0) ? 256 : 128;
  const int tiles = two ? ((N + 255) / 256) * (K / 256) : ((N + BM - 1) / BM) * ((K + bn - 1) / bn);
  int splits = (two ? gemm_sms() / 2 : gemm_sms()) / tiles;
  const int max_splits = (num_kb + 7) / 8;  // keep >= 8 k-blocks per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  GemmParams p{};
  p.MM = N; p.NN = K; p.KK = M;
  p.kb_per_split = (num_kb + splits - 1) / splits;
  p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.w = w;
  const bool add = accumulate != 0 && accumulate != CRV_DS_ZEROED;     // CRV_DS_ZEROED: dS already holds zeros
  p.reduce_out = (p.splits > 1 || add) ? 1 : 0;
  if (p.splits > 1 && accumulate == 0)
    CRV_CUDA(cudaMemsetAsync(dscores, 0, static_cast<size_t>(N) * K * sizeof(float), st));
  CUtensorMap tmA, tmB, tmO;
  int rc;
  if ((rc = make_map(&tmA, dy, 2, false, M, N, 64, 64))) return rc;
  if ((rc = make_map(&tmB, x, 2, false, M, K, 64, 64))) return rc;
  if ((rc = make_out_map(&tmO, dscores, false, N, K))) return rc;
  if (two) return launch2<true, true, kEpiScoreGrad, false>(tmA, tmB, tmO, p, st);
  if (bn == 256) return launch<true, true, 256, false, kEpiScoreGrad, false>(tmA, tmB, tmB, tmO, p, st);
  return launch<true, true, 128, false, kEpiScoreGrad, false>(tmA, tmB, tmB, tmO, p, st);
}

// ------------------------------------------------------------------------------------------------
// grouped launch: host side
// ------------------------------------------------------------------------------------------------
namespace crv {

static bool group_eligible(const crv_gemm_problem& q) {
  int MM, NN;
  if (q.kind == CRV_GEMM_FWD) { MM = q.M; NN = q.N; }
  else if (q.kind == CRV_GEMM_DX) { MM = q.M; NN = q.K; }
  else { MM = q.N; NN = q.K; }
  if (!use_2cta(MM, NN)) return false;
  if (q.act != CRV_ACT_NONE && q.out_dtype != CRV_DTYPE_BF16) return false;
  return true;
}

// Static round-robin over CTA pairs: tile t runs on pair t % pairs.  The split of a score-gradient problem's
// reduction decides both how many tiles it contributes and how long each is; pick, per problem, the split that
// minimises the longest pair (cost of a tile = its k-blocks + a constant for fill / epilogue; split tiles pay a
// little more for the reduce-add).  Problems are few and the candidates a handful, so this is a few microseconds
// of host time per DISTINCT group shape (memoised).
static int simulate_makespan(const GArgs& g, int pairs) {
  static thread_local int load[128];
  for (int i = 0; i < pairs; ++i) load[i] = 0;
  int t = 0;
  for (int i = 0; i < g.count; ++i) {
    const GProblem& p = g.p[i];
    const int total_kb = (p.KK + BK - 1) / BK;
    const int per_z = p.num_m * p.num_n;
    for (int z = 0; z < p.splits; ++z) {
      int kb = total_kb - z * p.kb_per_split;
      if (kb > p.kb_per_split) kb = p.kb_per_split;
      const int cost = kb + 3 + (p.epi == kGEpiScoreGrad && p.reduce_out ? 1 : 0);
      for (int r = 0; r < per_z; ++r, ++t) load[t % pairs] += cost;
    }
  }
  int mx = 0;
  for (int i = 0; i < pairs; ++i) mx = load[i] > mx ? load[i] : mx;
  return mx;
}

static void set_split(GProblem& p, int splits, int must_reduce) {
  const int total_kb = (p.KK + BK - 1) / BK;
  if (splits < 1) splits = 1;
  p.kb_per_split = (total_kb + splits - 1) / splits;
  p.splits = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.reduce_out = (p.splits > 1 || must_reduce) ? 1 : 0;
}

static void finish_tiles(GArgs& g) {
  int t = 0;
  for (int i = 0; i < g.count; ++i) {
    t += g.p[i].num_m * g.p[i].num_n * g.p[i].splits;
    g.p[i].tile_end = t;
  }
  g.total_tiles = t;
}

static void choose_splits(GArgs& g, const int* must_reduce, int pairs) {
  static const int cand[] = {1, 2, 3, 4, 6, 8, 9, 12, 16, 18, 24};
  for (int i = 0; i < g.count; ++i) {
    GProblem& p = g.p[i];
    if (p.epi != kGEpiScoreGrad) continue;
    const int total_kb = (p.KK + BK - 1) / BK;
    int best = 1, best_cost = 1 << 30;
    for (int c : cand) {
      if (c > 1 && total_kb / c < 8) break;   // keep >= 8 k-blocks per split
      set_split(p, c, must_reduce[i]);
      const int cost = simulate_makespan(g, pairs);
      if (cost < best_cost) { best_cost = cost; best = c; }
    }
    set_split(p, best, must_reduce[i]);
  }
}

template <int EPIMASK>
static int launch_group_t(GArgs& g, cudaStream_t stream) {
  using L = SmemG<5, 2>;
  auto kern = grouped_gemm2_kernel<5, 2, EPIMASK>;
  static bool configured = false;
  if (!configured) {
    CRV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  const int max_pairs = gemm_sms() / 2;
  const int pairs = g.total_tiles < max_pairs ? g.total_tiles : max_pairs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(64 + 32 * L::kEpiWarps);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  CRV_CUDA(cudaLaunchKernelEx(&cfg, kern, g));
  return launch_status();
}

// Instantiations by what a group can hold: forward (bf16 / fp32 stores), forward with GELU, backward (dX stores +
// score gradients), backward with gelu'.  The smallest one that covers the group's epilogue kinds is launched.
static int launch_group(GArgs& g, cudaStream_t stream) {
  constexpr int F = 1 << kGEpiF32, B = 1 << kGEpiBf16, S = 1 << kGEpiScoreGrad, G = 1 << kGEpiGelu, D = 1 << kGEpiGeluGrad;
  int mask = 0;
  for (int i = 0; i < g.count; ++i) mask |= 1 << g.p[i].epi;
  if ((mask & ~(F | B)) == 0) return launch_group_t<F | B>(g, stream);
  if ((mask & ~(B | G)) == 0) return launch_group_t<B | G>(g, stream);
  if ((mask & ~(F | B | S)) == 0) return launch_group_t<F | B | S>(g, stream);
  if ((mask & ~(B | S | D)) == 0) return launch_group_t<B | S | D>(g, stream);
  return launch_group_t<F | B | S | G | D>(g, stream);
}

static int single_launch(const crv_gemm_problem& q, void* stream) {
  if (q.act != CRV_ACT_NONE) return CRV_E_SHAPE;
  if (q.kind == CRV_GEMM_FWD)
    return crv_masked_linear_fwd(q.a, q.b, nullptr, nullptr, q.bias, q.out, q.out_dtype, q.M, q.N, q.K, stream);
  if (q.kind == CRV_GEMM_DX)
    return crv_masked_linear_bwd_dx(q.a, q.b, nullptr, nullptr, q.out, q.out_dtype, q.M, q.N, q.K, stream);
  return crv_masked_linear_bwd_ds(q.a, q.b, q.w_f32, static_cast<float*>(q.out), q.accumulate, q.M, q.N, q.K, stream);
}

}  // namespace crv

extern "C" int crv_masked_gemm_grouped(const crv_gemm_problem* pr, int count, void* stream) {
  if (!pr || count <= 0) return CRV_E_BADARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < count; ++i) {
    const crv_gemm_problem& q = pr[i];
    if (!q.a || !q.b || !q.out || q.M <= 0 || q.N <= 0 || q.K <= 0) return CRV_E_BADARG;
    if (q.kind < CRV_GEMM_FWD || q.kind > CRV_GEMM_DS) return CRV_E_BADARG;
    if (q.kind == CRV_GEMM_DS && (!q.w_f32 || q.act != CRV_ACT_NONE)) return CRV_E_BADARG;
    if (q.act != CRV_ACT_NONE && (q.act != CRV_ACT_GELU || !q.aux)) return CRV_E_BADARG;
    if (q.out_dtype != CRV_DTYPE_F32 && q.out_dtype != CRV_DTYPE_BF16) return CRV_E_BADARG;
    if ((q.N % 8) || (q.K % 8)) return CRV_E_SHAPE;
    if (!aligned16(q.a) || !aligned16(q.b) || !aligned16(q.out) || (q.aux && !aligned16(q.aux)) ||
        (q.w_f32 && !aligned16(q.w_f32)))
      return CRV_E_ALIGN;
  }
  const int pairs = gemm_sms() / 2;
  int i = 0;
  while (i < count) {
    if (!group_eligible(pr[i])) {
      int rc = single_launch(pr[i], stream);
      if (rc) return rc;
      ++i;
      continue;
    }
    GArgs g{};
    int must_reduce[kMaxGroup] = {0, 0, 0, 0};   // add into dS whatever the split
    int holds_junk[kMaxGroup] = {0, 0, 0, 0};    // dS must be cleared before a reduce-add
    int src[kMaxGroup];
    while (i < count && g.count < kMaxGroup && group_eligible(pr[i])) {
      const crv_gemm_problem& q = pr[i];
      GProblem& p = g.p[g.count];
      const bool obf = q.out_dtype == CRV_DTYPE_BF16;
      int rc;
      if (q.kind == CRV_GEMM_FWD) {            // A = X [M,K] K-major; B = Wm [N,K] K-major
        p.MM = q.M; p.NN = q.N; p.KK = q.K; p.a_mn = 0; p.b_mn = 0;
        if ((rc = make_map(&p.tmA, q.a, 2, false, q.M, q.K, BM, BK))) return rc;
        if ((rc = make_map(&p.tmB, q.b, 2, false, q.N, q.K, 128, BK))) return rc;
        if ((rc = make_out_map(&p.tmOut, q.out, obf, q.M, q.N))) return rc;
        p.bias = q.bias;
        p.epi = obf ? kGEpiBf16 : kGEpiF32;
        if (q.act == CRV_ACT_GELU) {
          if ((rc = make_out_map(&p.tmAux, q.aux, true, q.M, q.N))) return rc;
          p.epi = kGEpiGelu;
        }
      } else if (q.kind == CRV_GEMM_DX) {      // A = dY [M,N] K-major; B = Wm [N,K] = [KK][NN] MN-major
        p.MM = q.M; p.NN = q.K; p.KK = q.N; p.a_mn = 0; p.b_mn = 1;
        if ((rc = make_map(&p.tmA, q.a, 2, false, q.M, q.N, BM, BK))) return rc;
        if ((rc = make_map(&p.tmB, q.b, 2, false, q.N, q.K, 64, 64))) return rc;
        if ((rc = make_out_map(&p.tmOut, q.out, obf, q.M, q.K))) return rc;
        p.epi = obf ? kGEpiBf16 : kGEpiF32;
        if (q.act == CRV_ACT_GELU) {
          p.aux_in = static_cast<const uint16_t*>(q.aux);
          p.epi = kGEpiGeluGrad;
        }
      } else {                                  // A = dY [M,N] = [KK][MM] MN-major; B = X [M,K] = [KK][NN] MN-major
        p.MM = q.N; p.NN = q.K; p.KK = q.M; p.a_mn = 1; p.b_mn = 1;
        if ((rc = make_map(&p.tmA, q.a, 2, false, q.M, q.N, 64, 64))) return rc;
        if ((rc = make_map(&p.tmB, q.b, 2, false, q.M, q.K, 64, 64))) return rc;
        if ((rc = make_out_map(&p.tmOut, q.out, false, q.N, q.K))) return rc;
        p.w = q.w_f32;
        p.epi = kGEpiScoreGrad;
        must_reduce[g.count] = (q.accumulate != 0 && q.accumulate != CRV_DS_ZEROED) ? 1 : 0;
        holds_junk[g.count] = q.accumulate == 0 ? 1 : 0;
      }
      p.num_m = (p.MM + 255) / 256;
      p.num_n = (p.NN + 255) / 256;
      set_split(p, 1, must_reduce[g.count]);
      src[g.count] = i;
      ++g.count;
      ++i;
    }
    // two score-gradient problems of one group that write the same dS (a shared module applied to both modalities)
    // run concurrently: both must reduce-add, into a buffer that is cleared once if it holds junk
    bool dup[kMaxGroup] = {false, false, false, false};
    for (int a = 0; a < g.count; ++a)
      for (int b = a + 1; b < g.count; ++b)
        if (g.p[a].epi == kGEpiScoreGrad && g.p[b].epi == kGEpiScoreGrad && pr[src[a]].out == pr[src[b]].out) {
          dup[b] = true;
          must_reduce[a] = must_reduce[b] = 1;
        }
    finish_tiles(g);
    choose_splits(g, must_reduce, pairs);
    finish_tiles(g);
    for (int a = 0; a < g.count; ++a) {
      if (g.p[a].epi != kGEpiScoreGrad || dup[a]) continue;
      const crv_gemm_problem& q = pr[src[a]];
      if (holds_junk[a] && g.p[a].reduce_out)
        CRV_CUDA(cudaMemsetAsync(q.out, 0, static_cast<size_t>(q.N) * q.K * sizeof(float), st));
    }
    int rc = launch_group(g, st);
    if (rc) return rc;
  }
  return CRV_OK;
}
