// Exact k-th smallest value of many fp32 segments at once (Trainer.reset_threshold: 168 x
// torch.kthvalue, hg_transformers/mask_trainer_Robust_VQA.py:467-482; magnitude init,
// masking/maskers.py:204-215).
//
// Algorithm: most-significant-digit radix select over order-preserving 32-bit keys, 11 + 11 + 10
// bits.  Every pass is ONE persistent launch over all segments: the work is cut into chunks of
// kChunk elements that never straddle a segment, each CTA walks a contiguous range of chunks with
// 16-byte coalesced loads and keeps a 2048-bin histogram in shared memory, flushing it to the
// per-segment global histogram when the segment changes.  A one-CTA-per-segment scan kernel then
// locates the bin that holds rank k, narrows (prefix, k) and clears the histogram for the next pass.
// The result is the exact order statistic: bit-identical to sorting the segment.
//
// Key order: -inf < ... < -0 < +0 < ... < +inf < NaN (torch.kthvalue's CUDA ordering).
#include <vector>

#include "common.cuh"

namespace crv {

constexpr int kBins = 2048;
constexpr int kChunk = 8192;        // elements per work chunk (32 KB)
constexpr int kSelThreads = 256;

struct SegState {
  const float* ptr;
  long long n;
  long long k;             // remaining 1-based rank inside the current prefix bucket
  unsigned int prefix;     // key bits decided so far (high bits)
  int pad;
};

struct SelHeader {         // lives at the start of the workspace
  int count;
  int total_chunks;
};

__device__ __forceinline__ unsigned int float_key(float x, bool use_abs) {
  unsigned int u = __float_as_uint(x);
  if (use_abs) u &= 0x7FFFFFFFu;
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 0xFFFFFFFFu;  // NaN sorts last
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned int key) {
  if (key == 0xFFFFFFFFu) return __uint_as_float(0x7FC00000u);
  const unsigned int u = (key & 0x80000000u) ? (key & 0x7FFFFFFFu) : ~key;
  return __uint_as_float(u);
}

// pass 0: bits [31:21]   pass 1: bits [20:10] under prefix mask 0xFFE00000   pass 2: bits [9:0] under 0xFFFFFC00
__device__ __forceinline__ void pass_geometry(int pass, unsigned int& hi_mask, int& shift, unsigned int& bin_mask) {
  if (pass == 0) { hi_mask = 0u; shift = 21; bin_mask = 0x7FFu; }
  else if (pass == 1) { hi_mask = 0xFFE00000u; shift = 10; bin_mask = 0x7FFu; }
  else { hi_mask = 0xFFFFFC00u; shift = 0; bin_mask = 0x3FFu; }
}

__global__ void __launch_bounds__(kSelThreads)
select_hist_kernel(const SelHeader* __restrict__ hdr, const SegState* __restrict__ segs,
                   const int* __restrict__ cum_chunks, unsigned int* __restrict__ hist, int pass, int use_abs) {
  __shared__ unsigned int sh[kBins];
  const int count = hdr->count;
  const int total = hdr->total_chunks;
  const int per = (total + gridDim.x - 1) / gridDim.x;
  const int c_begin = blockIdx.x * per;
  const int c_end = min(c_begin + per, total);
  if (c_begin >= c_end) return;
  unsigned int hi_mask, bin_mask;
  int shift;
  pass_geometry(pass, hi_mask, shift, bin_mask);

  for (int i = threadIdx.x; i < kBins; i += kSelThreads) sh[i] = 0;
  __syncthreads();

  // locate the segment of the first chunk (upper bound in cum_chunks)
  int seg;
  {
    int lo = 0, hi = count;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cum_chunks[mid + 1] <= c_begin) lo = mid + 1; else hi = mid;
    }
    seg = lo;
  }
  for (int c = c_begin; c < c_end; ++c) {
    if (c >= cum_chunks[seg + 1]) {
      // segment change: flush and clear
      __syncthreads();
      for (int i = threadIdx.x; i < kBins; i += kSelThreads) {
        const unsigned int v = sh[i];
        if (v) { atomicAdd(&hist[static_cast<size_t>(seg) * kBins + i], v); sh[i] = 0; }
      }
      __syncthreads();
      while (c >= cum_chunks[seg + 1]) ++seg;
    }
    const SegState st = segs[seg];
    const unsigned int prefix = st.prefix;
    const long long off = static_cast<long long>(c - cum_chunks[seg]) * kChunk;
    const long long len = min(static_cast<long long>(kChunk), st.n - off);
    const float* base = st.ptr + off;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) == 0) {
      const int nvec = static_cast<int>(len >> 2);
      for (int i = threadIdx.x; i < nvec; i += kSelThreads) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(base) + i);
        const unsigned int k0 = float_key(v.x, use_abs), k1 = float_key(v.y, use_abs);
        const unsigned int k2 = float_key(v.z, use_abs), k3 = float_key(v.w, use_abs);
        const bool a0 = (k0 & hi_mask) == prefix, a1 = (k1 & hi_mask) == prefix;
        const bool a2 = (k2 & hi_mask) == prefix, a3 = (k3 & hi_mask) == prefix;
        const unsigned int b0 = (k0 >> shift) & bin_mask, b1 = (k1 >> shift) & bin_mask;
        const unsigned int b2 = (k2 >> shift) & bin_mask, b3 = (k3 >> shift) & bin_mask;
        if (a0 && a1 && a2 && a3 && b0 == b1 && b1 == b2 && b2 == b3) {
          atomicAdd(&sh[b0], 4u);  // tie-heavy data (scores start as exactly {0, 0.02})
        } else {
          if (a0) atomicAdd(&sh[b0], 1u);
          if (a1) atomicAdd(&sh[b1], 1u);
          if (a2) atomicAdd(&sh[b2], 1u);
          if (a3) atomicAdd(&sh[b3], 1u);
        }
      }
      for (int i = (nvec << 2) + threadIdx.x; i < len; i += kSelThreads) {
        const unsigned int k = float_key(base[i], use_abs);
        if ((k & hi_mask) == prefix) atomicAdd(&sh[(k >> shift) & bin_mask], 1u);
      }
    } else {
      for (int i = threadIdx.x; i < len; i += kSelThreads) {
        const unsigned int k = float_key(base[i], use_abs);
        if ((k & hi_mask) == prefix) atomicAdd(&sh[(k >> shift) & bin_mask], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kBins; i += kSelThreads) {
    const unsigned int v = sh[i];
    if (v) atomicAdd(&hist[static_cast<size_t>(seg) * kBins + i], v);
  }
}

// one CTA (kBins/2 = 1024 threads, 2 bins each) per segment
__global__ void __launch_bounds__(1024)
select_scan_kernel(SegState* __restrict__ segs, unsigned int* __restrict__ hist, float* __restrict__ out, int pass) {
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned int found_bin;
  __shared__ unsigned long long found_below;
  const int seg = blockIdx.x;
  unsigned int* h = hist + static_cast<size_t>(seg) * kBins;
  const int t = threadIdx.x;
  const unsigned int c0 = h[2 * t], c1 = h[2 * t + 1];
  h[2 * t] = 0;
  h[2 * t + 1] = 0;
  unsigned long long incl = static_cast<unsigned long long>(c0) + c1;
  const int lane = t & 31, wid = t >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) warp_tot[wid] = incl;
  if (t == 0) { found_bin = 0xFFFFFFFFu; found_below = 0; }
  __syncthreads();
  if (wid == 0) {
    unsigned long long w = warp_tot[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long up = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += up;
    }
    warp_tot[lane] = w;  // inclusive totals per warp
  }
  __syncthreads();
  const unsigned long long before_warp = wid ? warp_tot[wid - 1] : 0ull;
  const unsigned long long excl0 = before_warp + incl - c0 - c1;  // elements in bins < 2t
  const unsigned long long k = static_cast<unsigned long long>(segs[seg].k);
  // rank k lies in the first bin whose inclusive count reaches k
  if (c0 && k > excl0 && k <= excl0 + c0) { found_bin = 2 * t; found_below = excl0; }
  if (c1 && k > excl0 + c0 && k <= excl0 + c0 + c1) { found_bin = 2 * t + 1; found_below = excl0 + c0; }
  __syncthreads();
  if (t == 0) {
    unsigned int hi_mask, bin_mask;
    int shift;
    pass_geometry(pass, hi_mask, shift, bin_mask);
    SegState st = segs[seg];
    if (found_bin == 0xFFFFFFFFu) {
      // k out of range (k > n): clamp to the largest populated bin is not torch semantics; flag with NaN
      if (pass == 2) out[seg] = __uint_as_float(0x7FC00000u);
    } else {
      st.prefix |= found_bin << shift;
      st.k = static_cast<long long>(k - found_below);
      segs[seg] = st;
      if (pass == 2) out[seg] = key_float(st.prefix);
    }
  }
}

}  // namespace crv

using namespace crv;

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" size_t crv_kth_value_workspace_bytes(int count) {
  if (count <= 0) return 0;
  size_t b = 256;                                              // header
  b += align_up(sizeof(SegState) * count, 256);                // segment states
  b += align_up(sizeof(int) * (count + 1), 256);               // cumulative chunk counts
  b += align_up(sizeof(unsigned int) * kBins * count, 256);    // histograms
  return b;
}

extern "C" int crv_kth_value_batched(const float* const* ptrs_host, const long long* n_host, const long long* k_host,
                                     int count, int use_abs, float* thr_out, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  if (!ptrs_host || !n_host || !k_host || !thr_out || !workspace || count <= 0) return CRV_E_BADARG;
  if (workspace_bytes < crv_kth_value_workspace_bytes(count)) return CRV_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return CRV_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  // stage the small descriptor block on the host, one H2D copy
  const size_t off_seg = 256;
  const size_t off_cum = off_seg + align_up(sizeof(SegState) * count, 256);
  const size_t off_hist = off_cum + align_up(sizeof(int) * (count + 1), 256);
  const size_t stage_bytes = off_hist;
  // pageable staging: cudaMemcpyAsync from pageable memory returns after the source was consumed
  std::vector<unsigned char> host_buf(stage_bytes, 0);
  unsigned char* host = host_buf.data();
  SelHeader* hh = reinterpret_cast<SelHeader*>(host);
  SegState* hs = reinterpret_cast<SegState*>(host + off_seg);
  int* hc = reinterpret_cast<int*>(host + off_cum);
  long long total = 0;
  hc[0] = 0;
  for (int i = 0; i < count; ++i) {
    if (!ptrs_host[i] || n_host[i] <= 0 || k_host[i] < 1 || k_host[i] > n_host[i] ||
        (reinterpret_cast<uintptr_t>(ptrs_host[i]) & 3u)) {
      return CRV_E_BADARG;
    }
    hs[i].ptr = ptrs_host[i];
    hs[i].n = n_host[i];
    hs[i].k = k_host[i];
    hs[i].prefix = 0;
    hs[i].pad = 0;
    total += (n_host[i] + kChunk - 1) / kChunk;
    if (total > 0x7FFFFFFF) return CRV_E_SHAPE;
    hc[i + 1] = static_cast<int>(total);
  }
  hh->count = count;
  hh->total_chunks = static_cast<int>(total);
  unsigned char* dev = static_cast<unsigned char*>(workspace);
  cudaError_t e = cudaMemcpyAsync(dev, host, stage_bytes, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dev + off_hist, 0, sizeof(unsigned int) * kBins * count, st);
  if (e != cudaSuccess) return record(e);

  const SelHeader* dh = reinterpret_cast<const SelHeader*>(dev);
  SegState* ds = reinterpret_cast<SegState*>(dev + off_seg);
  const int* dc = reinterpret_cast<const int*>(dev + off_cum);
  unsigned int* dhist = reinterpret_cast<unsigned int*>(dev + off_hist);
  long long g = static_cast<long long>(num_sms()) * 8;
  if (g > total) g = total;
  int rc = CRV_OK;
  for (int pass = 0; pass < 3 && rc == CRV_OK; ++pass) {
    select_hist_kernel<<<static_cast<int>(g), kSelThreads, 0, st>>>(dh, ds, dc, dhist, pass, use_abs);
    rc = launch_status();
    if (rc) break;
    select_scan_kernel<<<count, 1024, 0, st>>>(ds, dhist, thr_out, pass);
    rc = launch_status();
  }
  return rc;
}
