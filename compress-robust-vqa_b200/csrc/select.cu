// Exact k-th smallest value of many fp32 segments at once (Trainer.reset_threshold: 168 x
// torch.kthvalue, hg_transformers/mask_trainer_Robust_VQA.py:467-482; magnitude init,
// masking/maskers.py:204-215).
//
// Two layers, both exact:
//
//  (1) RADIX CORE.  Most-significant-digit radix select over order-preserving 32-bit keys (11 + 11 + 10
//      bits).  Every pass is one persistent launch over all segments: chunks of kChunk elements never
//      straddle a segment, each CTA walks a contiguous chunk range with 16-byte coalesced loads and keeps
//      a 2048-bin histogram in shared memory, flushed per segment; a one-CTA-per-segment scan narrows
//      (prefix, k).  Its cost is the shared-memory atomic per element (measured 880 GB/s over the 829 MB of
//      LXMERT scores), so it is only used on small inputs:
//
//  (2) SAMPLE -> FILTER -> SELECT front end (segments larger than kSample elements).
//      a. gather a stratified sample of kSample elements per segment and radix-select two sample order
//         statistics lo <= hi that bracket rank k with ~4 sigma margin;
//      b. ONE streaming pass over the data (the HBM-bound part: 4 B per element, read once, no atomics
//         per element): count keys < lo, == lo, == hi in registers and compact the few keys strictly
//         between the pivots (about 2.5 % of the data) into a candidate buffer;
//      c. the counts decide exactly where rank k falls: it is lo, or hi, or the (k - below)-th smallest
//         candidate, which the radix core then selects.  If the pivots miss (possible only with
//         adversarial data) or the candidate buffer overflows, the segment falls back to the radix core on
//         the full data.  The result is always the exact order statistic, bit-identical to sorting.
//
// Key order: -inf < ... < -0 < +0 < ... < +inf < NaN (torch.kthvalue's CUDA ordering).
#include <algorithm>
#include <climits>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace crv {

constexpr int kBins = 2048;
constexpr int kChunk = 8192;        // elements per work chunk (32 KB)
constexpr int kSelThreads = 256;
constexpr int kSample = 16384;      // sample size per large segment
constexpr long long kSmallMax = 32768;    // segments up to this size are selected by one CTA
constexpr int kCandDiv = 4;         // candidate buffer capacity = n / kCandDiv + kChunk

struct SegState {
  const float* ptr;        // data the radix core works on (original segment, its sample, or its candidates)
  long long n;
  long long k;             // remaining 1-based rank inside the current prefix bucket
  unsigned int prefix;     // key bits decided so far (high bits)
  int done;                // 1 = result already written (scan kernels skip the segment)
  int use_abs;             // keys are taken on |x|
  int out_index;           // where the scan of pass 2 writes the selected value
};

struct BigSeg {            // one segment handled by the sample/filter front end
  const float* ptr;
  long long n, k;
  float* cand;             // candidate buffer (capacity cap)
  long long cap;
  float lo_f, hi_f;              // pivots (filled from the sample select)
  int lo_open, hi_open;          // 1 = no lower / upper pivot (rank window ran off the sample)
  unsigned long long c_lt, c_eq_lo, c_eq_hi, c_mid;   // counters of the filter pass
  int overflow;
  int out_index;
  int inclusive;           // filter mode chosen by pivots_kernel
  int pad_;
};

struct SelHeader {         // lives at the start of a core workspace
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

// ------------------------------------------------------------------------------------------------ radix core
// chunk table built on the device from the segments' current n (they may have been set by the filter pass)
__global__ void __launch_bounds__(1024) select_plan_kernel(SelHeader* hdr, const SegState* segs, int* cum_chunks, int count) {
  __shared__ int sh[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) { carry = 0; cum_chunks[0] = 0; }
  __syncthreads();
  for (int base = 0; base < count; base += 1024) {
    const int i = base + threadIdx.x;
    int c = 0;
    if (i < count && !segs[i].done) c = static_cast<int>((segs[i].n + kChunk - 1) / kChunk);
    sh[threadIdx.x] = c;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {           // Hillis-Steele inclusive scan
      const int v = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += v;
      __syncthreads();
    }
    if (i < count) cum_chunks[i + 1] = carry + sh[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) { hdr->count = count; hdr->total_chunks = carry; }
}

__global__ void __launch_bounds__(kSelThreads)
select_hist_kernel(const SelHeader* __restrict__ hdr, const SegState* __restrict__ segs,
                   const int* __restrict__ cum_chunks, unsigned int* __restrict__ hist, int pass) {
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
    const bool use_abs = st.use_abs != 0;
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
          atomicAdd(&sh[b0], 4u);
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

// one CTA (1024 threads, 2 bins each) per segment.  key_out != null: pass 2 writes the selected KEY there
// (sample select); otherwise it writes the float value to out[seg.out_index].
__global__ void __launch_bounds__(1024)
select_scan_kernel(SegState* __restrict__ segs, unsigned int* __restrict__ hist, float* __restrict__ out,
                   unsigned int* __restrict__ key_out, int pass) {
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned int found_bin;
  __shared__ unsigned long long found_below;
  const int seg = blockIdx.x;
  unsigned int* h = hist + static_cast<size_t>(seg) * kBins;
  const int t = threadIdx.x;
  const unsigned int c0 = h[2 * t], c1 = h[2 * t + 1];
  h[2 * t] = 0;
  h[2 * t + 1] = 0;
  if (segs[seg].done) return;
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
    warp_tot[lane] = w;
  }
  __syncthreads();
  const unsigned long long before_warp = wid ? warp_tot[wid - 1] : 0ull;
  const unsigned long long excl0 = before_warp + incl - c0 - c1;
  const unsigned long long k = static_cast<unsigned long long>(segs[seg].k);
  if (c0 && k > excl0 && k <= excl0 + c0) { found_bin = 2 * t; found_below = excl0; }
  if (c1 && k > excl0 + c0 && k <= excl0 + c0 + c1) { found_bin = 2 * t + 1; found_below = excl0 + c0; }
  __syncthreads();
  if (t == 0) {
    unsigned int hi_mask, bin_mask;
    int shift;
    pass_geometry(pass, hi_mask, shift, bin_mask);
    SegState st = segs[seg];
    if (found_bin == 0xFFFFFFFFu) {
      if (pass == 2) {
        if (key_out) key_out[seg] = 0xFFFFFFFFu;
        else out[st.out_index] = __uint_as_float(0x7FC00000u);   // k out of range
      }
    } else {
      st.prefix |= found_bin << shift;
      st.k = static_cast<long long>(k - found_below);
      segs[seg] = st;
      if (pass == 2) {
        if (key_out) key_out[seg] = st.prefix;
        else out[st.out_index] = key_float(st.prefix);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------- one-CTA select
// Segments of at most max_n elements (the samples, and the candidates the filter pass leaves) are selected by
// ONE CTA each in one launch: a min/max sweep normalises the keys to key - min, so the radix digits start at the
// highest bit in which the segment's keys actually differ (candidates lie between two close pivots and would
// otherwise all fall into one bin of the first two passes -- 32-way shared-memory atomic conflicts), then
// ceil(bits / 11) histogram passes over the L2-resident data.  Marks the segment done so the multi-CTA core that
// follows skips it.
constexpr int kSmallThreads = 512;      // 4 bins per thread; 4 CTAs resident per SM
constexpr int kSmallBpt = kBins / kSmallThreads;
__device__ __forceinline__ void block_find_bin(const unsigned int* sh, unsigned long long k, unsigned long long* warp_tot,
                                               unsigned int* found_bin, unsigned long long* found_below) {
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  unsigned int c[kSmallBpt];
  unsigned long long mine = 0;
#pragma unroll
  for (int j = 0; j < kSmallBpt; ++j) { c[j] = sh[kSmallBpt * t + j]; mine += c[j]; }
  unsigned long long incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) warp_tot[wid] = incl;
  if (t == 0) { *found_bin = 0xFFFFFFFFu; *found_below = 0; }
  __syncthreads();
  if (wid == 0) {
    unsigned long long w = lane < kSmallThreads / 32 ? warp_tot[lane] : 0ull;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long up = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += up;
    }
    if (lane < kSmallThreads / 32) warp_tot[lane] = w;
  }
  __syncthreads();
  unsigned long long excl = (wid ? warp_tot[wid - 1] : 0ull) + incl - mine;
#pragma unroll
  for (int j = 0; j < kSmallBpt; ++j) {
    if (c[j] && k > excl && k <= excl + c[j]) { *found_bin = kSmallBpt * t + j; *found_below = excl; }
    excl += c[j];
  }
  __syncthreads();
}

template <typename F>
__device__ __forceinline__ void for_each_key(const float* __restrict__ ptr, long long n, bool use_abs, F&& f) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0) {
    const long long nvec = n >> 2;
    for (long long i = threadIdx.x; i < nvec; i += blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(ptr) + i);
      f(float_key(v.x, use_abs)); f(float_key(v.y, use_abs)); f(float_key(v.z, use_abs)); f(float_key(v.w, use_abs));
    }
    for (long long i = (nvec << 2) + threadIdx.x; i < n; i += blockDim.x) f(float_key(ptr[i], use_abs));
  } else {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) f(float_key(ptr[i], use_abs));
  }
}

__global__ void __launch_bounds__(kSmallThreads)
select_small_kernel(SegState* __restrict__ segs, float* __restrict__ out, unsigned int* __restrict__ key_out,
                    long long max_n) {
  __shared__ unsigned int sh[kBins];
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned int wmin[32], wmax[32];   // kSmallThreads / 32 used
  __shared__ unsigned int found_bin;
  __shared__ unsigned long long found_below;
  const int seg = blockIdx.x;
  const SegState st = segs[seg];
  if (st.done || st.n > max_n || st.n <= 0) return;          // uniform per CTA
  const bool use_abs = st.use_abs != 0;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;

  unsigned int mn = 0xFFFFFFFFu, mx = 0u;
  for_each_key(st.ptr, st.n, use_abs, [&](unsigned int k) { mn = min(mn, k); mx = max(mx, k); });
  mn = __reduce_min_sync(0xffffffffu, mn);
  mx = __reduce_max_sync(0xffffffffu, mx);
  if (lane == 0) { wmin[wid] = mn; wmax[wid] = mx; }
  __syncthreads();
  mn = lane < kSmallThreads / 32 ? wmin[lane] : 0xFFFFFFFFu;
  mx = lane < kSmallThreads / 32 ? wmax[lane] : 0u;
  mn = __reduce_min_sync(0xffffffffu, mn);
  mx = __reduce_max_sync(0xffffffffu, mx);

  const unsigned int range = mx - mn;
  const int bits = 32 - __clz(range);                          // 0 when every key is equal
  unsigned int prefix = 0;                                     // in the key - mn domain
  unsigned long long k = static_cast<unsigned long long>(st.k);
  bool missing = false;
  for (int p = (bits + 10) / 11 - 1; p >= 0; --p) {
    const int shift = 11 * p;
    const unsigned int hi_mask = shift + 11 >= 32 ? 0u : (0xFFFFFFFFu << (shift + 11));
    for (int j = t; j < kBins; j += kSmallThreads) sh[j] = 0;
    __syncthreads();
    for_each_key(st.ptr, st.n, use_abs, [&](unsigned int key) {
      const unsigned int kp = key - mn;
      if ((kp & hi_mask) == prefix) atomicAdd(&sh[(kp >> shift) & 0x7FFu], 1u);
    });
    __syncthreads();
    block_find_bin(sh, k, warp_tot, &found_bin, &found_below);
    if (found_bin == 0xFFFFFFFFu) { missing = true; break; }   // k out of range (uniform)
    prefix |= found_bin << shift;
    k -= found_below;
    __syncthreads();
  }
  if (t == 0) {
    const unsigned int key = missing ? 0xFFFFFFFFu : prefix + mn;
    if (key_out) key_out[seg] = key;
    else out[st.out_index] = key_float(key);
    segs[seg].done = 1;
  }
}

// ------------------------------------------------------------------------------------------------ front end
// stratified sample: kSample / 8 evenly spaced runs of 8 consecutive elements (one 32-byte sector each)
__global__ void sample_gather_kernel(const BigSeg* __restrict__ big, float* __restrict__ samples, int use_abs) {
  const BigSeg s = big[blockIdx.y];
  float* dst = samples + static_cast<size_t>(blockIdx.y) * kSample;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < kSample; j += gridDim.x * blockDim.x) {
    const long long idx = (static_cast<long long>(j >> 3) * s.n) / (kSample >> 3) + (j & 7);   // < n for n > 8 * kSample / 8 * 8
    float v = __ldg(s.ptr + idx);
    if (use_abs) v = fabsf(v);
    dst[j] = v;
  }
}

// pivots selected on the samples -> BigSeg, as floats: the streaming pass compares in the float domain (one FSETP
// per test instead of building an integer key per element).  Float order equals key order except that it
// cannot tell -0 from +0 -- neither can torch.kthvalue's "<" -- and that NaNs compare false to everything,
// which files them under "above the upper pivot", exactly where the key order puts them.  An open end becomes
// an infinity: nothing is below -inf / above +inf, and elements equal to it are counted like any tied pivot.
// If the sample holds (almost) no element equal to either pivot the segment is filtered in INCLUSIVE mode: the
// pivots themselves become candidates and only "below lo" is counted -- 5 instructions per element instead of
// 10.  Pivot values that are heavily tied in the data (scores never touched by a gradient stay exactly at
// their initial 0 / 0.02) would flood the candidate buffer that way, so those segments use the EXCLUSIVE mode
// with the two equality counters.  One CTA per big segment sweeps its kSample samples.
constexpr int kTieSamples = 8;          // more sample hits on a pivot than this => exclusive mode
__global__ void __launch_bounds__(256)
pivots_kernel(BigSeg* __restrict__ big, const unsigned int* __restrict__ sample_keys, const float* __restrict__ samples) {
  __shared__ unsigned int ties;
  const int i = blockIdx.x;
  const float lo = big[i].lo_open ? -INFINITY : key_float(sample_keys[2 * i]);
  const float hi = big[i].hi_open ? INFINITY : key_float(sample_keys[2 * i + 1]);
  if (threadIdx.x == 0) ties = 0;
  __syncthreads();
  const float4* sp = reinterpret_cast<const float4*>(samples + static_cast<size_t>(i) * kSample);
  unsigned int c = 0;
  for (int j = threadIdx.x; j < kSample / 4; j += blockDim.x) {
    const float4 v = sp[j];
    c += (v.x == lo) + (v.y == lo) + (v.z == lo) + (v.w == lo);
    c += (v.x == hi) + (v.y == hi) + (v.z == hi) + (v.w == hi);
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&ties, c);
  __syncthreads();
  if (threadIdx.x == 0) {
    big[i].lo_f = lo;
    big[i].hi_f = hi;
    // the two pivots themselves are sample members: up to 2 hits are not ties.  NaN pivots never compare equal and
    // send the segment to the fallback either way.
    big[i].inclusive = (ties <= 2 + kTieSamples && lo < hi) ? 1 : 0;
  }
}

// THE streaming pass: one read of every big segment.  Every CTA walks a contiguous range of kChunk-element
// chunks in two halves; all loads of a half are issued before the first compare.  The three counters live in
// registers and the candidates strictly between the pivots in a shared-memory stage; both are flushed only
// when the CTA moves on to another segment (or the stage could overflow), so the global atomics are per CTA
// and segment, not per chunk.
constexpr int kHalf = kChunk / 2;
constexpr int kStage = kHalf + 2048;    // floats staged per CTA (24 KB); flushed once more than 2048 are waiting
__global__ void __launch_bounds__(kSelThreads, 5)
filter_kernel(BigSeg* __restrict__ big, const int* __restrict__ cum_chunks, int nbig, int total_chunks, int use_abs) {
  __shared__ __align__(16) float stage[kStage];
  __shared__ float priv[(kHalf / kSelThreads) * kSelThreads];   // [slot][thread]: a thread's candidates of one half
  __shared__ unsigned int n_stage;
  __shared__ unsigned int red[3][kSelThreads / 32];
  __shared__ long long out_base;
  const int per = (total_chunks + gridDim.x - 1) / gridDim.x;
  const int c_begin = blockIdx.x * per;
  const int c_end = min(c_begin + per, total_chunks);
  if (c_begin >= c_end) return;
  int seg;
  {
    int lo = 0, hi = nbig;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cum_chunks[mid + 1] <= c_begin) lo = mid + 1; else hi = mid;
    }
    seg = lo;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool absval = use_abs != 0;
  if (threadIdx.x == 0) n_stage = 0;
  __syncthreads();

  unsigned int c_lt = 0, c_el = 0, c_eh = 0;      // a CTA sees < 2^32 elements
  float lo_f = 0.f, hi_f = 0.f;
  const float* seg_ptr = nullptr;
  long long seg_n = 0;
  int seg_first_chunk = 0;
  int cur = -1;

  auto flush_stage = [&](BigSeg* bs) {        // called by all threads; the stage is quiescent (barrier before)
    const unsigned int ns = n_stage;
    if (ns == 0) return;                      // uniform
    if (threadIdx.x == 0) {
      long long b = static_cast<long long>(atomicAdd(&bs->c_mid, static_cast<unsigned long long>(ns)));
      if (b + ns > bs->cap) { bs->overflow = 1; b = -1; }
      out_base = b;
    }
    __syncthreads();
    const long long b = out_base;
    if (b >= 0)
      for (unsigned int i = threadIdx.x; i < ns; i += kSelThreads) bs->cand[b + i] = stage[i];
    __syncthreads();
    if (threadIdx.x == 0) n_stage = 0;
    __syncthreads();
  };
  auto flush_counts = [&](BigSeg* bs) {
    unsigned int v0 = c_lt, v1 = c_el, v2 = c_eh;
    for (int o = 16; o; o >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, o);
      v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      v2 += __shfl_xor_sync(0xffffffffu, v2, o);
    }
    if (lane == 0) { red[0][wid] = v0; red[1][wid] = v1; red[2][wid] = v2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t0 = 0, t1 = 0, t2 = 0;
      for (int w = 0; w < kSelThreads / 32; ++w) { t0 += red[0][w]; t1 += red[1][w]; t2 += red[2][w]; }
      if (t0) atomicAdd(&bs->c_lt, t0);
      if (t1) atomicAdd(&bs->c_eq_lo, t1);
      if (t2) atomicAdd(&bs->c_eq_hi, t2);
    }
    __syncthreads();
    c_lt = c_el = c_eh = 0;
  };
  bool inclusive = false;
  // compare one element; returns whether it is a candidate
  auto visit_excl = [&](float raw) -> bool {  // candidates strictly between the pivots, ties counted
    const float x = absval ? fabsf(raw) : raw;
    c_lt += x < lo_f;
    c_el += x == lo_f;
    c_eh += x == hi_f;
    return x > lo_f && x < hi_f;
  };
  auto visit_incl = [&](float raw) -> bool {  // candidates in [lo, hi]
    const float x = absval ? fabsf(raw) : raw;
    c_lt += x < lo_f;
    return x >= lo_f && x <= hi_f;
  };
  auto visit_stage = [&](float raw) {         // scalar tail / unaligned path
    if (inclusive ? visit_incl(raw) : visit_excl(raw)) stage[atomicAdd(&n_stage, 1u)] = absval ? fabsf(raw) : raw;
  };

  for (int c = c_begin; c < c_end; ++c) {
    while (c >= cum_chunks[seg + 1]) ++seg;
    if (seg != cur) {
      __syncthreads();                        // the previous chunk's stage writes are complete
      if (cur >= 0) { flush_stage(big + cur); flush_counts(big + cur); }
      cur = seg;
      lo_f = big[seg].lo_f; hi_f = big[seg].hi_f; inclusive = big[seg].inclusive != 0;
      seg_ptr = big[seg].ptr; seg_n = big[seg].n;
      seg_first_chunk = cum_chunks[seg];
    }
    const long long off = static_cast<long long>(c - seg_first_chunk) * kChunk;
    const int len = static_cast<int>(min(static_cast<long long>(kChunk), seg_n - off));
    const float* base = seg_ptr + off;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) == 0) {
      const int nvec = len >> 2;
      constexpr int kPer = kHalf / 4 / kSelThreads;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        __syncthreads();
        const bool full = n_stage > static_cast<unsigned int>(kStage - kHalf);
        __syncthreads();                      // everyone has read n_stage before anyone adds to it again
        if (full) flush_stage(big + cur);
        float4 v[kPer];
        const int i0 = h * (kHalf / 4) + threadIdx.x;
#pragma unroll
        for (int u = 0; u < kPer; ++u)
          if (i0 + u * kSelThreads < nvec) v[u] = __ldcs(reinterpret_cast<const float4*>(base) + i0 + u * kSelThreads);
        // classify: a candidate goes to the next free slot of this thread's PRIVATE column of `priv` (a predicated
        // store and a predicated add, no atomics); afterwards the thread reserves its stage slots with ONE
        // shared-memory atomic and moves its few candidates over.
        unsigned int mine = 0;
        auto keep = [&](bool is_cand, float raw) {
          if (is_cand) { priv[mine * kSelThreads + threadIdx.x] = absval ? fabsf(raw) : raw; ++mine; }
        };
        if (inclusive) {
#pragma unroll
          for (int u = 0; u < kPer; ++u)
            if (i0 + u * kSelThreads < nvec) {
              keep(visit_incl(v[u].x), v[u].x); keep(visit_incl(v[u].y), v[u].y);
              keep(visit_incl(v[u].z), v[u].z); keep(visit_incl(v[u].w), v[u].w);
            }
        } else {
#pragma unroll
          for (int u = 0; u < kPer; ++u)
            if (i0 + u * kSelThreads < nvec) {
              keep(visit_excl(v[u].x), v[u].x); keep(visit_excl(v[u].y), v[u].y);
              keep(visit_excl(v[u].z), v[u].z); keep(visit_excl(v[u].w), v[u].w);
            }
        }
        if (mine) {
          const unsigned int pos = atomicAdd(&n_stage, mine);
          for (unsigned int j = 0; j < mine; ++j) stage[pos + j] = priv[j * kSelThreads + threadIdx.x];
        }
      }
      for (int i = (nvec << 2) + threadIdx.x; i < len; i += kSelThreads) visit_stage(base[i]);   // < 4 elements
    } else {
      for (int h = 0; h < 2; ++h) {
        __syncthreads();
        const bool full = n_stage > static_cast<unsigned int>(kStage - kHalf);
        __syncthreads();
        if (full) flush_stage(big + cur);
        const int end = min(len, (h + 1) * kHalf);
        for (int i = h * kHalf + threadIdx.x; i < end; i += kSelThreads) visit_stage(base[i]);
      }
    }
  }
  __syncthreads();
  flush_stage(big + cur);
  flush_counts(big + cur);
}

// exact placement of rank k from the filter counters; prepares the final radix segments
__global__ void decide_kernel(const BigSeg* __restrict__ big, SegState* __restrict__ segs, float* __restrict__ out,
                              int nbig, int use_abs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nbig) return;
  const BigSeg b = big[i];
  SegState st;
  st.prefix = 0;
  st.done = 0;
  st.use_abs = 0;            // candidates already hold |x| when use_abs
  st.out_index = b.out_index;
  const unsigned long long k = static_cast<unsigned long long>(b.k);
  // inclusive mode leaves both equality counters at zero and the same placement logic applies
  const unsigned long long eq_hi = b.lo_f == b.hi_f ? 0ull : b.c_eq_hi;   // tied pivots: counted once, as lo
  const unsigned long long below = b.c_lt, at_lo = below + b.c_eq_lo, mid = at_lo + b.c_mid, at_hi = mid + eq_hi;
  bool fallback = b.overflow != 0;
  if (!fallback) {
    if (k <= below) fallback = true;                       // the pivot pair missed rank k from above
    else if (k <= at_lo) { out[b.out_index] = b.lo_f; st.done = 1; }
    else if (k <= mid) { st.ptr = b.cand; st.n = static_cast<long long>(b.c_mid); st.k = static_cast<long long>(k - at_lo); }
    else if (k <= at_hi) { out[b.out_index] = b.hi_f; st.done = 1; }
    else fallback = true;                                  // ... or from below
  }
  if (fallback) { st.ptr = b.ptr; st.n = b.n; st.k = b.k; st.use_abs = use_abs; st.done = 0; }
  if (st.done) { st.ptr = b.ptr; st.n = 0; st.k = 1; }
  segs[b.out_index] = st;   // the big segment's slot in the final radix pass
}

}  // namespace crv

using namespace crv;

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// workspace layout ------------------------------------------------------------------------------------------
struct WsLayout {
  size_t hdrA, segA, cumA, histA;      // core A: small segments + (later) the big segments' final select
  size_t segS;                         // the sample select's descriptors (2 virtual segments per big segment)
  size_t big, cumB, desc_bytes, keys, samples, cand, total;
};

static WsLayout ws_layout(int count, long long cand_floats) {
  WsLayout L{};
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  // descriptor block: filled on the host and uploaded with ONE copy
  L.segA = take(sizeof(SegState) * count);
  L.segS = take(sizeof(SegState) * 2 * count);
  L.big = take(sizeof(BigSeg) * count);
  L.cumB = take(sizeof(int) * (count + 1));
  L.desc_bytes = o;
  L.hdrA = take(sizeof(SelHeader));
  L.cumA = take(sizeof(int) * (count + 1));
  L.histA = take(sizeof(unsigned int) * kBins * count);
  L.keys = take(sizeof(unsigned int) * 2 * count);
  L.samples = take(sizeof(float) * static_cast<size_t>(kSample) * count);
  L.cand = take(sizeof(float) * static_cast<size_t>(cand_floats));
  L.total = o;
  return L;
}

// Worst-case size without knowing the segments: callers that know them use crv_kth_value_workspace_bytes_for.
extern "C" size_t crv_kth_value_workspace_bytes(int count) {
  if (count <= 0) return 0;
  return ws_layout(count, 0).total;   // no candidate space: every segment takes the radix core
}

extern "C" size_t crv_kth_value_workspace_bytes_for(const long long* n_host, int count) {
  if (count <= 0 || !n_host) return 0;
  long long cand = 0;
  for (int i = 0; i < count; ++i)
    if (n_host[i] > 4 * kSample) cand += n_host[i] / kCandDiv + kChunk;
  return ws_layout(count, cand).total;
}

static int run_core(unsigned char* dev, size_t hdr, size_t seg, size_t cum, size_t hist, int count, float* out,
                    unsigned int* key_out, cudaStream_t st) {
  SelHeader* dh = reinterpret_cast<SelHeader*>(dev + hdr);
  SegState* ds = reinterpret_cast<SegState*>(dev + seg);
  int* dc = reinterpret_cast<int*>(dev + cum);
  unsigned int* dhist = reinterpret_cast<unsigned int*>(dev + hist);
  select_plan_kernel<<<1, 1024, 0, st>>>(dh, ds, dc, count);
  int rc = launch_status();
  for (int pass = 0; pass < 3 && rc == CRV_OK; ++pass) {
    select_hist_kernel<<<num_sms() * 8, kSelThreads, 0, st>>>(dh, ds, dc, dhist, pass);
    rc = launch_status();
    if (rc) break;
    select_scan_kernel<<<count, 1024, 0, st>>>(ds, dhist, out, key_out, pass);
    rc = launch_status();
  }
  return rc;
}

extern "C" int crv_kth_value_batched(const float* const* ptrs_host, const long long* n_host, const long long* k_host,
                                     int count, int use_abs, float* thr_out, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  if (!ptrs_host || !n_host || !k_host || !thr_out || !workspace || count <= 0) return CRV_E_BADARG;
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return CRV_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // which segments can take the sample/filter front end with the workspace we were given?
  long long cand_need = 0;
  std::vector<int> is_big(count, 0);
  for (int i = 0; i < count; ++i) {
    if (!ptrs_host[i] || n_host[i] <= 0 || k_host[i] < 1 || k_host[i] > n_host[i] ||
        (reinterpret_cast<uintptr_t>(ptrs_host[i]) & 3u))
      return CRV_E_BADARG;
    if (n_host[i] > 4 * kSample) { is_big[i] = 1; cand_need += n_host[i] / kCandDiv + kChunk; }
  }
  WsLayout L = ws_layout(count, cand_need);
  if (workspace_bytes < L.total) {       // not enough room for candidates: radix core for everything
    cand_need = 0;
    std::fill(is_big.begin(), is_big.end(), 0);
    L = ws_layout(count, 0);
    if (workspace_bytes < L.total) return CRV_E_WORKSPACE;
  }
  unsigned char* dev = static_cast<unsigned char*>(workspace);

  // host image of the descriptor block (pageable: cudaMemcpyAsync returns after the source was consumed)
  std::vector<unsigned char> desc(L.desc_bytes, 0);
  SegState* segA = reinterpret_cast<SegState*>(desc.data() + L.segA);
  SegState* segS = reinterpret_cast<SegState*>(desc.data() + L.segS);
  BigSeg* big = reinterpret_cast<BigSeg*>(desc.data() + L.big);
  int* cumB = reinterpret_cast<int*>(desc.data() + L.cumB);
  cumB[0] = 0;
  float* cand_ptr = reinterpret_cast<float*>(dev + L.cand);
  float* samples = reinterpret_cast<float*>(dev + L.samples);
  int nbig = 0;
  for (int i = 0; i < count; ++i) {
    SegState s{};
    s.ptr = ptrs_host[i]; s.n = n_host[i]; s.k = k_host[i]; s.prefix = 0; s.use_abs = use_abs; s.out_index = i;
    s.done = is_big[i] ? 1 : 0;          // big segments are filled in by decide_kernel
    segA[i] = s;
    if (!is_big[i]) continue;
    BigSeg b{};
    b.ptr = ptrs_host[i]; b.n = n_host[i]; b.k = k_host[i]; b.out_index = i;
    b.cap = n_host[i] / kCandDiv + kChunk;
    b.cand = cand_ptr;
    cand_ptr += b.cap;
    // sample ranks bracketing k with ~4 sigma of the sample quantile
    const double p = static_cast<double>(k_host[i]) / static_cast<double>(n_host[i]);
    const double r = p * kSample;
    const double delta = 4.0 * sqrt(kSample * p * (1.0 - p)) + 8.0;
    long long r_lo = static_cast<long long>(floor(r - delta)), r_hi = static_cast<long long>(ceil(r + delta));
    b.lo_open = r_lo < 1;
    b.hi_open = r_hi > kSample;
    if (r_lo < 1) r_lo = 1;
    if (r_hi > kSample) r_hi = kSample;
    for (int e = 0; e < 2; ++e) {
      SegState v{};
      v.ptr = samples + static_cast<size_t>(nbig) * kSample;
      v.n = kSample; v.k = e ? r_hi : r_lo; v.prefix = 0; v.done = 0; v.use_abs = 0; v.out_index = 2 * nbig + e;
      segS[2 * static_cast<size_t>(nbig) + e] = v;
    }
    big[nbig] = b;
    cumB[nbig + 1] = cumB[nbig] + static_cast<int>((n_host[i] + kChunk - 1) / kChunk);
    ++nbig;
  }
  int rc = CRV_OK;
  BigSeg* dbig = reinterpret_cast<BigSeg*>(dev + L.big);
  SegState* dsegA = reinterpret_cast<SegState*>(dev + L.segA);
  CRV_CUDA(cudaMemsetAsync(dev + L.histA, 0, sizeof(unsigned int) * kBins * count, st));
  CRV_CUDA(cudaMemcpyAsync(dev, desc.data(), L.desc_bytes, cudaMemcpyHostToDevice, st));
  if (nbig) {
    // a. sample + pivot select
    sample_gather_kernel<<<dim3(8, nbig), 256, 0, st>>>(dbig, samples, use_abs);
    if ((rc = launch_status())) return rc;
    unsigned int* keys = reinterpret_cast<unsigned int*>(dev + L.keys);
    select_small_kernel<<<2 * nbig, kSmallThreads, 0, st>>>(reinterpret_cast<SegState*>(dev + L.segS), nullptr, keys, LLONG_MAX);
    if ((rc = launch_status())) return rc;
    pivots_kernel<<<nbig, 256, 0, st>>>(dbig, keys, samples);
    if ((rc = launch_status())) return rc;
    // b. the streaming pass
    const int total_chunks = cumB[nbig];
    const int grid = total_chunks < num_sms() * 5 ? total_chunks : num_sms() * 5;   // 5 resident CTAs per SM
    filter_kernel<<<grid, kSelThreads, 0, st>>>(dbig, reinterpret_cast<const int*>(dev + L.cumB), nbig, total_chunks,
                                               use_abs);
    if ((rc = launch_status())) return rc;
  }
  // c. small segments + whatever the big ones left over go through the radix core together
  if (nbig) {
    // the counters decide where rank k falls; the final-select descriptor of every big segment goes into its slot
    decide_kernel<<<(nbig + 127) / 128, 128, 0, st>>>(dbig, dsegA, thr_out, nbig, use_abs);
    if ((rc = launch_status())) return rc;
  }
  select_small_kernel<<<count, kSmallThreads, 0, st>>>(dsegA, thr_out, nullptr, kSmallMax);
  if ((rc = launch_status())) return rc;
  return run_core(dev, L.hdrA, L.segA, L.cumA, L.histA, count, thr_out, nullptr, st);   // segments above kSmallMax
}
