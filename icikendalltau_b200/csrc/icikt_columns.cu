// icikt_columns.cu -- K1: per-column preprocessing, done once per column instead of twice per
// pair as in the reference.
//
//   build_keys_kernel   setup_missing_matrix (R/utils.R:1-23) + the NA -> "below the minimum"
//                       substitution of src/kendallc.cpp:214-219, expressed as a sort key:
//                       missing rows get key 0, every other value an order-preserving 64-bit
//                       image of the double (-0.0 == +0.0, as compare_self :15-31 sees them).
//   segmented sort      sortedIndex (src/kendallc.cpp:6-12), one segment per column
//                       (cub::DeviceSegmentedSort; stability is irrelevant because rows of a
//                       tie group are interchangeable for every downstream count).
//   column_rank_kernel  compare_self + cumsum (:15-31, :250-251) -> dense ranks; tie-group
//                       sizes -> count_rank_tie sums (:103-118) in exact int64; the bit masks
//                       and the tied-row list (rows + dense group index) the pair kernel needs.
#include <cub/block/block_merge_sort.cuh>
#include <cub/device/device_segmented_sort.cuh>

#include <algorithm>
#include <cstdlib>

#include "icikt_internal.h"
#include "icikt_count.cuh"

namespace icikt {
namespace {

constexpr int RANK_THREADS = 512;

__device__ __forceinline__ unsigned long long order_key(double v) {
  if (v == 0.0) v = 0.0;  // -0.0 -> +0.0
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(256)
    build_keys_kernel(const double* __restrict__ data, long long ld, int n, int nstride, int wstride,
                      const double* __restrict__ global_na, int n_global_na, int na_inf,
                      unsigned long long* __restrict__ keys, uint16_t* __restrict__ vals,
                      uint32_t* __restrict__ nabits, const int col0) {
  const int col = blockIdx.y + col0;
  const int r = blockIdx.x * 256 + threadIdx.x;
  bool miss = false;
  if (r < n) {
    const double v = data[(size_t)col * ld + r];
    miss = (v != v) || (na_inf && isinf(v));
    for (int g = 0; g < n_global_na; ++g) miss = miss || (v == global_na[g]);
    keys[(size_t)col * nstride + r] = miss ? 0ull : order_key(v);
    vals[(size_t)col * nstride + r] = (uint16_t)r;
  }
  const uint32_t m = __ballot_sync(FULL, miss);
  if ((threadIdx.x & 31) == 0 && (r >> 5) < wstride) nabits[(size_t)col * wstride + (r >> 5)] = m;
}

// block-wide exclusive scan of one int per thread; returns the exclusive prefix, sets total
__device__ __forceinline__ int block_scan_excl(int v, int* warp_sums, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(FULL, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  int ws = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
  int wincl = ws;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(FULL, wincl, d);
    if (lane >= d) wincl += t;
  }
  total = __shfl_sync(FULL, wincl, 31);
  const int wexcl = __shfl_sync(FULL, wincl - ws, warp);
  __syncthreads();
  return wexcl + incl - v;
}

__device__ __forceinline__ long long block_sum_ll(long long v, long long* buf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
  if (lane == 0) buf[warp] = v;
  __syncthreads();
  long long t = (lane < (int)(blockDim.x >> 5)) ? buf[lane] : 0;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) t += __shfl_xor_sync(FULL, t, d);
  __syncthreads();
  return t;
}

// four sums with one pair of barriers; buf holds 4 * 32 values
__device__ __forceinline__ void block_sum_ll4(long long& a, long long& b, long long& c, long long& d, long long* buf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (int)(blockDim.x >> 5);
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    a += __shfl_xor_sync(FULL, a, s);
    b += __shfl_xor_sync(FULL, b, s);
    c += __shfl_xor_sync(FULL, c, s);
    d += __shfl_xor_sync(FULL, d, s);
  }
  if (lane == 0) {
    buf[warp] = a;
    buf[32 + warp] = b;
    buf[64 + warp] = c;
    buf[96 + warp] = d;
  }
  __syncthreads();
  a = lane < nw ? buf[lane] : 0;
  b = lane < nw ? buf[32 + lane] : 0;
  c = lane < nw ? buf[64 + lane] : 0;
  d = lane < nw ? buf[96 + lane] : 0;
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    a += __shfl_xor_sync(FULL, a, s);
    b += __shfl_xor_sync(FULL, b, s);
    c += __shfl_xor_sync(FULL, c, s);
    d += __shfl_xor_sync(FULL, d, s);
  }
  __syncthreads();
}

// Short columns (n <= 8192): the WHOLE per-column preprocessing in one kernel, one CTA per
// column, every sorted position held in registers (thread t owns positions t*ITEMS ..).
//   missing marking + order-preserving keys   (build_keys_kernel)
//   sort                                      (cub::BlockMergeSort in shared memory)
//   dense ranks, tie sums, first-group mask, tied-row list, statistics   (column_rank_kernel)
// Replaces six launches (~150 us for 100 columns of 5 000 rows, two thirds of it in
// cub::DeviceSegmentedSort) where the per-column work is a visible share of the whole job.
// Each phase is one pass over the registers plus one block scan.
struct KeyLess {
  __device__ __forceinline__ bool operator()(unsigned long long a, unsigned long long b) const { return a < b; }
};
template <int SORT_THREADS, int ITEMS>
__global__ void __launch_bounds__(SORT_THREADS)
    column_fused_kernel(const double* __restrict__ data, long long ld, int n, int nstride, int wstride,
                        const double* __restrict__ global_na, int n_global_na, int na_inf,
                        uint16_t* __restrict__ perm, uint16_t* __restrict__ rank, uint16_t* __restrict__ trow,
                        uint16_t* __restrict__ trun, uint16_t* __restrict__ tend, uint32_t* __restrict__ nabits,
                        uint32_t* __restrict__ firstbits, uint16_t* __restrict__ gstart, int gstride,
                        uint16_t* __restrict__ lgrp, ColStats* __restrict__ stats,
                        int32_t* __restrict__ max_tied, uint32_t* __restrict__ tord, const PipeConst pc,
                        const int large_tie, const int direct_budget, const int col0) {
  using Sort = cub::BlockMergeSort<unsigned long long, SORT_THREADS, ITEMS, uint16_t>;
  constexpr int CAP = SORT_THREADS * ITEMS;
  extern __shared__ __align__(16) unsigned char sort_smem[];
  __shared__ int warp_sums[32];
  __shared__ long long llbuf[128];
  __shared__ unsigned long long mnkey;
  __shared__ int n_large;
  __shared__ uint32_t descA[32], descB[32];
  typename Sort::TempStorage& temp = *reinterpret_cast<typename Sort::TempStorage*>(sort_smem);
  const int col = blockIdx.x + col0, tid = threadIdx.x;
  if (tid == 0) n_large = 0;
  unsigned long long keys[ITEMS];
  uint16_t vals[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {  // striped: a warp reads 32 consecutive rows
    const int r = i * SORT_THREADS + tid;
    bool miss = false;
    unsigned long long k = ~0ull;  // padding sorts last; no value maps to all-ones (NaN is missing)
    if (r < n) {
      const double v = data[(size_t)col * ld + r];
      miss = (v != v) || (na_inf && isinf(v));
      for (int g = 0; g < n_global_na; ++g) miss = miss || (v == global_na[g]);
      k = miss ? 0ull : order_key(v);
    }
    keys[i] = k;
    vals[i] = (uint16_t)r;
    const uint32_t m = __ballot_sync(FULL, miss);
    if ((tid & 31) == 0 && (r >> 5) < wstride) nabits[(size_t)col * wstride + (r >> 5)] = m;
  }
  Sort(temp).Sort(keys, vals, KeyLess());
  __syncthreads();  // the sort's shared memory is reused below
  unsigned long long* lastkey = reinterpret_cast<unsigned long long*>(sort_smem);           // [SORT_THREADS]
  uint16_t* gpos = reinterpret_cast<uint16_t*>(sort_smem + 8 * SORT_THREADS);                // [CAP + 2]
  uint32_t* bits = reinterpret_cast<uint32_t*>(sort_smem + 8 * SORT_THREADS + ((2 * (CAP + 2) + 15) & ~15));  // [CAP/32]
  uint32_t* whist = bits + ((CAP / 32 + 3) & ~3);                                            // [CAP]
  const int base = tid * ITEMS;  // blocked arrangement after the sort
  uint16_t* pm = perm + (size_t)col * nstride;
  uint16_t* rk = rank + (size_t)col * nstride;
  const int nwords = (n + 31) >> 5;

  int a_loc = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    if (base + i < n) pm[base + i] = vals[i];
    a_loc += (keys[i] == 0ull);
  }
  lastkey[tid] = keys[ITEMS - 1];
  int a;
  block_scan_excl(a_loc, warp_sums, a);  // missing rows sort first (key 0)
#pragma unroll
  for (int i = 0; i < ITEMS; ++i)
    if (base + i == a) mnkey = keys[i];
  for (int w = tid; w < nwords; w += SORT_THREADS) bits[w] = 0;
  __syncthreads();
  // NA substitute = min - 0.1 (src/kendallc.cpp:214-219); if that does not move the minimum in
  // fp64 the missing rows tie with it
  bool absorb = false;
  if (a > 0 && a < n) {
    const double mn = key_value(mnkey);
    absorb = (__dsub_rn(mn, 0.1) == mn);
  }
  // group starts (compare_self, :15-31) and dense ranks
  unsigned long long prev = tid > 0 ? lastkey[tid - 1] : 0ull;
  uint32_t fmask = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int t = base + i;
    const bool f = t < n && (t == 0 || (keys[i] != prev && !(t == a && absorb)));
    prev = keys[i];
    fmask |= (uint32_t)f << i;
  }
  int K;
  const int excl = block_scan_excl(__popc(fmask), warp_sums, K);
  {
    int r = excl - 1;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int t = base + i;
      if ((fmask >> i) & 1u) gpos[++r] = (uint16_t)t;
      if (t < n) rk[vals[i]] = (uint16_t)r;
    }
  }
  if (tid == 0) gpos[K] = (uint16_t)n;  // n <= 8192 here
  __syncthreads();
  for (int g = tid; g <= K; g += SORT_THREADS) gstart[(size_t)col * gstride + g] = gpos[g];
  // tie sums over group sizes (count_rank_tie, :103-118), exact int64
  long long s2 = 0, s3 = 0, s5 = 0, ntied = 0, lsq = 0;
  for (int g = tid + 1; g < K; g += SORT_THREADS) {
    const long long t = (long long)gpos[g + 1] - (long long)gpos[g];
    if (t >= large_tie) lsq += t * t;
  }
  lsq = block_sum_ll(lsq, llbuf);
  const int large_from = lsq > (long long)direct_budget * n ? large_tie : 0x7fffffff;
  for (int g = tid; g < K; g += SORT_THREADS) {
    const long long t = (long long)gpos[g + 1] - (long long)gpos[g];
    if (g > 0 && t >= large_from) {  // large tie group: also listed by (start position, size)
      const int k = atomicAdd(&n_large, 1);
      lgrp[(size_t)col * kLargeStride + 2 * k] = gpos[g];
      lgrp[(size_t)col * kLargeStride + 2 * k + 1] = (uint16_t)t;
    }
    if (g > 0 && t > 1) ntied += t;
    if (g == 0 && a > 0) continue;  // the NA group is kept apart for the local perspective
    s2 += t * (t - 1);
    s3 += t * (t - 1) * (t - 2);
    s5 += t * (t - 1) * (2 * t + 5);
  }
  block_sum_ll4(s2, s3, s5, ntied, llbuf);
  const int g0size = (K > 0) ? (int)gpos[1] : 0;
  const int first_run = g0size > 1 ? g0size : 0;
  // membership mask of the first group; rows of the other tied groups with their dense group index
  uint32_t tmask = 0, gmask = 0, lmask = 0;
  int after[ITEMS];  // rows of the item's group from the item to the group's end (tied items only)
  {
    int r = excl - 1;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int t = base + i;
      if ((fmask >> i) & 1u) ++r;
      if (t < first_run) atomicOr(&bits[vals[i] >> 5], 1u << (vals[i] & 31));
      if (t < n && t >= g0size) {
        const int sz = (int)gpos[r + 1] - (int)gpos[r];
        const bool tied = sz > 1;
        tmask |= (uint32_t)tied << i;
        lmask |= (uint32_t)(sz >= large_from) << i;
        after[i] = (int)gpos[r + 1] - t;
        gmask |= (uint32_t)(tied && (int)gpos[r] == t) << i;
      }
    }
  }
  int tot2;
  const int excl2 = block_scan_excl(__popc(tmask) | (__popc(gmask) << 16), warp_sums, tot2);
  {
    int pos = excl2 & 0xffff, gcount = excl2 >> 16;
    uint16_t* tr = trow + (size_t)col * nstride;
    uint16_t* tg = trun + (size_t)col * nstride;
    uint16_t* te = tend + (size_t)col * nstride;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if ((gmask >> i) & 1u) ++gcount;
      if ((tmask >> i) & 1u) {
        const bool large = (lmask >> i) & 1u;
        tr[pos] = vals[i];
        tg[pos] = (uint16_t)((gcount - 1) | (large ? kLargeFlag : 0u));
        te[pos] = (uint16_t)(large ? pos : pos + after[i]);  // tied rows of a group are consecutive in the list
        ++pos;
      }
    }
  }
  // ---- walk order for the pair kernel's direct comparison of small tie groups: a tied row is
  // compared with the rows behind it in its group (`walk` of them), so consecutive rows of a group
  // walk t-1, t-2, ... 0 steps and a warp taking them in list order idles half of the time.  Here the
  // tied rows are sorted by walk length, longest first (counting sort), and listed as
  // (list index << 16 | walk): the lanes of a warp then walk equally far.
  if ((tot2 & 0xffff) > 0) {  // block-uniform
    for (int w = tid; w < n; w += SORT_THREADS) whist[w] = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i)
      if ((tmask >> i) & 1u) {
        const uint32_t walk = ((lmask >> i) & 1u) ? 0u : (uint32_t)(after[i] - 1);
        after[i] = (int)(walk | (atomicAdd(&whist[walk], 1u) << 16));  // rank inside the walk class
      }
    __syncthreads();
    {  // class offsets, longest walk first
      const int per = (n + SORT_THREADS - 1) / SORT_THREADS;
      const int hi = n - 1 - tid * per, lo = max(hi - per + 1, 0);
      int sum = 0;
      for (int v = hi; v >= lo; --v) sum += (int)whist[v];
      int total;
      int run = block_scan_excl(sum, warp_sums, total);
      for (int v = hi; v >= lo; --v) {
        const int c = (int)whist[v];
        whist[v] = (uint32_t)run;
        run += c;
      }
    }
    __syncthreads();
    uint32_t* to = tord + (size_t)col * nstride;
    int pos = excl2 & 0xffff;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i)
      if ((tmask >> i) & 1u) {
        const uint32_t walk = (uint32_t)after[i] & 0xffffu;
        to[whist[walk] + ((uint32_t)after[i] >> 16)] = ((uint32_t)pos << 16) | walk;
        ++pos;
      }
  }
  __syncthreads();  // bits complete
  for (int w = tid; w < nwords; w += SORT_THREADS) firstbits[(size_t)col * wstride + w] = bits[w];
  if (tid == 0) {
    ColStats s;
    s.n_na = a;
    s.first_run = first_run;
    s.n_tied = (int)ntied;
    s.n_groups = K;
    int L = 1;
    while ((1 << L) < K) ++L;
    s.levels = L;
    s.g0extra = (a > 0) ? g0size - a : 0;
    s.flags = (absorb ? 1 : 0) | (n_large << 8);
    s.n_tgroups = tot2 >> 16;
    s.s2o = s2;
    s.s3o = s3;
    s.s5o = s5;
    s.cconst = 0;
    stats[col] = s;
    // what the pair kernel's launch tiers are selected by (on the device)
    if (n_large > 0) {
      atomicMax(max_tied + 0, 1);
      atomicMax(max_tied + 1, n_large);
    }
    atomicMax(max_tied + 2, K);
  }
  // ---- cconst: pass A (icikt_count.cuh) over the column's own sorted dense ranks.  They hold no
  // inversion, so what the bucket-free pass counts on them is exactly the per-column constant the
  // pair kernel subtracts.  One 8-key run per thread (kk = 1), two u16 buffers in the shared
  // memory the sort and the tables above no longer need.
  if (K < 2) return;  // constant or all-missing column: cconst stays 0 (block-uniform)
  __syncthreads();    // gpos / bits have been read by everyone
  {
    const int nw = (n + 255) >> 8;  // warps whose 256 positions hold keys; the others only keep the barriers
    const int capc = nw << 8;
    const int L = max(1, 32 - __clz(K - 1));
    const uint32_t bufA = smem_addr(sort_smem), bufB = bufA + 2u * (uint32_t)capc;
    int r = excl - 1;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if ((fmask >> i) & 1u) ++r;
      if (base + i < n) Mem<false>::st16(bufA + 2u * (uint32_t)(base + i), (uint32_t)r);
    }
    const uint32_t pad = (1u << L) - 1u;
    for (int q = n + tid; q < capc; q += SORT_THREADS) Mem<false>::st16(bufA + 2u * (uint32_t)q, pad);
    __syncthreads();
    unsigned long long acc = 0;
    if ((tid >> 5) < nw) {
      count_pass<false>(bufA, bufB, 1, nw, L, descA, descB, tid & 31, tid >> 5, pc, acc);
    } else {  // count_pass has two block-wide barriers per two-bit level
      for (int lv = (L - 1) & ~1; lv >= 0; lv -= 2) {
        __syncthreads();
        __syncthreads();
      }
    }
    const long long total = block_sum_ll((long long)acc, llbuf);
    if (tid == 0) stats[col].cconst = (uint64_t)total;
  }
}

template <int SORT_THREADS, int ITEMS>
int launch_column_fused(const double* d_data, int64_t ld, const double* d_global_na, int n_global_na, int na_inf,
                        ColumnTables& tab, cudaStream_t stream, int large_tie, int direct_budget, int col0,
                        int ncols) {
  using Sort = cub::BlockMergeSort<unsigned long long, SORT_THREADS, ITEMS, uint16_t>;
  constexpr int CAP = SORT_THREADS * ITEMS;
  const size_t post = 8 * SORT_THREADS + ((2 * (CAP + 2) + 15) & ~15) + 4 * ((CAP / 32 + 3) & ~3) + 4 * (size_t)CAP;
  const size_t pass_a = 2 * 2 * (size_t)SORT_THREADS * 8;  // two u16 buffers of 8 keys per thread
  const size_t smem = std::max(std::max(sizeof(typename Sort::TempStorage), post), pass_a);
  PipeConst pc;
  pc.one = 1u;
  pc.two = 2u;
  pc.c64k = 65536u;
  auto kern = column_fused_kernel<SORT_THREADS, ITEMS>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
  kern<<<(unsigned)ncols, SORT_THREADS, smem, stream>>>(d_data, ld, (int)tab.n, (int)tab.nstride, (int)tab.wstride,
                                                        d_global_na, n_global_na, na_inf, tab.perm, tab.rank,
                                                        tab.trow, tab.trun, tab.tend, tab.nabits, tab.firstbits, tab.gstart,
                                                        (int)tab.gstride, tab.lgrp, tab.stats, tab.max_tied, tab.tord, pc, large_tie, direct_budget, col0);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

__global__ void __launch_bounds__(RANK_THREADS)
    column_rank_kernel(const unsigned long long* __restrict__ skeys, int n, int nstride, int wstride,
                       uint16_t* __restrict__ perm, uint16_t* __restrict__ rank,
                       uint16_t* __restrict__ trow, uint16_t* __restrict__ trun, uint16_t* __restrict__ tend,
                       uint32_t* __restrict__ firstbits,
                       uint32_t* __restrict__ gpos_all, uint16_t* __restrict__ gstart_tab, int gstride,
                       uint16_t* __restrict__ lgrp, ColStats* __restrict__ stats,
                       int32_t* __restrict__ max_tied, uint32_t* __restrict__ tord,
                       uint16_t* __restrict__ wrank_all, const int large_tie, const int direct_budget,
                       const int col0) {
  // rows compared directly walk less than 2048 steps: groups below large_tie (<= 2048) rows, or larger
  // ones whose squares sum to at most direct_budget (<= 63) * n (launch_columns clamps both)
  __shared__ int n_large;
  __shared__ int warp_sums[32];
  __shared__ long long llbuf[128];
  __shared__ uint32_t bits[2048];
  __shared__ uint32_t whist[2048];
  const int col = blockIdx.x + col0;
  const int tid = threadIdx.x;
  const unsigned long long* sk = skeys + (size_t)col * nstride;
  uint16_t* pm = perm + (size_t)col * nstride;
  uint16_t* rk = rank + (size_t)col * nstride;
  uint16_t* tr = trow + (size_t)col * nstride;
  uint16_t* tg = trun + (size_t)col * nstride;
  uint16_t* te = tend + (size_t)col * nstride;
  uint32_t* gpos = gpos_all + (size_t)col * (nstride + 64);
  const int n32 = (n + 31) & ~31;
  const int nwords = n32 >> 5;
  if (tid == 0) n_large = 0;

  // missing rows sort first (key 0)
  long long a_ll = 0;
  for (int t = tid; t < n; t += RANK_THREADS) a_ll += (sk[t] == 0ull);
  const int a = (int)block_sum_ll(a_ll, llbuf);
  // NA substitute = min - 0.1 (src/kendallc.cpp:214-219); if that does not move the minimum in
  // fp64 (|min| huge or -Inf) the missing rows tie with it
  bool absorb = false;
  if (a > 0 && a < n) {
    const double mn = key_value(sk[a]);
    absorb = (__dsub_rn(mn, 0.1) == mn);
  }

  // dense ranks in sorted order
  int carry = 0;
  for (int t0 = 0; t0 < n32; t0 += RANK_THREADS) {
    const int t = t0 + tid;
    int flag = 0;
    if (t < n) flag = (t == 0) || ((sk[t] != sk[t - 1]) && !(t == a && absorb));
    int total;
    const int excl = block_scan_excl(flag, warp_sums, total);
    if (t < n) {
      const int r = carry + excl + flag - 1;
      rk[pm[t]] = (uint16_t)r;
      if (flag) gpos[r] = (uint32_t)t;
    }
    carry += total;
  }
  const int K = carry;
  if (tid == 0) gpos[K] = (uint32_t)n;
  __syncthreads();
  for (int g = tid; g <= K; g += RANK_THREADS) gstart_tab[(size_t)col * gstride + g] = (uint16_t)gpos[g];  // n <= 65535

  // tie sums over group sizes (count_rank_tie, src/kendallc.cpp:103-118), exact int64
  long long s2 = 0, s3 = 0, s5 = 0, ntied = 0, lsq = 0;
  for (int g = tid + 1; g < K; g += RANK_THREADS) {
    const long long t = (long long)gpos[g + 1] - (long long)gpos[g];
    if (t >= large_tie) lsq += t * t;
  }
  lsq = block_sum_ll(lsq, llbuf);
  const long long large_from = lsq > (long long)direct_budget * n ? large_tie : 0x7fffffff;
  for (int g = tid; g < K; g += RANK_THREADS) {
    const long long t = (long long)gpos[g + 1] - (long long)gpos[g];
    if (g > 0 && t >= large_from) {  // large tie group: also listed by (start position, size)
      const int k = atomicAdd(&n_large, 1);
      lgrp[(size_t)col * kLargeStride + 2 * k] = (uint16_t)gpos[g];
      lgrp[(size_t)col * kLargeStride + 2 * k + 1] = (uint16_t)t;
    }
    if (g > 0 && t > 1) ntied += t;
    if (g == 0 && a > 0) continue;  // the NA group is kept apart for the local perspective
    s2 += t * (t - 1);
    s3 += t * (t - 1) * (t - 2);
    s5 += t * (t - 1) * (2 * t + 5);
  }
  block_sum_ll4(s2, s3, s5, ntied, llbuf);
  const int g0size = (K > 0) ? (int)gpos[1] : 0;
  const int first_run = g0size > 1 ? g0size : 0;

  // membership mask of the first group
  for (int w = tid; w < nwords; w += RANK_THREADS) bits[w] = 0;
  __syncthreads();
  for (int t = tid; t < first_run; t += RANK_THREADS) {
    const uint32_t row = pm[t];
    atomicOr(&bits[row >> 5], 1u << (row & 31));
  }
  __syncthreads();
  for (int w = tid; w < nwords; w += RANK_THREADS) firstbits[(size_t)col * wstride + w] = bits[w];

  // rows of the other tied groups, in sorted order, with a dense index of their group
  carry = 0;
  int gcarry = 0;
  for (int t0 = 0; t0 < n32; t0 += RANK_THREADS) {
    const int t = t0 + tid;
    int flag = 0, r = 0, gstart = 0, large = 0, after = 0;
    uint16_t row = 0;
    if (t < n && t >= g0size) {
      row = pm[t];
      r = rk[row];
      const uint32_t sz = gpos[r + 1] - gpos[r];
      flag = sz > 1;
      large = (long long)sz >= large_from;
      after = (int)gpos[r + 1] - t;
      gstart = flag && (gpos[r] == (uint32_t)t);
    }
    int total, gtotal;
    const int excl = block_scan_excl(flag, warp_sums, total);
    const int gincl = block_scan_excl(gstart, warp_sums, gtotal) + gstart;
    if (flag) {
      tr[carry + excl] = row;
      tg[carry + excl] = (uint16_t)((gcarry + gincl - 1) | (large ? kLargeFlag : 0u));
      te[carry + excl] = (uint16_t)(large ? carry + excl : carry + excl + after);
    }
    carry += total;
    gcarry += gtotal;
  }

  // walk order of the tied rows for the pair kernel's direct comparison (see column_fused_kernel):
  // counting sort by walk length, longest first; the rank inside a class is parked in the sort's
  // value buffer, which is free by now
  if (carry > 0) {  // block-uniform
    const int m = carry;
    uint16_t* wr = wrank_all + (size_t)col * nstride;
    uint32_t* to = tord + (size_t)col * nstride;
    for (int w = tid; w < 2048; w += RANK_THREADS) whist[w] = 0;
    __syncthreads();  // also: te[] above was written by other threads
    for (int k = tid; k < m; k += RANK_THREADS) {
      const int e = te[k];
      const int walk = e > k ? e - k - 1 : 0;
      wr[k] = (uint16_t)atomicAdd(&whist[walk], 1u);
    }
    __syncthreads();
    {
      constexpr int PER = 2048 / RANK_THREADS;
      const int hi = 2047 - tid * PER;
      int sum = 0;
#pragma unroll
      for (int q = 0; q < PER; ++q) sum += (int)whist[hi - q];
      int total;
      int run = block_scan_excl(sum, warp_sums, total);
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int c = (int)whist[hi - q];
        whist[hi - q] = (uint32_t)run;
        run += c;
      }
    }
    __syncthreads();
    for (int k = tid; k < m; k += RANK_THREADS) {
      const int e = te[k];
      const uint32_t walk = e > k ? (uint32_t)(e - k - 1) : 0u;
      to[whist[walk] + wr[k]] = ((uint32_t)k << 16) | walk;
    }
  }

  if (tid == 0) {
    ColStats s;
    s.n_na = a;
    s.first_run = first_run;
    s.n_tied = (int)ntied;
    s.n_groups = K;
    int L = 1;
    while ((1 << L) < K) ++L;
    s.levels = L;
    s.g0extra = (a > 0) ? g0size - a : 0;
    s.flags = (absorb ? 1 : 0) | (n_large << 8);
    s.n_tgroups = gcarry;
    s.s2o = s2;
    s.s3o = s3;
    s.s5o = s5;
    s.cconst = 0;
    stats[col] = s;
    if (n_large > 0) {
      atomicMax(max_tied + 0, 1);
      atomicMax(max_tied + 1, n_large);
    }
    atomicMax(max_tied + 2, K);
  }
}

__global__ void seg_offsets_kernel(long long* begin, long long* end, int c0, int c1, long long nstride, long long n) {
  const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (c < c1) {
    begin[c] = c * nstride;
    end[c] = c * nstride + n;
  }
}

// tier maxima from the statistics of all columns (what the column kernels raise with atomicMax in a
// full run): [0] any large tie group, [1] most large groups, [2] most distinct values of a column
__global__ void max_tied_kernel(const ColStats* __restrict__ stats, int C, int32_t* __restrict__ max_tied) {
  int nl = 0, k = 0;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    nl = max(nl, stats[c].flags >> 8);
    k = max(k, stats[c].n_groups);
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    nl = max(nl, __shfl_xor_sync(FULL, nl, d));
    k = max(k, __shfl_xor_sync(FULL, k, d));
  }
  if ((threadIdx.x & 31) == 0) {
    if (nl > 0) {
      atomicMax(max_tied + 0, 1);
      atomicMax(max_tied + 1, nl);
    }
    atomicMax(max_tied + 2, k);
  }
}

}  // namespace

int launch_max_tied(ColumnTables& tab, cudaStream_t stream) {
  if (cudaMemsetAsync(tab.max_tied, 0, 4 * sizeof(int32_t), stream) != cudaSuccess) return -1;
  const int C = (int)tab.C;
  max_tied_kernel<<<std::min(64, (C + 255) / 256), 256, 0, stream>>>(tab.stats, C, tab.max_tied);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

size_t columns_cub_bytes(int64_t n, int64_t C, int64_t nstride) {
  size_t bytes = 0;
  cub::DeviceSegmentedSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr,
                                      (unsigned long long*)nullptr, (const uint16_t*)nullptr,
                                      (uint16_t*)nullptr, (long long)(nstride * C), (long long)C,
                                      (const long long*)nullptr, (const long long*)nullptr);
  (void)n;
  return bytes;
}

bool columns_fused(int64_t n) { return n <= 8192 && !getenv("ICIKT_NO_FUSED_COLUMNS"); }

int launch_columns(const double* d_data, int64_t ld, const double* d_global_na, int n_global_na,
                   int na_inf, ColumnTables& tab, ColumnWork& wk, const TiledShape& sh,
                   unsigned char* scratch, cudaStream_t stream, int64_t col_lo, int64_t col_hi) {
  const int n = (int)tab.n;
  const int col0 = (int)col_lo, C = (int)(col_hi - col_lo);  // the columns of this call
  if (C <= 0) return 0;
  const int nstride = (int)tab.nstride, wstride = (int)tab.wstride;
  int launches = 0;
  // tie-group thresholds (icikt_common.cuh); the environment overrides are for tuning sweeps
  int large_tie = kLargeTie, direct_budget = kDirectBudget;
  if (const char* e = getenv("ICIKT_LARGE_TIE")) large_tie = std::min(2048, std::max(2, atoi(e)));
  if (const char* e = getenv("ICIKT_DIRECT_BUDGET")) direct_budget = std::min(63, std::max(0, atoi(e)));
  // the bit arrays are written in full (words below n32/32) by the kernels below; the padding
  // words up to wstride were zeroed once when the plan was created
  // short columns: 512 threads (more CTAs per SM when there are many columns), else 1024
  // full runs reset the tier maxima here and the kernels below raise them; a partial run (sharded
  // preprocessing) is followed by launch_max_tied once the statistics of every column are in place
  if (col0 == 0 && C == (int)tab.C && cudaMemsetAsync(tab.max_tied, 0, 4 * sizeof(int32_t), stream) != cudaSuccess) return -1;
  if (columns_fused(n)) {
    int l;
#define ICIKT_FUSED(T, I) \
  l = launch_column_fused<T, I>(d_data, ld, d_global_na, n_global_na, na_inf, tab, stream, large_tie, direct_budget, col0, C)
    if (n <= 512) ICIKT_FUSED(512, 1);
    else if (n <= 1024) ICIKT_FUSED(512, 2);
    else if (n <= 2048) ICIKT_FUSED(512, 4);
    else if (n <= 3072) ICIKT_FUSED(512, 6);
    else if (n <= 4096) ICIKT_FUSED(1024, 4);
    else if (n <= 5120) ICIKT_FUSED(1024, 5);
    else if (n <= 6144) ICIKT_FUSED(1024, 6);
    else ICIKT_FUSED(1024, 8);
#undef ICIKT_FUSED
    if (l < 0) return -1;
    return launches + l;  // the fused kernel computes the pass-A constants itself
  } else {
    seg_offsets_kernel<<<(C + 255) / 256, 256, 0, stream>>>(wk.seg_begin, wk.seg_end, col0, col0 + C, nstride, n);
    ++launches;
    const int n32 = (n + 31) & ~31;
    dim3 grid((n32 + 255) / 256, C);
    build_keys_kernel<<<grid, 256, 0, stream>>>(d_data, ld, n, nstride, wstride, d_global_na, n_global_na,
                                                na_inf, wk.keys_in, wk.vals_in, tab.nabits, col0);
    ++launches;
    if (cudaGetLastError() != cudaSuccess) return -1;
    size_t bytes = wk.cub_bytes;
    if (cub::DeviceSegmentedSort::SortPairs(wk.cub_temp, bytes, (const unsigned long long*)wk.keys_in,
                                            wk.keys_out, (const uint16_t*)wk.vals_in, tab.perm,
                                            (long long)nstride * tab.C, (long long)C,
                                            (const long long*)wk.seg_begin + col0, (const long long*)wk.seg_end + col0,
                                            stream) != cudaSuccess)
      return -1;
    launches += 3;  // cub partitions the segments into size classes: up to three sort kernels
    column_rank_kernel<<<C, RANK_THREADS, 0, stream>>>(wk.keys_out, n, nstride, wstride, tab.perm, tab.rank,
                                                       tab.trow, tab.trun, tab.tend, tab.firstbits,
                                                       wk.gpos, tab.gstart, (int)tab.gstride, tab.lgrp, tab.stats, tab.max_tied,
                                                       tab.tord, wk.vals_in, large_tie, direct_budget, col0);
    ++launches;
    if (cudaGetLastError() != cudaSuccess) return -1;
  }
  const int cl = launch_column_consts(tab, sh, scratch, stream, col_lo, col_hi);
  if (cl < 0) return -1;
  return launches + cl;
}

}  // namespace icikt
