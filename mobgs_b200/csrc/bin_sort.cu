// K3': tile binning + per-tile (depth, index) radix sort.
//
// gsplat builds one global list of (tile | depth) 64-bit keys and runs a device-wide radix sort
// over all intersections (4-6 passes over 12 B/entry through HBM).  Here the tile id never enters
// the key: a counting pass + prefix sum gives every (sub-frame, tile) its own segment, entries
// are scattered into their segment, and one CTA sorts each segment on the 64-bit key
// (depth_bits << 32 | gaussian_index) — in shared memory when the segment fits, ping-ponging
// through L2-resident global buffers when it does not.  Digits on which all keys of a segment
// agree are skipped, so a typical segment takes 5-7 byte passes that never leave the SM.
// The resulting order — ascending depth, ties by ascending Gaussian index — is exactly what
// gsplat's stable sort of (tile, depth) keys emitted in Gaussian order produces.
#include "common.cuh"
#include "tile_bin.cuh"

namespace mobgs {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kRankSortMax = 768;    // segments up to this size use the O(n^2 / 256) rank sort

__global__ void __launch_bounds__(256) tile_count_kernel(const __grid_constant__ MobgsTileCount a, int tiles_x, int tiles_y) {
  const size_t iv = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (iv >= (size_t)a.K * a.N) return;
  const int k = (int)(iv / a.N);
  const int gid = (int)(iv - (size_t)k * a.N);
  if (gid < a.lists.g_begin[k] || gid >= a.lists.g_end[k]) return;
  const size_t i = (size_t)a.lists.rec_k[k] * a.N + gid;     // physical record
  const int radius = a.radii[i];
  if (radius <= 0) return;
  const GaussGeom g = load_geom(a.records + i * kRecFloats);
  int* counts = a.tile_counts + (size_t)k * tiles_x * tiles_y;
  for_each_tile(g, radius, a.width, a.height, tiles_x, tiles_y, a.tight,
                [&](int tile) { atomicAdd(counts + tile, 1); });
}

// counting pass that records (segment, slot, key) entries: see bin_count_record (tile_bin.cuh)
__global__ void __launch_bounds__(256) tile_count_entries_kernel(const __grid_constant__ MobgsTileCount a, int tiles_x, int tiles_y) {
  const size_t iv = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool live = iv < (size_t)a.K * a.N;
  int k = 0, gid = 0, radius = 0;
  size_t i = 0;
  if (live) {
    k = (int)(iv / a.N);
    gid = (int)(iv - (size_t)k * a.N);
    live = gid >= a.lists.g_begin[k] && gid < a.lists.g_end[k];
  }
  if (live) {
    i = (size_t)a.lists.rec_k[k] * a.N + gid;     // physical record
    radius = a.radii[i];
    live = radius > 0;
  }
  GaussGeom g = {0.f, 0.f, 0.f, 1.f, 0.f, 1.f};
  float depth = 0.f;
  if (live) { g = load_geom(a.records + i * kRecFloats); depth = a.depths[i]; }
  const BinTarget b = {a.tile_counts, a.entries, a.entry_capacity, a.entry_cursor, a.width, a.height, tiles_x, tiles_y, a.tight};
  bin_count_record(b, k, gid, live, g, radius, depth);
}

// emit pass over the entries of tile_count_entries_kernel: keys[tile_offsets[seg] + slot] = depth_bits << 32 | gid
__global__ void __launch_bounds__(256) tile_scatter_kernel(const __grid_constant__ MobgsTileSort a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = min((int64_t)*a.n_entries, a.entry_capacity);
  if (i >= n) return;
  const uint4 e = __ldcs(reinterpret_cast<const uint4*>(a.entries) + i);
  const int64_t dst = (int64_t)a.tile_offsets[e.x] + (int)e.y;
  if (dst < a.capacity) a.keys[dst] = ((uint64_t)e.w << 32) | e.z;
}

// exclusive prefix sum over n ints by a single CTA; out[n] = total.  Thread t owns the contiguous chunk
// [t c, (t + 1) c), c = ceil(n / 1024) rounded up to a multiple of 4 (int4 accesses): one pass sums it, a block scan
// gives its base, a second pass writes the prefixes — two barriers instead of four per 1024 elements.
__global__ void __launch_bounds__(1024) scan_kernel(const int* __restrict__ in, int* __restrict__ out, int n) {
  __shared__ int warp_tot[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = (((n + 1023) / 1024) + 3) & ~3;
  const int beg = min(n, (int)threadIdx.x * c), end = min(n, beg + c);
  const bool vec = (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  int sum = 0;
  if (vec) {
    int j = beg;
    for (; j + 4 <= end; j += 4) { const int4 v = *reinterpret_cast<const int4*>(in + j); sum += v.x + v.y + v.z + v.w; }
    for (; j < end; ++j) sum += in[j];
  } else {
    for (int j = beg; j < end; ++j) sum += in[j];
  }
  int s = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
  if (lane == 31) warp_tot[warp] = s;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    warp_tot[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  int run = (warp ? warp_tot[warp - 1] : 0) + s - sum;   // exclusive prefix of this thread's chunk
  if (vec) {
    int j = beg;
    for (; j + 4 <= end; j += 4) {
      const int4 v = *reinterpret_cast<const int4*>(in + j);
      int4 o;
      o.x = run; o.y = o.x + v.x; o.z = o.y + v.y; o.w = o.z + v.z;
      run = o.w + v.w;
      *reinterpret_cast<int4*>(out + j) = o;
    }
    for (; j < end; ++j) { const int v = in[j]; out[j] = run; run += v; }
  } else {
    for (int j = beg; j < end; ++j) { const int v = in[j]; out[j] = run; run += v; }
  }
  if (threadIdx.x == 1023) out[n] = warp_tot[31];
}

__global__ void __launch_bounds__(256) tile_emit_kernel(const __grid_constant__ MobgsTileSort a, int tiles_x, int tiles_y) {
  const size_t iv = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (iv >= (size_t)a.K * a.N) return;
  const int k = (int)(iv / a.N);
  const int gid = (int)(iv - (size_t)k * a.N);
  if (gid < a.lists.g_begin[k] || gid >= a.lists.g_end[k]) return;
  const size_t i = (size_t)a.lists.rec_k[k] * a.N + gid;
  const int radius = a.radii[i];
  if (radius <= 0) return;
  const GaussGeom g = load_geom(a.records + i * kRecFloats);
  const size_t tbase = (size_t)k * tiles_x * tiles_y;
  const uint64_t key = ((uint64_t)__float_as_uint(a.depths[i]) << 32) | (uint32_t)gid;
  for_each_tile(g, radius, a.width, a.height, tiles_x, tiles_y, a.tight, [&](int tile) {
    const int slot = a.tile_offsets[tbase + tile] + atomicAdd(a.tile_cursor + tbase + tile, 1);
    if ((int64_t)slot < a.capacity) a.keys[slot] = key;
  });
}

// ------------------------------------------------------------------------------------------
// one CTA sorts one (sub-frame, tile) segment
// ------------------------------------------------------------------------------------------
// Small segments (n <= kRankSortMax, the common case): all-pairs rank sort out of 6 KB of static shared
// memory — keys are unique (they embed the Gaussian index), so rank = #{keys smaller} is the final
// position.  A kernel of its own so that its occupancy is not capped by the radix path's 72 KB buffers.
__global__ void __launch_bounds__(kSortThreads) tile_rank_sort_kernel(MobgsTileSort a) {
  __shared__ __align__(16) uint64_t buf[kRankSortMax + 1];
  const int seg = blockIdx.x;
  const int beg = a.tile_offsets[seg];
  int n = a.tile_offsets[seg + 1] - beg;
  if ((int64_t)beg + n > a.capacity) n = (int)max((int64_t)0, a.capacity - beg);
  if (n <= 0 || n > kRankSortMax) return;
  const uint64_t* gkeys = a.keys + beg;
  if (n == 1) {
    if (threadIdx.x == 0) a.sorted_ids[beg] = (int)(gkeys[0] & 0xffffffffu);
    return;
  }
  for (int i = threadIdx.x; i < n; i += kSortThreads) buf[i] = gkeys[i];
  if (threadIdx.x == 0) buf[n] = ~0ull;          // sentinel: never smaller than a key (pairs are read two at a time)
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kSortThreads) {
    const uint64_t key = buf[i];
    int rank = 0;
#pragma unroll 4
    for (int j = 0; j < n; j += 2) {
      const ulonglong2 kk = *reinterpret_cast<const ulonglong2*>(buf + j);   // one LDS.128 = two keys
      rank += (kk.x < key) + (kk.y < key);
    }
    a.sorted_ids[beg + rank] = (int)(key & 0xffffffffu);
  }
}

// Bucketed rank sort (the default for n <= kRankSortMax): the all-pairs count above costs n / 2 iterations per key.  Here
// the keys are first split into kBuckets ranges of their depth bits — bucket = (depth_bits - min) >> shift, monotone in
// the key, so buckets are already in final order — with one shared-memory atomic per key (histogram position), an
// exclusive scan over the buckets, and a scatter into bucket order; each key then ranks itself against its own bucket
// only (a handful of LDS.64 instead of n / 2 LDS.128).  Threads are assigned to keys in bucket order, so a warp's lanes
// loop over the same one or two buckets.  Depth ties share a bucket and are ordered by the index half of the key, as
// before; if every key has the same depth the kernel degenerates to the all-pairs count.
#ifndef MOBGS_RANK_SORT_BUCKETS
#define MOBGS_RANK_SORT_BUCKETS 1
#endif
constexpr int kBuckets = 64;
struct SmallSortSmem {
  uint64_t buf[kRankSortMax];
  int hist[kBuckets], start[kBuckets + 1];
  uint32_t wmin[kSortWarps], wmax[kSortWarps];
};

// sorts gkeys[0, n), 2 <= n <= kRankSortMax, and writes the Gaussian indices to out[0, n); called by all threads of the CTA
__device__ __forceinline__ void small_sort(SmallSortSmem& sm, const uint64_t* gkeys, int n, int32_t* out) {
  constexpr int kPer = (kRankSortMax + kSortThreads - 1) / kSortThreads;   // keys per thread (3)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t key[kPer];
  uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int i = threadIdx.x + j * kSortThreads;
    key[j] = i < n ? gkeys[i] : 0ull;
    if (i < n) { const uint32_t d = (uint32_t)(key[j] >> 32); lo = min(lo, d); hi = max(hi, d); }
  }
  if (threadIdx.x < kBuckets) sm.hist[threadIdx.x] = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { sm.wmin[warp] = lo; sm.wmax[warp] = hi; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) { lo = min(lo, sm.wmin[w]); hi = max(hi, sm.wmax[w]); }
  const uint32_t range = hi - lo;
  const int shift = range < (uint32_t)kBuckets ? 0 : (32 - __clz(range)) - 6;   // (range >> shift) < kBuckets = 2^6
  int slot[kPer], bkt[kPer];
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int i = threadIdx.x + j * kSortThreads;
    bkt[j] = (int)(((uint32_t)(key[j] >> 32) - lo) >> shift);
    slot[j] = i < n ? atomicAdd(&sm.hist[bkt[j]], 1) : 0;
  }
  __syncthreads();
  if (warp == 0) {          // exclusive scan over the 64 buckets, two per lane
    const int c0 = sm.hist[2 * lane], c1 = sm.hist[2 * lane + 1];
    int s = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
    sm.start[2 * lane] = s - c0 - c1;
    sm.start[2 * lane + 1] = s - c1;
    if (lane == 31) sm.start[kBuckets] = s;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int i = threadIdx.x + j * kSortThreads;
    if (i < n) sm.buf[sm.start[bkt[j]] + slot[j]] = key[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kSortThreads) {
    const uint64_t k = sm.buf[i];
    const int b = (int)(((uint32_t)(k >> 32) - lo) >> shift);
    const int b0 = sm.start[b], b1 = sm.start[b + 1];
    int rank = b0;
    for (int q = b0; q < b1; ++q) rank += sm.buf[q] < k;
    out[rank] = (int)(k & 0xffffffffu);
  }
  __syncthreads();          // the buffers may be reused by the caller's next call
}

__global__ void __launch_bounds__(kSortThreads) tile_bucket_sort_kernel(MobgsTileSort a) {
  __shared__ __align__(16) SmallSortSmem sm;
  const int seg = blockIdx.x;
  const int beg = a.tile_offsets[seg];
  int n = a.tile_offsets[seg + 1] - beg;
  if ((int64_t)beg + n > a.capacity) n = (int)max((int64_t)0, a.capacity - beg);
  if (n <= 0 || n > kRankSortMax) return;
  if (n == 1) {
    if (threadIdx.x == 0) a.sorted_ids[beg] = (int)(a.keys[beg] & 0xffffffffu);
    return;
  }
  small_sort(sm, a.keys + beg, n, a.sorted_ids + beg);
}

// Large segments (n > kRankSortMax): MSD partition on the 64-bit key + small sorts.  A range is split into 256 (n <= 4096)
// or 2048 buckets of (key - min) >> shift — monotone in the key, so the buckets are already in final order — by three
// streaming passes (min / max, histogram, scatter into the other key buffer).  Keys in buckets of at most kDirectMax
// entries then rank themselves inside their bucket straight from the partitioned array (a short loop of cached loads,
// threads in bucket order); buckets up to kRankSortMax take the shared-memory small sort; larger ones (skewed depth
// distributions) are pushed on a small stack and partitioned again with their own, narrower key range — keys are unique,
// so every level strictly narrows the range and the recursion ends.  Persistent CTAs take segments from an atomic
// counter.  (This replaced an LSD radix sort — 5-7 histogram / scatter passes per segment, ping-ponging through global
// memory above 4096 keys — that took 39 ms on the heavy-footprint workload's 242 M entries.)
constexpr int kMsdBucketsMax = 2048;
constexpr int kDirectMax = 64;        // bucket sizes ranked directly from the partitioned array
constexpr int kMsdStack = 256;        // pending ranges per CTA
constexpr int kMsdListMax = 512;      // buckets of one partition that need their own sort

struct MsdSmem {
  SmallSortSmem small;
  int hist[kMsdBucketsMax];
  int start[kMsdBucketsMax + 1];
  unsigned long long rmin[kSortWarps], rmax[kSortWarps];
  int warp_tot[kSortWarps];
  int stk_off[kMsdStack], stk_len[kMsdStack], stk_lvl[kMsdStack];
  int big_list[kMsdListMax];
  int stk_n, big_n, seg;
};

// all-pairs rank straight from global memory: O(m^2 / threads); only when the range stack is full
__device__ __forceinline__ void slow_sort(const uint64_t* keys, int m, int32_t* out) {
  for (int i = threadIdx.x; i < m; i += kSortThreads) {
    const uint64_t k = keys[i];
    int rank = 0;
    for (int q = 0; q < m; ++q) rank += keys[q] < k;
    out[rank] = (int)(k & 0xffffffffu);
  }
}

// One partition step of the range [off, off + len) of a segment whose keys currently live in A (B = the other buffer;
// both already offset to the segment's begin).  Called by all threads of the CTA.
__device__ void msd_partition(MsdSmem& sm, const uint64_t* A, uint64_t* B, int32_t* out, int off, int len, int lvl) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb_bits = len > 4096 ? 11 : 8, nb = 1 << nb_bits;
  // 1. key range
  unsigned long long lo = ~0ull, hi = 0ull;
  for (int i = threadIdx.x; i < len; i += kSortThreads) { const unsigned long long k = A[off + i]; lo = min(lo, k); hi = max(hi, k); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { sm.rmin[warp] = lo; sm.rmax[warp] = hi; }
  for (int b = threadIdx.x; b < nb; b += kSortThreads) sm.hist[b] = 0;
  if (threadIdx.x == 0) sm.big_n = 0;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) { lo = min(lo, sm.rmin[w]); hi = max(hi, sm.rmax[w]); }
  const unsigned long long range = hi - lo;
  if (range == 0ull) {
    // cannot happen with real keys (they embed the Gaussian index, so they are unique); the unwritten keys of an
    // overflowed speculative attempt can all be equal, and that attempt's lists only have to be harmless
    for (int i = threadIdx.x; i < len; i += kSortThreads) out[off + i] = (int)(A[off + i] & 0xffffffffu);
    __syncthreads();
    return;
  }
  const int bits = 64 - __clzll((long long)range);
  const int shift = bits > nb_bits ? bits - nb_bits : 0;          // (range >> shift) < nb
  // 2. histogram
  for (int i = threadIdx.x; i < len; i += kSortThreads) atomicAdd(&sm.hist[(int)((A[off + i] - lo) >> shift)], 1);
  __syncthreads();
  // 3. exclusive scan over nb buckets (nb / 256 consecutive buckets per thread); large buckets go on big_list
  {
    const int per = nb / kSortThreads;                             // 1 or 8
    const int b0 = threadIdx.x * per;
    int sum = 0;
    for (int j = 0; j < per; ++j) sum += sm.hist[b0 + j];
    int sc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
    if (lane == 31) sm.warp_tot[warp] = sc;
    __syncthreads();
    int base = sc - sum;
    for (int w = 0; w < warp; ++w) base += sm.warp_tot[w];
    for (int j = 0; j < per; ++j) {
      const int c = sm.hist[b0 + j];
      sm.start[b0 + j] = base;
      sm.hist[b0 + j] = 0;                                          // becomes the scatter cursor
      if (c > kDirectMax) {
        const int slot = atomicAdd(&sm.big_n, 1);
        if (slot < kMsdListMax) sm.big_list[slot] = b0 + j;
      }
      base += c;
    }
    if (threadIdx.x == kSortThreads - 1) sm.start[nb] = base;
  }
  __syncthreads();
  // 4. scatter into bucket order
  for (int i = threadIdx.x; i < len; i += kSortThreads) {
    const unsigned long long k = A[off + i];
    const int b = (int)((k - lo) >> shift);
    B[off + sm.start[b] + atomicAdd(&sm.hist[b], 1)] = k;
  }
  __syncthreads();
  // 5. small buckets: every key ranks itself inside its bucket
  for (int i = threadIdx.x; i < len; i += kSortThreads) {
    const unsigned long long k = B[off + i];
    const int b = (int)((k - lo) >> shift);
    const int b0 = sm.start[b], b1 = sm.start[b + 1];
    if (b1 - b0 > kDirectMax) continue;
    int rank = b0;
    for (int q = b0; q < b1; ++q) rank += B[off + q] < k;
    out[off + rank] = (int)(k & 0xffffffffu);
  }
  // 6. the others: shared-memory small sort, or another partition level
  const int big_n = sm.big_n;
  const bool listed = big_n <= kMsdListMax;
  const int n_iter = listed ? big_n : nb;                          // list overflow: look at every bucket
  for (int it = 0; it < n_iter; ++it) {
    const int b = listed ? sm.big_list[it] : it;
    const int b0 = sm.start[b], m = sm.start[b + 1] - b0;
    if (m <= kDirectMax) continue;
    if (m <= kRankSortMax) {
      small_sort(sm.small, B + off + b0, m, out + off + b0);
    } else {
      __syncthreads();
      const int top = sm.stk_n;
      __syncthreads();
      if (top < kMsdStack) {
        if (threadIdx.x == 0) { sm.stk_off[top] = off + b0; sm.stk_len[top] = m; sm.stk_lvl[top] = lvl + 1; sm.stk_n = top + 1; }
      } else {
        slow_sort(B + off + b0, m, out + off + b0);
      }
      __syncthreads();
    }
  }
  __syncthreads();
}

constexpr int kMsdChunk = 32;         // segments claimed per atomic: one warp looks at their sizes
__global__ void __launch_bounds__(kSortThreads) tile_msd_sort_kernel(MobgsTileSort a, int nt) {
  __shared__ __align__(16) MsdSmem sm;
  __shared__ int chunk_big[kMsdChunk];
  __shared__ int chunk_n;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) { sm.seg = atomicAdd(a.tile_cursor, kMsdChunk); chunk_n = 0; }    // counter zeroed by the launcher
    __syncthreads();
    const int c0 = sm.seg;
    if (c0 >= nt) return;
    if (threadIdx.x < kMsdChunk && c0 + threadIdx.x < nt) {
      const int seg = c0 + threadIdx.x;
      const int beg = a.tile_offsets[seg];
      int n = a.tile_offsets[seg + 1] - beg;
      if ((int64_t)beg + n > a.capacity) n = (int)max((int64_t)0, a.capacity - beg);
      if (n > kRankSortMax) chunk_big[atomicAdd(&chunk_n, 1)] = seg;      // the others are tile_bucket_sort_kernel's
    }
    __syncthreads();
    const int nbig = chunk_n;
    for (int bi = 0; bi < nbig; ++bi) {
      const int seg = chunk_big[bi];
      const int beg = a.tile_offsets[seg];
      int n = a.tile_offsets[seg + 1] - beg;
      if ((int64_t)beg + n > a.capacity) n = (int)max((int64_t)0, a.capacity - beg);
      __syncthreads();
      if (threadIdx.x == 0) { sm.stk_n = 1; sm.stk_off[0] = 0; sm.stk_len[0] = n; sm.stk_lvl[0] = 0; }
      for (;;) {
        __syncthreads();
        const int top = sm.stk_n;
        if (top == 0) break;
        const int off = sm.stk_off[top - 1], len = sm.stk_len[top - 1], lvl = sm.stk_lvl[top - 1];
        __syncthreads();
        if (threadIdx.x == 0) sm.stk_n = top - 1;
        __syncthreads();
        uint64_t* A = ((lvl & 1) ? a.keys_tmp : a.keys) + beg;
        uint64_t* B = ((lvl & 1) ? a.keys : a.keys_tmp) + beg;
        msd_partition(sm, A, B, a.sorted_ids + beg, off, len, lvl);
      }
    }
  }
}

}  // namespace mobgs

using namespace mobgs;

extern "C" int mobgs_tile_count(const MobgsTileCount* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->K >= 1 && a->K <= MOBGS_MAX_K && a->N >= 0 && a->width > 0 && a->height > 0, "bad extents");
  MOBGS_REQUIRE(a->tile_counts && a->tile_offsets, "NULL workspace");
  cudaStream_t s = (cudaStream_t)stream;
  const int tiles_x = (a->width + kTile - 1) / kTile, tiles_y = (a->height + kTile - 1) / kTile;
  const int nt = a->K * tiles_x * tiles_y;
  if (a->counts_ready) {       // mobgs_synth_project_fwd already counted (MobgsSynthFwd.bin_tile_counts): prefix sum only
    scan_kernel<<<1, 1024, 0, s>>>(a->tile_counts, a->tile_offsets, nt);
    return check_launch("tile_count (scan)");
  }
  cudaMemsetAsync(a->tile_counts, 0, sizeof(int) * (size_t)nt, s);
  if (a->entries) {
    MOBGS_REQUIRE(a->entry_cursor && a->entry_capacity >= 0, "entries need entry_cursor and a capacity");
    cudaMemsetAsync(a->entry_cursor, 0, sizeof(int), s);
  }
  if (a->N > 0) {
    MOBGS_REQUIRE(a->records && a->radii, "NULL records / radii");
    const size_t total = (size_t)a->K * a->N;
    if (a->entries) {
      MOBGS_REQUIRE(a->depths, "entries need depths");
      tile_count_entries_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(*a, tiles_x, tiles_y);
    } else {
      tile_count_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(*a, tiles_x, tiles_y);
    }
  }
  scan_kernel<<<1, 1024, 0, s>>>(a->tile_counts, a->tile_offsets, nt);
  return check_launch("tile_count");
}

extern "C" int mobgs_tile_emit_sort(const MobgsTileSort* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->K >= 1 && a->K <= MOBGS_MAX_K && a->N >= 0 && a->width > 0 && a->height > 0, "bad extents");
  MOBGS_REQUIRE(a->tile_offsets && a->tile_cursor, "NULL workspace");
  if (a->N == 0 || a->capacity == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->keys && a->keys_tmp && a->sorted_ids, "NULL pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int tiles_x = (a->width + kTile - 1) / kTile, tiles_y = (a->height + kTile - 1) / kTile;
  const int nt = a->K * tiles_x * tiles_y;
  if (a->entries) {
    // the counting pass already recorded (segment, slot, key) of every intersection: streaming scatter
    MOBGS_REQUIRE(a->n_entries && a->entry_capacity >= 0, "entries need n_entries and their capacity");
    if (a->entry_capacity > 0)
      tile_scatter_kernel<<<(unsigned)((a->entry_capacity + 255) / 256), 256, 0, s>>>(*a);
  } else {
    MOBGS_REQUIRE(a->records && a->radii && a->depths, "NULL pointer");
    cudaMemsetAsync(a->tile_cursor, 0, sizeof(int) * (size_t)nt, s);
    const size_t total = (size_t)a->K * a->N;
    tile_emit_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(*a, tiles_x, tiles_y);
  }
#if MOBGS_RANK_SORT_BUCKETS
  tile_bucket_sort_kernel<<<nt, kSortThreads, 0, s>>>(*a);
#else
  tile_rank_sort_kernel<<<nt, kSortThreads, 0, s>>>(*a);
#endif
  static int sort_ctas = 0;           // persistent CTAs for the large segments
  if (!sort_ctas) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sort_ctas = 4 * sms;
  }
  cudaMemsetAsync(a->tile_cursor, 0, sizeof(int), s);     // (the emit pass is done with it) segment counter of the MSD sort
  tile_msd_sort_kernel<<<min(sort_ctas, nt), kSortThreads, 0, s>>>(*a, nt);
  return check_launch("tile_emit_sort");
}
