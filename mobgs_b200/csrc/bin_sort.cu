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

namespace mobgs {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortSmemCap = 4096;   // keys per segment sorted entirely in shared memory
constexpr int kRankSortMax = 768;    // segments up to this size use the O(n^2 / 256) rank sort

struct TileRect { int x0, y0, x1, y1; };

__device__ __forceinline__ TileRect tile_rect(float mx, float my, int radius, int tiles_x, int tiles_y) {
  // gsplat isect_tiles: tile_min inclusive, tile_max exclusive, float->uint casts saturate at 0
  const float tr = (float)radius / kTile, tx = mx / kTile, ty = my / kTile;
  TileRect r;
  r.x0 = min(max(0, (int)floorf(tx - tr)), tiles_x);
  r.y0 = min(max(0, (int)floorf(ty - tr)), tiles_y);
  r.x1 = min(max(0, (int)ceilf(tx + tr)), tiles_x);
  r.y1 = min(max(0, (int)ceilf(ty + tr)), tiles_y);
  return r;
}

// Smallest sigma = 0.5(a dx^2 + c dy^2) + b dx dy over the rectangle of pixel centres of a tile.
// nb_c = -b / c and nb_a = -b / a are computed once per Gaussian (approximate reciprocals: their error is
// far inside the +0.01 margin on tau).
__device__ __forceinline__ float min_sigma_rect(float mx, float my, float a, float b, float c, float nb_c, float nb_a,
                                                float xlo, float xhi, float ylo, float yhi) {
  const float dxl = xlo - mx, dxh = xhi - mx, dyl = ylo - my, dyh = yhi - my;
  if (dxl <= 0.f && dxh >= 0.f && dyl <= 0.f && dyh >= 0.f) return 0.f;
  float best = 3.4e38f;
  // vertical edges: dx fixed, minimise over dy in [dyl, dyh]
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const float dx = e ? dxh : dxl;
    const float dy = fminf(dyh, fmaxf(dyl, nb_c * dx));
    best = fminf(best, 0.5f * (a * dx * dx + c * dy * dy) + b * dx * dy);
  }
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const float dy = e ? dyh : dyl;
    const float dx = fminf(dxh, fmaxf(dxl, nb_a * dy));
    best = fminf(best, 0.5f * (a * dx * dx + c * dy * dy) + b * dx * dy);
  }
  return best;
}

struct GaussGeom { float mx, my, opac, ca, cb, cc; };

__device__ __forceinline__ GaussGeom load_geom(const float* rec) {
  const float4 r0 = *reinterpret_cast<const float4*>(rec);
  const float2 r1 = *reinterpret_cast<const float2*>(rec + 4);
  return {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y};
}

// Visits every tile the reference would list for this Gaussian (optionally minus the provably
// empty ones) and calls f(tile_index).
template <typename F>
__device__ __forceinline__ void for_each_tile(const GaussGeom& g, int radius, int width, int height,
                                              int tiles_x, int tiles_y, int tight, F f) {
  const TileRect r = tile_rect(g.mx, g.my, radius, tiles_x, tiles_y);
  float tau = 0.f, nb_c = 0.f, nb_a = 0.f;
  if (tight) {
    nb_c = -g.cb * __fdividef(1.f, g.cc);
    nb_a = -g.cb * __fdividef(1.f, g.ca);
    // a pixel contributes iff opac * exp(-sigma) >= 1/255  <=>  sigma <= log(255 opac)
    tau = __logf(255.f * g.opac) + 0.01f;   // +0.01: safety margin for fp rounding
    if (!(tau >= 0.f)) return;
  }
  for (int ty = r.y0; ty < r.y1; ++ty)
    for (int tx = r.x0; tx < r.x1; ++tx) {
      if (tight) {
        const float xlo = tx * kTile + 0.5f, ylo = ty * kTile + 0.5f;
        const float xhi = fminf((float)(tx * kTile + kTile), (float)width) - 0.5f;
        const float yhi = fminf((float)(ty * kTile + kTile), (float)height) - 0.5f;
        if (min_sigma_rect(g.mx, g.my, g.ca, g.cb, g.cc, nb_c, nb_a, xlo, xhi, ylo, yhi) > tau) continue;
      }
      f(ty * tiles_x + tx);
    }
}

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

// Counting pass that also REMEMBERS what it found (a.entries != NULL): the atomicAdd on the tile counter returns the
// entry's slot inside its segment, so (segment, slot, key) is everything the emit pass needs — it becomes a streaming
// scatter over the entries (tile_scatter_kernel) instead of a second pass over all K*N records that repeats the
// exact tile tests and hits the same L2 atomics again.  Entries land in one global array in arbitrary order; a warp
// reserves its range with ONE atomicAdd on the cursor (prefix sum of the lanes' hit counts).  A Gaussian's hits are
// kept as a bit mask over its tile rectangle between counting and writing (rectangles of more than 64 tiles repeat
// the test instead).  Entries beyond entry_capacity are dropped; the counters stay exact, so the caller sees the
// overflow in tile_offsets[K*T] and redoes the binning with the two-pass kernels (the lists of the overflowed attempt
// hold unwritten keys: the blend kernels clamp every Gaussian index they read).
struct BinEntry { int seg; int slot; uint32_t gid; uint32_t depth_bits; };   // 16 bytes, stored as one uint4

__global__ void __launch_bounds__(256) tile_count_entries_kernel(const __grid_constant__ MobgsTileCount a, int tiles_x, int tiles_y) {
  const size_t iv = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int tiles = tiles_x * tiles_y;
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
  TileRect r = {0, 0, 0, 0};
  float tau = 0.f, nb_c = 0.f, nb_a = 0.f;
  if (live) {
    g = load_geom(a.records + i * kRecFloats);
    r = tile_rect(g.mx, g.my, radius, tiles_x, tiles_y);
    if (a.tight) {
      nb_c = -g.cb * __fdividef(1.f, g.cc);
      nb_a = -g.cb * __fdividef(1.f, g.ca);
      tau = __logf(255.f * g.opac) + 0.01f;
      if (!(tau >= 0.f)) live = false;
    }
  }
  const int rw = live ? r.x1 - r.x0 : 0, rh = live ? r.y1 - r.y0 : 0;
  auto hit = [&](int tx, int ty) {
    if (!a.tight) return true;
    const float xlo = tx * kTile + 0.5f, ylo = ty * kTile + 0.5f;
    const float xhi = fminf((float)(tx * kTile + kTile), (float)a.width) - 0.5f;
    const float yhi = fminf((float)(ty * kTile + kTile), (float)a.height) - 0.5f;
    return !(min_sigma_rect(g.mx, g.my, g.ca, g.cb, g.cc, nb_c, nb_a, xlo, xhi, ylo, yhi) > tau);
  };
  const bool small = rw * rh <= 64;
  unsigned long long mask = 0ull;
  int n = 0;
  for (int y = 0; y < rh; ++y)
    for (int x = 0; x < rw; ++x)
      if (hit(r.x0 + x, r.y0 + y)) {
        if (small) mask |= 1ull << (y * rw + x);
        ++n;
      }
  // one reservation per warp
  int incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;
  int base = 0;
  if (lane == 31) base = atomicAdd(a.entry_cursor, total);
  base = __shfl_sync(0xffffffffu, base, 31);
  if (n == 0) return;
  int64_t pos = (int64_t)base + incl - n;
  int* counts = a.tile_counts + (size_t)k * tiles;
  const uint32_t dbits = __float_as_uint(a.depths[i]);
  uint4* entries = reinterpret_cast<uint4*>(a.entries);
  for (int y = 0; y < rh; ++y)
    for (int x = 0; x < rw; ++x) {
      const bool h = small ? ((mask >> (y * rw + x)) & 1ull) != 0ull : hit(r.x0 + x, r.y0 + y);
      if (!h) continue;
      const int tile = (r.y0 + y) * tiles_x + (r.x0 + x);
      const int slot = atomicAdd(counts + tile, 1);
      if (pos < a.entry_capacity) entries[pos] = make_uint4((unsigned)(k * tiles + tile), (unsigned)slot, (unsigned)gid, dbits);
      ++pos;
    }
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
__device__ __forceinline__ void radix_pass(const uint64_t* src, uint64_t* dst, int n, int shift,
                                           int (*warp_hist)[256], int* digit_base) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = (((n + kSortWarps - 1) / kSortWarps) + 31) & ~31;
  const int beg = min(n, warp * chunk), end = min(n, beg + chunk);
  int* hist = warp_hist[warp];
  for (int d = lane; d < 256; d += 32) hist[d] = 0;
  __syncwarp();
  for (int i = beg; i < end; i += 32) {
    const int idx = i + lane;
    const bool valid = idx < end;
    const int digit = valid ? (int)((src[idx] >> shift) & 0xff) : 256;
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    if (valid && lane == (__ffs(peers) - 1)) hist[digit] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {
    // thread d owns digit d: exclusive scan over warps, then over digits
    const int d = threadIdx.x;
    int tot = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) { const int c = warp_hist[w][d]; warp_hist[w][d] = tot; tot += c; }
    int s = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
    if (lane == 31) digit_base[warp] = s;
    __syncthreads();
    int prev = 0;
    for (int w = 0; w < warp; ++w) prev += digit_base[w];
    const int base = prev + s - tot;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) warp_hist[w][d] += base;
  }
  __syncthreads();
  for (int i = beg; i < end; i += 32) {
    const int idx = i + lane;
    const bool valid = idx < end;
    uint64_t key = 0;
    int digit = 256;
    if (valid) { key = src[idx]; digit = (int)((key >> shift) & 0xff); }
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    if (valid) {
      const int rank = __popc(peers & ((1u << lane) - 1u));
      dst[hist[digit] + rank] = key;
    }
    __syncwarp();
    if (valid && lane == (__ffs(peers) - 1)) hist[digit] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
}

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
__global__ void __launch_bounds__(kSortThreads) tile_bucket_sort_kernel(MobgsTileSort a) {
  __shared__ __align__(16) uint64_t buf[kRankSortMax];
  __shared__ int hist[kBuckets], start[kBuckets + 1];
  __shared__ uint32_t wmin[kSortWarps], wmax[kSortWarps];
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
  if (threadIdx.x < kBuckets) hist[threadIdx.x] = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { wmin[warp] = lo; wmax[warp] = hi; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) { lo = min(lo, wmin[w]); hi = max(hi, wmax[w]); }
  const uint32_t range = hi - lo;
  const int shift = range < (uint32_t)kBuckets ? 0 : (32 - __clz(range)) - 6;   // (range >> shift) < kBuckets = 2^6
  int slot[kPer], bkt[kPer];
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int i = threadIdx.x + j * kSortThreads;
    bkt[j] = (int)(((uint32_t)(key[j] >> 32) - lo) >> shift);
    slot[j] = i < n ? atomicAdd(&hist[bkt[j]], 1) : 0;
  }
  __syncthreads();
  if (warp == 0) {          // exclusive scan over the 64 buckets, two per lane
    const int c0 = hist[2 * lane], c1 = hist[2 * lane + 1];
    int s = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
    start[2 * lane] = s - c0 - c1;
    start[2 * lane + 1] = s - c1;
    if (lane == 31) start[kBuckets] = s;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int i = threadIdx.x + j * kSortThreads;
    if (i < n) buf[start[bkt[j]] + slot[j]] = key[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kSortThreads) {
    const uint64_t k = buf[i];
    const int b = (int)(((uint32_t)(k >> 32) - lo) >> shift);
    const int b0 = start[b], b1 = start[b + 1];
    int rank = b0;
    for (int q = b0; q < b1; ++q) rank += buf[q] < k;
    a.sorted_ids[beg + rank] = (int)(k & 0xffffffffu);
  }
}

// Large segments (n > kRankSortMax, rare): LSD radix sort, one CTA per segment at a time.  The grid is a
// few CTAs per SM; each scans 256 segment sizes at once and sorts the large ones it finds.
__device__ void radix_sort_segment(const MobgsTileSort& a, int seg, uint64_t* bufA, uint64_t* bufB,
                                   int (*warp_hist)[256], int* digit_base, unsigned long long* red_or,
                                   unsigned long long* red_and) {
  const int beg = a.tile_offsets[seg];
  int n = a.tile_offsets[seg + 1] - beg;
  if ((int64_t)beg + n > a.capacity) n = (int)max((int64_t)0, a.capacity - beg);
  uint64_t* gkeys = a.keys + beg;
  const bool in_smem = n <= kSortSmemCap;
  uint64_t* src = in_smem ? bufA : gkeys;
  uint64_t* dst = in_smem ? bufB : a.keys_tmp + beg;

  unsigned long long vor = 0ull, vand = ~0ull;
  for (int i = threadIdx.x; i < n; i += kSortThreads) {
    const uint64_t key = gkeys[i];
    if (in_smem) bufA[i] = key;
    vor |= key; vand &= key;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vor |= __shfl_xor_sync(0xffffffffu, vor, o);
    vand &= __shfl_xor_sync(0xffffffffu, vand, o);
  }
  if ((threadIdx.x & 31) == 0) { red_or[threadIdx.x >> 5] = vor; red_and[threadIdx.x >> 5] = vand; }
  __syncthreads();
  vor = 0ull; vand = ~0ull;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) { vor |= red_or[w]; vand &= red_and[w]; }
  const unsigned long long diff = vor ^ vand;

  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 8 * pass;
    if (((diff >> shift) & 0xffull) == 0ull) continue;   // all keys share this digit
    radix_pass(src, dst, n, shift, warp_hist, digit_base);
    uint64_t* t = src; src = dst; dst = t;
  }
  for (int i = threadIdx.x; i < n; i += kSortThreads) a.sorted_ids[beg + i] = (int)(src[i] & 0xffffffffu);
}

__global__ void __launch_bounds__(kSortThreads) tile_sort_kernel(MobgsTileSort a, int nt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bufA = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* bufB = bufA + kSortSmemCap;
  int (*warp_hist)[256] = reinterpret_cast<int (*)[256]>(bufB + kSortSmemCap);
  __shared__ int digit_base[kSortWarps];
  __shared__ unsigned long long red_or[kSortWarps], red_and[kSortWarps];
  __shared__ int big[kSortThreads];
  __shared__ int nbig;
  for (int c0 = blockIdx.x * kSortThreads; c0 < nt; c0 += gridDim.x * kSortThreads) {
    if (threadIdx.x == 0) nbig = 0;
    __syncthreads();
    const int seg = c0 + threadIdx.x;
    if (seg < nt) {
      const int beg = a.tile_offsets[seg];
      int n = a.tile_offsets[seg + 1] - beg;
      if ((int64_t)beg + n > a.capacity) n = (int)max((int64_t)0, a.capacity - beg);
      if (n > kRankSortMax) big[atomicAdd(&nbig, 1)] = seg;
    }
    __syncthreads();
    const int nb = nbig;
    for (int b = 0; b < nb; ++b) {
      radix_sort_segment(a, big[b], bufA, bufB, warp_hist, digit_base, red_or, red_and);
      __syncthreads();
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
  MOBGS_REQUIRE(a->tile_offsets && (a->tile_cursor || a->entries), "NULL workspace");
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
  const size_t smem = 2 * sizeof(uint64_t) * kSortSmemCap + sizeof(int) * kSortWarps * 256;
  cudaFuncSetAttribute(tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#if MOBGS_RANK_SORT_BUCKETS
  tile_bucket_sort_kernel<<<nt, kSortThreads, 0, s>>>(*a);
#else
  tile_rank_sort_kernel<<<nt, kSortThreads, 0, s>>>(*a);
#endif
  static int sort_ctas = 0;           // a few CTAs per SM (the radix buffers allow 3)
  if (!sort_ctas) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sort_ctas = 3 * sms;
  }
  tile_sort_kernel<<<min(sort_ctas, (nt + kSortThreads - 1) / kSortThreads), kSortThreads, smem, s>>>(*a, nt);
  return check_launch("tile_emit_sort");
}
