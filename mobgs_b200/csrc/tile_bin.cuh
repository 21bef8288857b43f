// Tile binning of one projected Gaussian, shared by bin_sort.cu (mobgs_tile_count) and synth_project.cu (the counting pass
// fused into the projection): gsplat's tile rectangle, the exact "can any pixel centre of this tile reach alpha >= 1/255"
// test, and the count-and-record step.
#pragma once
#include "common.cuh"

namespace mobgs {

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

// Counting pass that also REMEMBERS what it found: the atomicAdd on the tile counter returns the entry's slot inside its
// segment, so (segment, slot, key) is everything the emit pass needs — it becomes a streaming scatter over the entries
// (tile_scatter_kernel) instead of a second pass over all K*N records that repeats the exact tile tests and hits the
// same L2 atomics again.  Entries land in one global array in arbitrary order; a converged warp reserves its range with
// ONE atomicAdd on the cursor (prefix sum of the lanes' hit counts; a diverged warp falls back to one per lane).  A
// Gaussian's hits are kept as a bit mask over its tile rectangle between counting and writing (rectangles of more than
// 64 tiles repeat the test instead).  Entries beyond `capacity` are dropped; the counters stay exact, so the caller sees
// the overflow in tile_offsets[K*T] and redoes the binning with the two-pass kernels (the lists of the overflowed
// attempt hold unwritten keys: the blend kernels clamp every Gaussian index they read).
struct BinTarget {
  int32_t* tile_counts;     // [lists * T]
  void* entries;            // [capacity] x uint4 (segment, slot, Gaussian index, depth bits); NULL = count only
  int64_t capacity;
  int32_t* cursor;
  int width, height, tiles_x, tiles_y, tight;
};

// one (list, Gaussian) pair per calling thread; `live` = the pair exists (in range, radius > 0).  Must be called by
// every thread that reaches the call site together (it looks at __activemask()).
__device__ __forceinline__ void bin_count_record(const BinTarget& b, int list, int gid, bool live, const GaussGeom& g,
                                                 int radius, float depth) {
  const int lane = threadIdx.x & 31;
  const int tiles = b.tiles_x * b.tiles_y;
  TileRect r = {0, 0, 0, 0};
  float tau = 0.f, nb_c = 0.f, nb_a = 0.f;
  if (live) {
    r = tile_rect(g.mx, g.my, radius, b.tiles_x, b.tiles_y);
    if (b.tight) {
      nb_c = -g.cb * __fdividef(1.f, g.cc);
      nb_a = -g.cb * __fdividef(1.f, g.ca);
      tau = __logf(255.f * g.opac) + 0.01f;
      if (!(tau >= 0.f)) live = false;
    }
  }
  const int rw = live ? r.x1 - r.x0 : 0, rh = live ? r.y1 - r.y0 : 0;
  auto hit = [&](int tx, int ty) {
    if (!b.tight) return true;
    const float xlo = tx * kTile + 0.5f, ylo = ty * kTile + 0.5f;
    const float xhi = fminf((float)(tx * kTile + kTile), (float)b.width) - 0.5f;
    const float yhi = fminf((float)(ty * kTile + kTile), (float)b.height) - 0.5f;
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
  int64_t pos = 0;
  if (b.entries) {
    const unsigned am = __activemask();
    if (am == 0xffffffffu) {
      int incl = n;       // one reservation per warp
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      int base = 0;
      if (total > 0 && lane == 31) base = atomicAdd(b.cursor, total);
      base = __shfl_sync(0xffffffffu, base, 31);
      pos = (int64_t)base + incl - n;
    } else if (n > 0) {
      pos = atomicAdd(b.cursor, n);
    }
  }
  if (n == 0) return;
  int* counts = b.tile_counts + (size_t)list * tiles;
  const uint32_t dbits = __float_as_uint(depth);
  uint4* entries = reinterpret_cast<uint4*>(b.entries);
  for (int y = 0; y < rh; ++y)
    for (int x = 0; x < rw; ++x) {
      const bool h = small ? ((mask >> (y * rw + x)) & 1ull) != 0ull : hit(r.x0 + x, r.y0 + y);
      if (!h) continue;
      const int tile = (r.y0 + y) * b.tiles_x + (r.x0 + x);
      const int slot = atomicAdd(counts + tile, 1);
      if (entries && pos < b.capacity) entries[pos] = make_uint4((unsigned)(list * tiles + tile), (unsigned)slot, (unsigned)gid, dbits);
      ++pos;
    }
}

}  // namespace mobgs
