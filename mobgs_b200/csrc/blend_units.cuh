// Sub-warp pixel units of the blend kernels and the exact (unit, Gaussian) culling test.
// MOBGS_HD so that tests/host_math/ can brute-force the test on the CPU (every pixel with
// alpha >= 1/255 must lie in a unit whose mask bit is set).
#pragma once
#include "gs_math.cuh"

namespace mobgs {

// The 256 pixels of a 16x16 tile are split into rectangular units of L consecutive lanes
// (32: 8x4 px, one unit per warp; 16: 4x4 px, two per warp; 8: 4x2 px, four per warp).  Every unit
// owns a compacted list of the batch entries that can reach it with alpha >= 1/255, so a warp works
// on up to 32 / L different Gaussians at a time (one per unit) instead of dragging all 32 lanes
// through every Gaussian that touches any of them: with the 2-10 px footprints of the benchmark scene
// a (tile, Gaussian) entry has ~25 contributing pixels, i.e. 3.2 warp iterations as 16x2 strips,
// 2.35 as 8x4 blocks, ~1.7 as pairs of 4x4 units.
template <int L>
struct UnitGeom {
  static_assert(L == 32 || L == 16 || L == 8, "unit size must be 32, 16 or 8 lanes");
  static constexpr int kUW = L == 32 ? 8 : 4;      // unit width  (pixels)
  static constexpr int kUH = L / kUW;              // unit height (pixels)
  static constexpr int kUX = 16 / kUW;             // units per tile row
  static constexpr int kUY = 16 / kUH;             // unit rows ("bands") per tile
  static constexpr int kUnits = 256 / L;           // units per tile (8 / 16 / 32): one mask bit each
  static constexpr int kUPW = 32 / L;              // units per warp
  static constexpr unsigned kAll = kUnits == 32 ? 0xffffffffu : ((1u << (kUnits & 31)) - 1u);
};

MOBGS_HD float um_log(float x) {
#if defined(__CUDA_ARCH__)
  return __logf(x);
#else
  return logf(x);
#endif
}
MOBGS_HD float um_rcp(float x) {
#if defined(__CUDA_ARCH__)
  return __fdividef(1.f, x);
#else
  return 1.f / x;
#endif
}
MOBGS_HD float um_sqrt(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return sqrtf(x);
#endif
}

// Which units of the tile can a Gaussian (record words r0 = x y opac ca, r1 = cb cc ..) reach with
// alpha >= 1/255, i.e. which unit rectangles (of pixel centres) intersect the ellipse
// sigma <= tau = log(255 opac)?  Exact up to the safety margins (+0.01 on tau, +1e-3 px on the
// extents): per band of unit rows the ellipse restricted to the band is convex, its projection onto
// x is the interval [xl, xr] whose ends are the ellipse's left / right boundary curves evaluated at
// the band row closest to the ellipse's left- / rightmost point.  A unit outside it contributes
// nothing, so it skips the Gaussian without evaluating it.
template <int L>
MOBGS_HD unsigned unit_mask(float mx, float my, float opac, float a, float b, float c, float tile_x0, float tile_y0) {
  using G = UnitGeom<L>;
  const float tau = um_log(255.f * opac) + 0.01f;
  if (!(tau >= 0.f)) return 0u;
  const float det = a * c - b * b;
  if (!(det > 0.f) || !(a > 0.f)) return G::kAll;
  const float q = 2.f * tau;
  const float inv_det = um_rcp(det), inv_a = um_rcp(a);
  const float ey = um_sqrt(q * a * inv_det) + 1e-3f;          // vertical half-extent
  const float ex = um_sqrt(q * c * inv_det);                  // horizontal half-extent
  const float dy_r = -b * ex * um_rcp(c);                     // row offset of the rightmost point
  const float qa = q * a;
  const float rx = mx - tile_x0, ry = my - tile_y0;           // tile-local centre
  unsigned m = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int band = 0; band < G::kUY; ++band) {
    const float dyl = (float)(band * G::kUH) + 0.5f - ry, dyh = dyl + (float)(G::kUH - 1);
    if (dyl > ey || dyh < -ey) continue;
    const float dr = fminf(dyh, fmaxf(dyl, dy_r)), dl = fminf(dyh, fmaxf(dyl, -dy_r));
    const float xr = (-b * dr + um_sqrt(fmaxf(qa - det * dr * dr, 0.f))) * inv_a + 1e-3f;
    const float xl = (-b * dl - um_sqrt(fmaxf(qa - det * dl * dl, 0.f))) * inv_a - 1e-3f;
    // unit ux covers tile-local pixel centres [kUW ux + 0.5, kUW ux + kUW - 0.5]
    const float ur = fminf(floorf((rx + xr - 0.5f) * (1.f / G::kUW)), (float)(G::kUX - 1));
    const float ul = fmaxf(ceilf((rx + xl + 0.5f) * (1.f / G::kUW) - 1.f), 0.f);
    if (ul <= ur) {
      const int iur = (int)ur, iul = (int)ul;
      m |= ((2u << iur) - (1u << iul)) << (band * G::kUX);
    }
  }
  return m;
}

}  // namespace mobgs
