// K9 / f4: simple_knn._C.distCUDA2 — for every point the mean of the squared
// distances to its 3 nearest neighbours (the point itself excluded), used by
// GaussianModel.create_from_pcd* to initialise the scales (scene/gaussian_model.py:420, :514; N ~ 30 k,
// once at start-up).  simple-knn (gitlab.inria.fr/bkerbl/simple-knn, un-vendored, unpinned) sorts the points
// along a Morton curve and prunes boxes; its result is the exact 3-NN mean, which at these sizes a tiled
// brute-force pass delivers in well under a millisecond: every thread owns one query, the points stream
// through shared memory in tiles of 1024, three running minima per thread.  O(N^2) — meant for the
// initialisation point clouds (<= a few 100 k points), not for the 1 M-Gaussian scene.
#include "common.cuh"

namespace mobgs {

constexpr int kKnnThreads = 256;
constexpr int kKnnTile = 1024;

__global__ void __launch_bounds__(kKnnThreads) knn3_mean_dist2_kernel(const float* __restrict__ pts, float* __restrict__ out, int n) {
  __shared__ float sx[kKnnTile], sy[kKnnTile], sz[kKnnTile];
  const int q = blockIdx.x * kKnnThreads + threadIdx.x;
  const bool live = q < n;
  const float qx = live ? pts[3 * q] : 0.f, qy = live ? pts[3 * q + 1] : 0.f, qz = live ? pts[3 * q + 2] : 0.f;
  float b0 = 3.402823466e38f, b1 = b0, b2 = b0;
  for (int t0 = 0; t0 < n; t0 += kKnnTile) {
    const int m = min(kKnnTile, n - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += kKnnThreads) {
      sx[i] = pts[3 * (t0 + i)]; sy[i] = pts[3 * (t0 + i) + 1]; sz[i] = pts[3 * (t0 + i) + 2];
    }
    __syncthreads();
    if (live) {
#pragma unroll 4
      for (int i = 0; i < m; ++i) {
        const float dx = qx - sx[i], dy = qy - sy[i], dz = qz - sz[i];
        const float d = dx * dx + dy * dy + dz * dz;
        if (d < b2 && t0 + i != q) {
          if (d < b1) { b2 = b1; if (d < b0) { b1 = b0; b0 = d; } else b1 = d; } else b2 = d;
        }
      }
    }
  }
  if (live) out[q] = (b0 + b1 + b2) / 3.0f;
}

}  // namespace mobgs

using namespace mobgs;

extern "C" int mobgs_knn3_mean_dist2(const float* points, float* out, int32_t n, void* stream) {
  MOBGS_REQUIRE(n >= 0, "n < 0");
  if (n == 0) return MOBGS_OK;
  MOBGS_REQUIRE(points && out, "NULL pointer");
  MOBGS_REQUIRE(n >= 4, "distCUDA2 needs at least 4 points (3 neighbours), got %d", n);
  knn3_mean_dist2_kernel<<<(n + kKnnThreads - 1) / kKnnThreads, kKnnThreads, 0, (cudaStream_t)stream>>>(points, out, n);
  return check_launch("knn3_mean_dist2");
}
