// f3: Camera.cam_ray for K cameras in one launch (scene/cameras.py:132-146 with
// get_pixels_torch :244-253, pixels_to_local_viewdirs_torch :255-266, pixels_to_viewdirs_torch :268-284):
//     l(p) = normalise( ((x + 0.5 - ppx) / sfx, (y + 0.5 - ppy) / sfy, 1) )        local direction of pixel p
//     d(p) = normalise( R_k l(p) )                                                  R_k = camera-to-world rotation
//     cam_ray[k] = [ c_k (3 planes, broadcast) | d (3 planes) ]                     [K,6,H,W]
// The reference builds this per warped sub-frame camera with a meshgrid + ~10 elementwise launches + a
// batched 3x3 matmul over H*W rows (K times per blurry view).  Forward: streaming write, 24 B per
// (pixel, camera).  Backward (R and c carry the pose gradient): 12 sums over the pixels of
// each camera, warp shuffle -> CTA -> one atomicAdd per CTA and value.
#include "common.cuh"
#include "ray_math.cuh"

namespace mobgs {

constexpr int kRayThreads = 256;

__device__ __forceinline__ void local_dir(const MobgsCameraRays& a, int p, float (&l)[3]) {
  const int y = p / a.W, x = p - y * a.W;
  const RayIntr in = {a.ppx, a.ppy, a.sfx, a.sfy};
  ray_local_dir(in, x, y, l);
}

__global__ void __launch_bounds__(kRayThreads) camera_rays_fwd_kernel(const __grid_constant__ MobgsCameraRays a) {
  const int k = blockIdx.y;
  const RayCam m = load_ray_cam(a.rot, a.centre, k);
  const int P = a.H * a.W;
  float* out = a.rays + (size_t)k * 6 * P;
  for (int p = blockIdx.x * kRayThreads + threadIdx.x; p < P; p += gridDim.x * kRayThreads) {
    float l[3];
    local_dir(a, p, l);
    const float wx = m.r[0] * l[0] + m.r[1] * l[1] + m.r[2] * l[2];
    const float wy = m.r[3] * l[0] + m.r[4] * l[1] + m.r[5] * l[2];
    const float wz = m.r[6] * l[0] + m.r[7] * l[1] + m.r[8] * l[2];
    const float n = sqrtf(wx * wx + wy * wy + wz * wz);
    out[p] = m.c[0]; out[P + p] = m.c[1]; out[2 * P + p] = m.c[2];
    out[3 * P + p] = wx / n; out[4 * P + p] = wy / n; out[5 * P + p] = wz / n;
  }
}

__global__ void __launch_bounds__(kRayThreads) camera_rays_bwd_kernel(const __grid_constant__ MobgsCameraRays a) {
  __shared__ float red[12][kRayThreads / 32];
  const int k = blockIdx.y;
  const RayCam m = load_ray_cam(a.rot, a.centre, k);
  const int P = a.H * a.W;
  const float* g = a.v_rays + (size_t)k * 6 * P;
  float acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.f;
  for (int p = blockIdx.x * kRayThreads + threadIdx.x; p < P; p += gridDim.x * kRayThreads) {
    float l[3];
    local_dir(a, p, l);
    const float wx = m.r[0] * l[0] + m.r[1] * l[1] + m.r[2] * l[2];
    const float wy = m.r[3] * l[0] + m.r[4] * l[1] + m.r[5] * l[2];
    const float wz = m.r[6] * l[0] + m.r[7] * l[1] + m.r[8] * l[2];
    const float inv_n = 1.f / sqrtf(wx * wx + wy * wy + wz * wz);
    const float dx = wx * inv_n, dy = wy * inv_n, dz = wz * inv_n;
    const float gx = __ldg(g + 3 * P + p), gy = __ldg(g + 4 * P + p), gz = __ldg(g + 5 * P + p);
    const float dot = dx * gx + dy * gy + dz * gz;
    const float vx = (gx - dx * dot) * inv_n, vy = (gy - dy * dot) * inv_n, vz = (gz - dz * dot) * inv_n;   // d/dw of w/|w|
#pragma unroll
    for (int j = 0; j < 3; ++j) { acc[j] += vx * l[j]; acc[3 + j] += vy * l[j]; acc[6 + j] += vz * l[j]; }
    acc[9] += __ldg(g + p); acc[10] += __ldg(g + P + p); acc[11] += __ldg(g + 2 * P + p);
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const float s = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0) red[i][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kRayThreads / 32; ++w) s += red[threadIdx.x][w];
    float* dst = threadIdx.x < 9 ? a.v_rot + 9 * k + threadIdx.x : a.v_centre + 3 * k + (threadIdx.x - 9);
    atomicAdd(dst, s);
  }
}

}  // namespace mobgs

using namespace mobgs;

static int ray_grid(const MobgsCameraRays* a) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int want = (a->H * a->W + kRayThreads - 1) / kRayThreads;
  const int per_cam = max(1, (8 * sms + a->K - 1) / a->K);      // ~8 CTAs per SM over all cameras
  return min(want, per_cam);
}

extern "C" int mobgs_camera_rays_fwd(const MobgsCameraRays* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->K >= 1 && a->K <= 65535 && a->H > 0 && a->W > 0, "bad extents");
  MOBGS_REQUIRE(a->rot && a->centre && a->rays, "NULL pointer");
  MOBGS_REQUIRE(a->sfx != 0.f && a->sfy != 0.f, "zero scale factor");
  camera_rays_fwd_kernel<<<dim3(ray_grid(a), a->K), kRayThreads, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("camera_rays_fwd");
}

extern "C" int mobgs_camera_rays_bwd(const MobgsCameraRays* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->K >= 1 && a->K <= 65535 && a->H > 0 && a->W > 0, "bad extents");
  MOBGS_REQUIRE(a->rot && a->centre && a->v_rays && a->v_rot && a->v_centre, "NULL pointer");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(a->v_rot, 0, sizeof(float) * 9 * a->K, s);
  cudaMemsetAsync(a->v_centre, 0, sizeof(float) * 3 * a->K, s);
  camera_rays_bwd_kernel<<<dim3(ray_grid(a), a->K), kRayThreads, 0, s>>>(*a);
  return check_launch("camera_rays_bwd");
}
