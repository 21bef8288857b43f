// Camera ray of one pixel from 12 pose floats (scene/cameras.py:132-146 with get_pixels_torch :244-253,
// pixels_to_local_viewdirs_torch :255-266, pixels_to_viewdirs_torch :268-284) — shared by camera_rays.cu (which
// materialises Camera.cam_ray) and by the fused decoder epilogue / prologue of blend.cu (which never does):
//     l(p) = normalise( ((x + 0.5 - ppx) / sfx, (y + 0.5 - ppy) / sfy, 1) )        local direction of pixel p
//     d(p) = normalise( R l(p) )                                                    R = camera-to-world rotation
//     ray  = [ c | d ]                                                              c = camera centre
#pragma once

namespace mobgs {

struct RayCam { float r[9], c[3]; };

struct RayIntr { float ppx, ppy, sfx, sfy; };

__device__ __forceinline__ RayCam load_ray_cam(const float* rot, const float* centre, int k) {
  RayCam m;
#pragma unroll
  for (int i = 0; i < 9; ++i) m.r[i] = rot[9 * k + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) m.c[i] = centre[3 * k + i];
  return m;
}

// FAST = false: IEEE sqrt / divisions, the arithmetic of camera_rays.cu (pinned to the reference's torch ops);
// FAST = true (the in-register rays of the blend kernels): one MUFU.RSQ per normalisation and a multiplication by the
// reciprocal scale factors — 2^-22 relative error per component, ~20 instead of ~60 instructions per pixel.
template <bool FAST = false>
__device__ __forceinline__ void ray_local_dir(const RayIntr& in, int x, int y, float (&l)[3]) {
  if (FAST) {
    const float lx = ((float)x + 0.5f - in.ppx) * __frcp_rn(in.sfx), ly = ((float)y + 0.5f - in.ppy) * __frcp_rn(in.sfy);
    const float rn = rsqrtf(lx * lx + ly * ly + 1.f);
    l[0] = lx * rn; l[1] = ly * rn; l[2] = rn;
  } else {
    const float lx = ((float)x + 0.5f - in.ppx) / in.sfx, ly = ((float)y + 0.5f - in.ppy) / in.sfy;
    const float n = sqrtf(lx * lx + ly * ly + 1.f);
    l[0] = lx / n; l[1] = ly / n; l[2] = 1.f / n;
  }
}

// world direction before normalisation
__device__ __forceinline__ void ray_rotate(const float* r, const float (&l)[3], float (&w)[3]) {
  w[0] = r[0] * l[0] + r[1] * l[1] + r[2] * l[2];
  w[1] = r[3] * l[0] + r[4] * l[1] + r[5] * l[2];
  w[2] = r[6] * l[0] + r[7] * l[1] + r[8] * l[2];
}

// pose = R (9, row-major) | c (3); rays[6] = [c | d]
template <bool FAST = false>
__device__ __forceinline__ void pixel_ray(const float* pose, const RayIntr& in, int x, int y, float (&rays)[6]) {
  float l[3], w[3];
  ray_local_dir<FAST>(in, x, y, l);
  ray_rotate(pose, l, w);
  const float n2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const float inv_n = FAST ? rsqrtf(n2) : 1.f / sqrtf(n2);
  rays[0] = pose[9]; rays[1] = pose[10]; rays[2] = pose[11];
  if (FAST) { rays[3] = w[0] * inv_n; rays[4] = w[1] * inv_n; rays[5] = w[2] * inv_n; }
  else { const float n = sqrtf(n2); rays[3] = w[0] / n; rays[4] = w[1] / n; rays[5] = w[2] / n; }
}

// VJP of pixel_ray: g[6] = d loss / d rays -> the 12 pose-gradient terms of this pixel (v[0..8] = d R, v[9..11] = d c)
template <bool FAST = false>
__device__ __forceinline__ void pixel_ray_vjp(const float* pose, const RayIntr& in, int x, int y, const float (&g)[6],
                                              float (&v)[12]) {
  float l[3], w[3];
  ray_local_dir<FAST>(in, x, y, l);
  ray_rotate(pose, l, w);
  const float n2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const float inv_n = FAST ? rsqrtf(n2) : 1.f / sqrtf(n2);
  const float dx = w[0] * inv_n, dy = w[1] * inv_n, dz = w[2] * inv_n;
  const float dot = dx * g[3] + dy * g[4] + dz * g[5];
  const float vx = (g[3] - dx * dot) * inv_n, vy = (g[4] - dy * dot) * inv_n, vz = (g[5] - dz * dot) * inv_n;   // d/dw of w/|w|
#pragma unroll
  for (int j = 0; j < 3; ++j) { v[j] = vx * l[j]; v[3 + j] = vy * l[j]; v[6 + j] = vz * l[j]; }
  v[9] = g[0]; v[10] = g[1]; v[11] = g[2];
}

}  // namespace mobgs
