// Depth -> camera-space surface normals: main_utils.get_normals (main_utils.py:95-141), which train.py runs once per
// view and step on the centre render's depth (train.py:590) through numpy pixel grids, a host->device copy of the
// [H,W,3] view directions and ~12 torch launches.  One launch here:
//     y = (v + off - ppy) / sfy,  x = (u + off - ppx - y skew) / sfx           local view direction (x, y, 1) of pixel (u, v)
//     c(u, v) = (x, y, 1) z(u, v)                                               back-projected point
//     n = normalise( (c(u+1, v) - c(u-1, v)) x (c(u, v-1) - c(u, v+1)) )        F.normalize: n / max(|n|, 1e-12)
// for interior pixels, zero on the one-pixel border (F.pad constant).  z [B,H,W] -> normals [B,3,H,W].
// Streaming: 4 B read (neighbours from L1/L2) + 12 B written per pixel; HBM-bound.
#include "common.cuh"

namespace mobgs {

constexpr int kNrmThreads = 256;

__global__ void __launch_bounds__(kNrmThreads) depth_normals_kernel(const __grid_constant__ MobgsNormals a) {
  const int64_t P = (int64_t)a.width * a.height;
  const int64_t total = P * a.B;
  const int64_t stride = (int64_t)gridDim.x * kNrmThreads;
  for (int64_t i = (int64_t)blockIdx.x * kNrmThreads + threadIdx.x; i < total; i += stride) {
    const int b = (int)(i / P);
    const int64_t pp = i - (int64_t)b * P;
    const int v = (int)(pp / a.width), u = (int)(pp - (int64_t)v * a.width);
    float n0 = 0.f, n1 = 0.f, n2 = 0.f;
    if (u > 0 && u < a.width - 1 && v > 0 && v < a.height - 1) {
      const float* z = a.z + (int64_t)b * P;
      // view directions exactly as the reference evaluates them (fp32, division — not a reciprocal multiply)
      const float yc = __fdiv_rn((float)v + a.pixel_offset - a.ppy, a.sfy);
      const float yt = __fdiv_rn((float)(v - 1) + a.pixel_offset - a.ppy, a.sfy);
      const float yb = __fdiv_rn((float)(v + 1) + a.pixel_offset - a.ppy, a.sfy);
      const float xl = __fdiv_rn((float)(u - 1) + a.pixel_offset - a.ppx - __fmul_rn(yc, a.skew), a.sfx);
      const float xr = __fdiv_rn((float)(u + 1) + a.pixel_offset - a.ppx - __fmul_rn(yc, a.skew), a.sfx);
      const float xt = __fdiv_rn((float)u + a.pixel_offset - a.ppx - __fmul_rn(yt, a.skew), a.sfx);
      const float xb = __fdiv_rn((float)u + a.pixel_offset - a.ppx - __fmul_rn(yb, a.skew), a.sfx);
      const float zl = __ldg(z + pp - 1), zr = __ldg(z + pp + 1);
      const float zt = __ldg(z + pp - a.width), zb = __ldg(z + pp + a.width);
      // left_to_right = right - left, bottom_to_top = top - bottom
      const float lx = __fmul_rn(xr, zr) - __fmul_rn(xl, zl), ly = __fmul_rn(yc, zr) - __fmul_rn(yc, zl), lz = zr - zl;
      const float tx = __fmul_rn(xt, zt) - __fmul_rn(xb, zb), ty = __fmul_rn(yt, zt) - __fmul_rn(yb, zb), tz = zt - zb;
      const float cx = __fmul_rn(ly, tz) - __fmul_rn(lz, ty);
      const float cy = __fmul_rn(lz, tx) - __fmul_rn(lx, tz);
      const float cz = __fmul_rn(lx, ty) - __fmul_rn(ly, tx);
      const float len = sqrtf(__fmul_rn(cx, cx) + __fmul_rn(cy, cy) + __fmul_rn(cz, cz));
      const float den = fmaxf(len, 1e-12f);
      n0 = __fdiv_rn(cx, den); n1 = __fdiv_rn(cy, den); n2 = __fdiv_rn(cz, den);
    }
    float* o = a.normals + (int64_t)b * 3 * P + pp;
    o[0] = n0; o[P] = n1; o[2 * P] = n2;
  }
}

}  // namespace mobgs

using namespace mobgs;

extern "C" int mobgs_depth_normals(const MobgsNormals* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->B >= 0 && a->width > 0 && a->height > 0, "bad extents");
  MOBGS_REQUIRE(a->sfx != 0.f && a->sfy != 0.f, "scale factors must be non-zero");
  if (a->B == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->z && a->normals, "NULL pointer");
  static int max_ctas = 0;
  if (!max_ctas) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    max_ctas = 8 * sms;
  }
  const int64_t total = (int64_t)a->B * a->width * a->height;
  const int64_t want = (total + kNrmThreads - 1) / kNrmThreads;
  const int grid = want < (int64_t)max_ctas ? (int)want : max_ctas;
  depth_normals_kernel<<<grid, kNrmThreads, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("depth_normals");
}
