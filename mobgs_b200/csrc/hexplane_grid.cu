// HexPlane feature gather + VJP (scene/hexplane.py:19-108): the non-GEMM half of
// the deformation network's backward.  The reference runs 18 F.grid_sample launches forward and 18
// backward with [N,32] intermediates; here one kernel each way.  8 lanes share a point, each owns 4
// of the 32 channels of every plane (one 16-byte texel slice: a point's corner fetch is one coalesced
// 128-byte line); the coordinate gradient is reduced over the 8 lanes with three shuffles and the
// plane gradient leaves as 16-byte vector reductions into channels-last planes.
#include "common.cuh"

namespace mobgs {

constexpr int kGC = 32;     // channels per plane

struct Samp { int o00, o01, o10, o11; float tx, ty, sx, sy; };   // sx/sy: d(pixel coord)/d(normalised coord), 0 if clipped

// F.grid_sample(bilinear, align_corners=True, padding_mode='border') addressing on a channels-last plane
__device__ __forceinline__ Samp sample_at(float x, float y, int Wd, int Hd) {
  Samp s;
  float ix = (x + 1.f) * 0.5f * (float)(Wd - 1), iy = (y + 1.f) * 0.5f * (float)(Hd - 1);
  // clip_coordinates_set_grad: no gradient through a clipped coordinate
  s.sx = (ix > 0.f && ix < (float)(Wd - 1)) ? 0.5f * (float)(Wd - 1) : 0.f;
  s.sy = (iy > 0.f && iy < (float)(Hd - 1)) ? 0.5f * (float)(Hd - 1) : 0.f;
  ix = fminf(fmaxf(ix, 0.f), (float)(Wd - 1));
  iy = fminf(fmaxf(iy, 0.f), (float)(Hd - 1));
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const int x1 = min(x0 + 1, Wd - 1), y1 = min(y0 + 1, Hd - 1);
  s.tx = ix - fx; s.ty = iy - fy;
  s.o00 = (y0 * Wd + x0) * kGC; s.o01 = (y0 * Wd + x1) * kGC;
  s.o10 = (y1 * Wd + x0) * kGC; s.o11 = (y1 * Wd + x1) * kGC;
  return s;
}

__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_lerp4(float4 v00, float4 v01, float4 v10, float4 v11, float tx, float ty) {
  const float w00 = (1.f - tx) * (1.f - ty), w01 = tx * (1.f - ty), w10 = (1.f - tx) * ty, w11 = tx * ty;
  return make_float4(v00.x * w00 + v01.x * w01 + v10.x * w10 + v11.x * w11, v00.y * w00 + v01.y * w01 + v10.y * w10 + v11.y * w11,
                     v00.z * w00 + v01.z * w01 + v10.z * w10 + v11.z * w11, v00.w * w00 + v01.w * w01 + v10.w * w10 + v11.w * w11);
}

__device__ __forceinline__ void load_coords(const MobgsHexFeat& a, int g, float c[4], float scale[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float k = 2.0f / (a.aabb[3 + i] - a.aabb[i]);
    const float v = (a.pts[3 * g + i] - a.aabb[i]) * k - 1.0f;
    scale[i] = (v >= -1.f && v <= 1.f) ? k : 0.f;      // torch.clamp passes the gradient inside [min, max]
    c[i] = fminf(fmaxf(v, -1.f), 1.f);
  }
  c[3] = a.times[g];
}

template <bool BWD>
__global__ void __launch_bounds__(256) hexplane_features_kernel(const __grid_constant__ MobgsHexFeat a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = t >> 3, c4 = t & 7;
  const bool live = g < a.N;
  const int gi = live ? g : a.N - 1;           // keep all 8 lanes of a group in the shuffles
  float c[4], scale[3];
  load_coords(a, gi, c, scale);
  float gc[4] = {0.f, 0.f, 0.f, 0.f};
  const int F = a.levels * kGC;
  for (int l = 0; l < a.levels; ++l) {
    float4 v[6];
    int pi = 0;
#pragma unroll
    for (int ca = 0; ca < 4; ++ca)
#pragma unroll
      for (int cb = ca + 1; cb < 4; ++cb, ++pi) {
        const int id = l * 6 + pi;
        const Samp s = sample_at(c[ca], c[cb], a.plane_w[id], a.plane_h[id]);
        const float* p = a.planes[id] + 4 * c4;
        v[pi] = f4_lerp4(__ldg(reinterpret_cast<const float4*>(p + s.o00)), __ldg(reinterpret_cast<const float4*>(p + s.o01)),
                         __ldg(reinterpret_cast<const float4*>(p + s.o10)), __ldg(reinterpret_cast<const float4*>(p + s.o11)), s.tx, s.ty);
      }
    if (!BWD) {
      float4 prod = v[0];
#pragma unroll
      for (int q = 1; q < 6; ++q) prod = f4_mul(prod, v[q]);
      if (live) *reinterpret_cast<float4*>(a.feat + (size_t)g * F + l * kGC + 4 * c4) = prod;
      continue;
    }
    // exclusive products: e[p] = prod_{q != p} v[q]   (prefix x suffix; no division)
    float4 e[6];
    {
      float4 pre = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
      for (int q = 0; q < 6; ++q) { e[q] = pre; pre = f4_mul(pre, v[q]); }
      float4 suf = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
      for (int q = 5; q >= 0; --q) { e[q] = f4_mul(e[q], suf); suf = f4_mul(suf, v[q]); }
    }
    const float4 gf = live ? __ldg(reinterpret_cast<const float4*>(a.g_feat + (size_t)g * F + l * kGC + 4 * c4))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    pi = 0;
#pragma unroll
    for (int ca = 0; ca < 4; ++ca)
#pragma unroll
      for (int cb = ca + 1; cb < 4; ++cb, ++pi) {
        const int id = l * 6 + pi;
        const Samp s = sample_at(c[ca], c[cb], a.plane_w[id], a.plane_h[id]);
        const float4 gv = f4_mul(gf, e[pi]);                     // d loss / d (this plane's sample)
        const float* p = a.planes[id] + 4 * c4;
        const float4 v00 = __ldg(reinterpret_cast<const float4*>(p + s.o00)), v01 = __ldg(reinterpret_cast<const float4*>(p + s.o01));
        const float4 v10 = __ldg(reinterpret_cast<const float4*>(p + s.o10)), v11 = __ldg(reinterpret_cast<const float4*>(p + s.o11));
        // plane gradient: four 16-byte reductions
        if (live) {
          float* gp = a.g_planes[id] + 4 * c4;
          const float w00 = (1.f - s.tx) * (1.f - s.ty), w01 = s.tx * (1.f - s.ty), w10 = (1.f - s.tx) * s.ty, w11 = s.tx * s.ty;
          red_add_v4(gp + s.o00, gv.x * w00, gv.y * w00, gv.z * w00, gv.w * w00);
          red_add_v4(gp + s.o01, gv.x * w01, gv.y * w01, gv.z * w01, gv.w * w01);
          red_add_v4(gp + s.o10, gv.x * w10, gv.y * w10, gv.z * w10, gv.w * w10);
          red_add_v4(gp + s.o11, gv.x * w11, gv.y * w11, gv.z * w11, gv.w * w11);
        }
        // coordinate gradient: d sample / d ix = (v01 - v00)(1 - ty) + (v11 - v10) ty, etc.
        float gx = gv.x * ((v01.x - v00.x) * (1.f - s.ty) + (v11.x - v10.x) * s.ty) + gv.y * ((v01.y - v00.y) * (1.f - s.ty) + (v11.y - v10.y) * s.ty) +
                   gv.z * ((v01.z - v00.z) * (1.f - s.ty) + (v11.z - v10.z) * s.ty) + gv.w * ((v01.w - v00.w) * (1.f - s.ty) + (v11.w - v10.w) * s.ty);
        float gy = gv.x * ((v10.x - v00.x) * (1.f - s.tx) + (v11.x - v01.x) * s.tx) + gv.y * ((v10.y - v00.y) * (1.f - s.tx) + (v11.y - v01.y) * s.tx) +
                   gv.z * ((v10.z - v00.z) * (1.f - s.tx) + (v11.z - v01.z) * s.tx) + gv.w * ((v10.w - v00.w) * (1.f - s.tx) + (v11.w - v01.w) * s.tx);
        gc[ca] += gx * s.sx;
        gc[cb] += gy * s.sy;
      }
  }
  if (BWD) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      gc[i] += __shfl_xor_sync(0xffffffffu, gc[i], 1);
      gc[i] += __shfl_xor_sync(0xffffffffu, gc[i], 2);
      gc[i] += __shfl_xor_sync(0xffffffffu, gc[i], 4);
    }
    if (live && c4 == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) a.g_pts[3 * g + i] = gc[i] * scale[i];
      if (a.g_times) a.g_times[g] = gc[3];
    }
  }
}

}  // namespace mobgs

using namespace mobgs;

static int check_hexfeat(const MobgsHexFeat* a, bool bwd) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->N >= 0 && a->levels >= 1 && a->levels <= 4, "bad N / levels");
  if (a->N == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->pts && a->times, "NULL input");
  for (int i = 0; i < a->levels * 6; ++i) {
    MOBGS_REQUIRE(a->planes[i] && a->plane_w[i] >= 1 && a->plane_h[i] >= 1, "bad plane %d", i);
    if (bwd) MOBGS_REQUIRE(a->g_planes[i], "NULL g_planes[%d]", i);
  }
  if (bwd) MOBGS_REQUIRE(a->g_feat && a->g_pts, "NULL gradient pointer");
  else MOBGS_REQUIRE(a->feat, "NULL feat");
  return MOBGS_OK;
}

extern "C" int mobgs_hexplane_features_fwd(const MobgsHexFeat* a, void* stream) {
  if (int e = check_hexfeat(a, false)) return e;
  if (a->N == 0) return MOBGS_OK;
  const size_t threads = (size_t)a->N * 8;
  hexplane_features_kernel<false><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("hexplane_features_fwd");
}

extern "C" int mobgs_hexplane_features_bwd(const MobgsHexFeat* a, void* stream) {
  if (int e = check_hexfeat(a, true)) return e;
  if (a->N == 0) return MOBGS_OK;
  const size_t threads = (size_t)a->N * 8;
  hexplane_features_kernel<true><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("hexplane_features_bwd");
}
