// Per-pixel epilogue: expected-depth division + Sandwich RGB decoder + K-sub-frame mean, forward
// and VJP.  Replaces two cuDNN 1x1 convolutions (TF32 by default on this
// hardware), three elementwise launches and a stack+mean per decode with one fp32 kernel.
// One thread = one pixel, looping over the K sub-frames; HBM-bound (reads 44 + 24 B, writes 16 B
// per pixel and sub-frame).  The 90 decoder-weight gradients are reduced warp -> CTA -> one
// atomicAdd per CTA of a persistent grid.
#include "common.cuh"

namespace mobgs {

constexpr int kDecThreads = 256;
constexpr float kEdFloor = 1e-10f;
constexpr float kMeanEps = 1e-10f;

struct DecW { float w1[72]; float w2[18]; };

__device__ __forceinline__ void load_w(DecW& w, const float* w1, const float* w2, float* smem) {
  for (int i = threadIdx.x; i < 90; i += blockDim.x) smem[i] = i < 72 ? w1[i] : w2[i - 72];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 72; ++i) w.w1[i] = smem[i];
#pragma unroll
  for (int i = 0; i < 18; ++i) w.w2[i] = smem[72 + i];
}

__device__ __forceinline__ void load_px(const float* img, float v[10]) {
  const float2* p = reinterpret_cast<const float2*>(img);
#pragma unroll
  for (int i = 0; i < 5; ++i) { const float2 t = p[i]; v[2 * i] = t.x; v[2 * i + 1] = t.y; }
}

// x = [spec(3), timefeat(3), rays(6)]
__device__ __forceinline__ void sandwich_fwd(const DecW& w, const float v[10], const float rays[6],
                                             float x[12], float hpre[6], float out[3]) {
#pragma unroll
  for (int i = 0; i < 6; ++i) { x[i] = v[3 + i]; x[6 + i] = rays[i]; }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += w.w1[12 * j + i] * x[i];
    hpre[j] = s;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float s = v[c];
#pragma unroll
    for (int j = 0; j < 6; ++j) s += w.w2[6 * c + j] * fmaxf(hpre[j], 0.f);
    out[c] = 1.0f / (1.0f + expf(-s));
  }
}

__global__ void __launch_bounds__(kDecThreads) decode_fwd_kernel(MobgsDecodeFwd a) {
  __shared__ float sw[96];
  DecW w;
  load_w(w, a.w1, a.w2, sw);
  const size_t P = (size_t)a.width * a.height;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    float mean[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < a.K; ++k) {
      float v[10], rays[6], x[12], hpre[6], out[3];
      load_px(a.img + ((size_t)k * P + p) * 10, v);
      const float* rp = a.rays + (a.rays_per_k ? (size_t)k * 6 * P : 0) + p;
#pragma unroll
      for (int i = 0; i < 6; ++i) rays[i] = rp[i * P];
      sandwich_fwd(w, v, rays, x, hpre, out);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (a.rgb) a.rgb[((size_t)k * 3 + c) * P + p] = out[c];
        mean[c] += out[c];
      }
      if (a.depth) a.depth[(size_t)k * P + p] = v[9] / fmaxf(a.alpha[(size_t)k * P + p], kEdFloor);
    }
    if (a.mean) {
      const float inv = 1.0f / (float)a.K;
#pragma unroll
      for (int c = 0; c < 3; ++c) a.mean[c * P + p] = mean[c] * inv + kMeanEps;
    }
  }
}

__global__ void __launch_bounds__(kDecThreads) decode_bwd_kernel(MobgsDecodeBwd a) {
  __shared__ float sw[96];
  __shared__ float sred[90];
  DecW w;
  load_w(w, a.w1, a.w2, sw);
  if (threadIdx.x < 90) sred[threadIdx.x] = 0.f;
  __syncthreads();
  float gw[90];
#pragma unroll
  for (int i = 0; i < 90; ++i) gw[i] = 0.f;
  const size_t P = (size_t)a.width * a.height;
  const float invK = 1.0f / (float)a.K;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    float gm[3] = {0.f, 0.f, 0.f};
    if (a.g_mean) {
#pragma unroll
      for (int c = 0; c < 3; ++c) gm[c] = a.g_mean[c * P + p] * invK;
    }
    float gr_shared[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < a.K; ++k) {
      float v[10], rays[6], x[12], hpre[6], out[3];
      const size_t kp = (size_t)k * P + p;
      load_px(a.img + kp * 10, v);
      const float* rp = a.rays + (a.rays_per_k ? (size_t)k * 6 * P : 0) + p;
#pragma unroll
      for (int i = 0; i < 6; ++i) rays[i] = rp[i * P];
      sandwich_fwd(w, v, rays, x, hpre, out);
      float gv[10];
      float gpre[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float g = gm[c];
        if (a.g_rgb) g += a.g_rgb[((size_t)k * 3 + c) * P + p];
        gpre[c] = g * out[c] * (1.f - out[c]);
        gv[c] = gpre[c];
      }
      float ghpre[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const float h = fmaxf(hpre[j], 0.f);
        float gh = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) { gh += w.w2[6 * c + j] * gpre[c]; gw[72 + 6 * c + j] += gpre[c] * h; }
        ghpre[j] = hpre[j] > 0.f ? gh : 0.f;
      }
      float gx[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) { s += w.w1[12 * j + i] * ghpre[j]; gw[12 * j + i] += ghpre[j] * x[i]; }
        gx[i] = s;
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) gv[3 + i] = gx[i];
      // expected depth
      const float al = a.alpha[kp];
      const float den = fmaxf(al, kEdFloor);
      const float gd = a.g_depth ? a.g_depth[kp] : 0.f;
      gv[9] = gd / den;
      a.v_alpha[kp] = al > kEdFloor ? -gd * v[9] / (den * den) : 0.f;
      float2* vo = reinterpret_cast<float2*>(a.v_img + kp * 10);
#pragma unroll
      for (int i = 0; i < 5; ++i) vo[i] = make_float2(gv[2 * i], gv[2 * i + 1]);
      if (a.v_rays) {
        if (a.rays_per_k) {
#pragma unroll
          for (int i = 0; i < 6; ++i) a.v_rays[((size_t)k * 6 + i) * P + p] = gx[6 + i];
        } else {
#pragma unroll
          for (int i = 0; i < 6; ++i) gr_shared[i] += gx[6 + i];
        }
      }
    }
    if (a.v_rays && !a.rays_per_k) {
#pragma unroll
      for (int i = 0; i < 6; ++i) atomicAdd(a.v_rays + i * P + p, gr_shared[i]);
    }
  }
  // 90 weight gradients: warp shuffle -> shared -> one atomic per CTA
#pragma unroll
  for (int i = 0; i < 90; ++i) {
    const float s = warp_sum(gw[i]);
    if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(&sred[i], s);
  }
  __syncthreads();
  if (threadIdx.x < 90) {
    const float s = sred[threadIdx.x];
    if (s != 0.f) atomicAdd(threadIdx.x < 72 ? a.v_w1 + threadIdx.x : a.v_w2 + (threadIdx.x - 72), s);
  }
}

}  // namespace mobgs

using namespace mobgs;

static int decode_grid(size_t P) {
  const size_t want = (P + kDecThreads - 1) / kDecThreads;
  const size_t cap = 148 * 8;   // persistent: 8 CTAs per SM
  return (int)(want < cap ? want : cap);
}

extern "C" int mobgs_decode_fwd(const MobgsDecodeFwd* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->K >= 1 && a->width > 0 && a->height > 0, "bad extents");
  MOBGS_REQUIRE(a->img && a->alpha && a->rays && a->w1 && a->w2, "NULL input");
  const size_t P = (size_t)a->width * a->height;
  decode_fwd_kernel<<<decode_grid(P), kDecThreads, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("decode_fwd");
}

extern "C" int mobgs_decode_bwd(const MobgsDecodeBwd* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->K >= 1 && a->width > 0 && a->height > 0, "bad extents");
  MOBGS_REQUIRE(a->img && a->alpha && a->rays && a->w1 && a->w2, "NULL input");
  MOBGS_REQUIRE(a->v_img && a->v_alpha && a->v_w1 && a->v_w2, "NULL output");
  const size_t P = (size_t)a->width * a->height;
  decode_bwd_kernel<<<decode_grid(P), kDecThreads, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("decode_bwd");
}
