// Per-pixel epilogue: expected-depth division + Sandwich RGB decoder + K-sub-frame mean, forward
// and VJP.  Replaces two cuDNN 1x1 convolutions (TF32 by default on this
// hardware), three elementwise launches and a stack+mean per decode with one fp32 kernel.
// One thread = one pixel, looping over the K sub-frames; HBM-bound (reads 44 + 24 B, writes 16 B
// per pixel and sub-frame).  The 90 decoder-weight gradients are reduced warp -> CTA -> one
// atomicAdd per CTA of a persistent grid.
#include "common.cuh"
#include "decode_math.cuh"

namespace mobgs {

constexpr int kDecThreads = 128;
// img10 is channels-last (40 B per pixel): a CTA moves its 256-pixel block through shared memory
// with coalesced 8-byte accesses instead of letting every thread walk its own 40-byte record.
__device__ __forceinline__ void stage_in(float* simg, const float* gsrc, int npx) {
  const float2* src = reinterpret_cast<const float2*>(gsrc);
  float2* dst = reinterpret_cast<float2*>(simg);
  for (int i = threadIdx.x; i < npx * 5; i += kDecThreads) dst[i] = __ldg(src + i);
}
__device__ __forceinline__ void stage_out(float* gdst, const float* simg, int npx) {
  const float2* src = reinterpret_cast<const float2*>(simg);
  float2* dst = reinterpret_cast<float2*>(gdst);
  for (int i = threadIdx.x; i < npx * 5; i += kDecThreads) dst[i] = src[i];
}

// Work item = (sub-frame k, block of kDecThreads pixels); a persistent grid strides over K * nblk
// items so that the sub-frames of a pixel are independent pieces of work in flight at once.
__global__ void __launch_bounds__(kDecThreads, 4) decode_fwd_kernel(MobgsDecodeFwd a) {
  __shared__ __align__(16) float sw[96];
  __shared__ __align__(16) float simg[kDecThreads * 10];
  DecW w;
  load_w(w, a.w1, a.w2, sw);
  const size_t P = (size_t)a.width * a.height;
  const size_t nblk = (P + kDecThreads - 1) / kDecThreads;
  for (size_t item = blockIdx.x; item < nblk * a.K; item += gridDim.x) {
    const int k = (int)(item / nblk);
    const size_t pb = item - (size_t)k * nblk;
    const size_t p0 = pb * kDecThreads, p = p0 + threadIdx.x;
    const int npx = (int)min((size_t)kDecThreads, P - p0);
    const bool valid = threadIdx.x < npx;
    float rays[6] = {0, 0, 0, 0, 0, 0}, al = 1.f;
    if (valid) {     // issue the planar loads before the staging barrier so they overlap it
      const float* rp = a.rays + (a.rays_per_k ? (size_t)k * 6 * P : 0) + p;
#pragma unroll
      for (int i = 0; i < 6; ++i) rays[i] = __ldg(rp + i * P);
      al = __ldg(a.alpha + (size_t)k * P + p);
    }
    __syncthreads();
    stage_in(simg, a.img + ((size_t)k * P + p0) * 10, npx);
    __syncthreads();
    if (!valid) continue;
    float v[10], x[12], hpre[6], out[3];
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = simg[threadIdx.x * 10 + i];
    sandwich_fwd(w, v, rays, x, hpre, out);
#pragma unroll
    for (int c = 0; c < 3; ++c) a.rgb[((size_t)k * 3 + c) * P + p] = out[c];
    if (a.depth) a.depth[(size_t)k * P + p] = v[9] / fmaxf(al, kEdFloor);
  }
}

// blur model: mean over the K decoded sub-frames + 1e-10 (train.py:540-541)
__global__ void __launch_bounds__(256) subframe_mean_kernel(const float* __restrict__ rgb, float* __restrict__ mean,
                                                            int K, size_t n) {
  const float inv = 1.0f / (float)K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += __ldg(rgb + (size_t)k * n + i);
    mean[i] = s * inv + kMeanEps;
  }
}

__global__ void __launch_bounds__(kDecThreads, 3) decode_bwd_kernel(MobgsDecodeBwd a) {
  __shared__ __align__(16) float sw[96];
  __shared__ float sred[90];
  __shared__ __align__(16) float simg[kDecThreads * 10];
  DecW w;
  load_w(w, a.w1, a.w2, sw);
  if (threadIdx.x < 90) sred[threadIdx.x] = 0.f;
  float gw[90];
#pragma unroll
  for (int i = 0; i < 90; ++i) gw[i] = 0.f;
  const size_t P = (size_t)a.width * a.height;
  const size_t nblk = (P + kDecThreads - 1) / kDecThreads;
  const float invK = 1.0f / (float)a.K;
  for (size_t item = blockIdx.x; item < nblk * a.K; item += gridDim.x) {
    const int k = (int)(item / nblk);
    const size_t pb = item - (size_t)k * nblk;
    const size_t p0 = pb * kDecThreads, p = p0 + threadIdx.x;
    const int npx = (int)min((size_t)kDecThreads, P - p0);
    const bool valid = threadIdx.x < npx;
    const size_t kp0 = (size_t)k * P + p0, kp = kp0 + threadIdx.x;
    float rays[6] = {0, 0, 0, 0, 0, 0}, al = 1.f, gd = 0.f, g_in[3] = {0.f, 0.f, 0.f};
    if (valid) {
      const float* rp = a.rays + (a.rays_per_k ? (size_t)k * 6 * P : 0) + p;
#pragma unroll
      for (int i = 0; i < 6; ++i) rays[i] = __ldg(rp + i * P);
      al = __ldg(a.alpha + kp);
      if (a.g_depth) gd = __ldg(a.g_depth + kp);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (a.g_mean) g_in[c] = __ldg(a.g_mean + c * P + p) * invK;
        if (a.g_rgb) g_in[c] += __ldg(a.g_rgb + ((size_t)k * 3 + c) * P + p);
      }
    }
    __syncthreads();
    stage_in(simg, a.img + kp0 * 10, npx);
    __syncthreads();
    float gv[10];
    if (valid) {
      float v[10], x[12], hpre[6], out[3];
#pragma unroll
      for (int i = 0; i < 10; ++i) v[i] = simg[threadIdx.x * 10 + i];
      sandwich_fwd(w, v, rays, x, hpre, out);
      asm volatile("" ::: "memory");   // do not keep the forward's 90 weights live in registers
      float gpre[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        gpre[c] = g_in[c] * out[c] * (1.f - out[c]);
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
      asm volatile("" ::: "memory");
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
      const float den = fmaxf(al, kEdFloor);
      gv[9] = gd / den;
      a.v_alpha[kp] = al > kEdFloor ? -gd * v[9] / (den * den) : 0.f;
      if (a.v_rays) {
        if (a.rays_per_k) {
#pragma unroll
          for (int i = 0; i < 6; ++i) a.v_rays[((size_t)k * 6 + i) * P + p] = gx[6 + i];
        } else {
#pragma unroll
          for (int i = 0; i < 6; ++i) atomicAdd(a.v_rays + i * P + p, gx[6 + i]);
        }
      }
    }
    __syncthreads();            // everyone has read its pixel: reuse the buffer for the gradient
    if (valid) {
#pragma unroll
      for (int i = 0; i < 10; ++i) simg[threadIdx.x * 10 + i] = gv[i];
    }
    __syncthreads();
    stage_out(a.v_img + kp0 * 10, simg, npx);
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

static int decode_grid(size_t items, int per_sm) {
  const size_t cap = (size_t)148 * per_sm;   // persistent grid
  return (int)(items < cap ? items : cap);
}

extern "C" int mobgs_decode_fwd(const MobgsDecodeFwd* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->K >= 1 && a->width > 0 && a->height > 0, "bad extents");
  MOBGS_REQUIRE(a->img && a->alpha && a->rays && a->w1 && a->w2, "NULL input");
  MOBGS_REQUIRE(a->rgb, "rgb output must not be NULL (the mean is reduced from it)");
  const size_t P = (size_t)a->width * a->height;
  const size_t nblk = (P + kDecThreads - 1) / kDecThreads;
  decode_fwd_kernel<<<decode_grid(nblk * a->K, 16), kDecThreads, 0, (cudaStream_t)stream>>>(*a);
  if (a->mean) {
    const size_t n = 3 * P;
    subframe_mean_kernel<<<decode_grid((n + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(a->rgb, a->mean, a->K, n);
  }
  return check_launch("decode_fwd");
}

extern "C" int mobgs_decode_bwd(const MobgsDecodeBwd* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->K >= 1 && a->width > 0 && a->height > 0, "bad extents");
  MOBGS_REQUIRE(a->img && a->alpha && a->rays && a->w1 && a->w2, "NULL input");
  MOBGS_REQUIRE(a->v_img && a->v_alpha && a->v_w1 && a->v_w2, "NULL output");
  const size_t P = (size_t)a->width * a->height;
  const size_t nblk = (P + kDecThreads - 1) / kDecThreads;
  decode_bwd_kernel<<<decode_grid(nblk * a->K, 9), kDecThreads, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("decode_bwd");
}

extern "C" int mobgs_subframe_mean(const float* rgb, float* mean, int32_t K, int64_t n, void* stream) {
  MOBGS_REQUIRE(rgb && mean && K >= 1 && n >= 0, "bad arguments");
  if (n == 0) return MOBGS_OK;
  subframe_mean_kernel<<<decode_grid((size_t)(n + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(rgb, mean, K, (size_t)n);
  return check_launch("subframe_mean");
}
