// f2: photometric loss of the training step, fused forward + backward:
//     photo_loss = l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))        train.py:621-628
// l1_loss = mean |x - y| (utils/loss_utils.py:233-239, mask=None); ssim = mean of the SSIM map built from
// five 11x11 Gaussian-window (sigma 1.5, zero padded, per channel) filterings of x, y, x^2, y^2, x y
// (utils/loss_utils.py:351-382).  The reference runs 5 grouped conv2d + ~20 elementwise launches forward
// and as many backward; here one forward kernel (separable window out of shared memory, register
// blocked, both reductions, and the three per-pixel partials the backward needs) and one backward kernel.
//
// Forward, per pixel q:   mu1, mu2, X2 = w*x^2, Y2 = w*y^2, XY = w*xy
//   A = 2 mu1 mu2 + C1, B = 2 (XY - mu1 mu2) + C2, C = mu1^2 + mu2^2 + C1, D = (X2 - mu1^2) + (Y2 - mu2^2) + C2
//   ssim = A B / (C D);  partials w.r.t. the filter outputs that depend on x:
//   d_mu1 = 2 mu2 (B - A) / (C D) - 2 mu1 ssim (D - C) / (C D),  d_X2 = -ssim / D,  d_XY = 2 A / (C D)
// Backward, per pixel p:  v_x = v * [ s_l1 sign(x - y) + s_ssim ( w*d_mu1 + 2 x w*d_X2 + y w*d_XY ) ]
// (the window is symmetric, so the adjoint of the filtering is the same filtering).
#include "common.cuh"

namespace mobgs {

// 32x32-pixel output tile per CTA of 256 threads.  Both filter passes are register blocked: a thread
// produces 4 consecutive outputs along the filter direction from 14 inputs it loads once (26.7 shared-
// memory loads per output pixel instead of 132 with one output per thread, which left the kernels bound
// by the LSU pipe); every input is scattered into the (up to 4) outputs whose window covers it.
constexpr int kLT = 32;                  // output tile (pixels)
constexpr int kLR = 5;                   // window radius (11 taps)
constexpr int kLH = kLT + 2 * kLR;       // 42: tile + halo
constexpr int kLS = 45;                  // halo row stride: 4 consecutive rows start 13 banks apart, the 8 lanes of a row
                                         // read 4 floats apart -> the horizontal pass is conflict free
constexpr int kLO = kLT + 1;             // filtered row stride
constexpr int kLThreads = 256;
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

// plane tile (with halo, zero outside the image) -> shared memory [kLH][kLS]
__device__ __forceinline__ void load_halo(const float* __restrict__ src, int H, int W, int y0, int x0, float* dst) {
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  for (int r = ly; r < kLH; r += kLThreads / 32) {
    const int y = y0 + r - kLR;
    const bool yin = y >= 0 && y < H;
#pragma unroll
    for (int c = lx; c < kLH; c += 32) {
      const int x = x0 + c - kLR;
      dst[r * kLS + c] = (yin && x >= 0 && x < W) ? __ldg(src + (size_t)y * W + x) : 0.f;
    }
  }
}

__global__ void __launch_bounds__(kLThreads) photo_loss_fwd_kernel(const __grid_constant__ MobgsPhotoLossFwd a) {
  __shared__ float sx[kLH * kLS], sy[kLH * kLS];
  __shared__ float sh[5][kLH * kLO];            // horizontally filtered x, y, x^2, y^2, xy
  __shared__ float red[2][kLThreads / 32];
  const int plane = blockIdx.z, y0 = blockIdx.y * kLT, x0 = blockIdx.x * kLT;
  const size_t poff = (size_t)plane * a.H * a.W;
  load_halo(a.img + poff, a.H, a.W, y0, x0, sx);
  load_halo(a.gt + poff, a.H, a.W, y0, x0, sy);
  __syncthreads();
  // horizontal pass: item = (row r, group of 4 output columns)
  for (int item = threadIdx.x; item < kLH * (kLT / 4); item += kLThreads) {
    const int r = item >> 3, c0 = (item & 7) * 4;
    float acc[5][4];
#pragma unroll
    for (int m = 0; m < 5; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[m][j] = 0.f;
#pragma unroll
    for (int i = 0; i < 14; ++i) {
      const float x = sx[r * kLS + c0 + i], y = sy[r * kLS + c0 + i];
      const float xx = x * x, yy = y * y, xy = x * y;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i - j >= 0 && i - j <= 2 * kLR) {
          const float w = a.window[i - j];
          acc[0][j] += w * x; acc[1][j] += w * y; acc[2][j] += w * xx; acc[3][j] += w * yy; acc[4][j] += w * xy;
        }
      }
    }
#pragma unroll
    for (int m = 0; m < 5; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) sh[m][r * kLO + c0 + j] = acc[m][j];
  }
  __syncthreads();
  // vertical pass: thread = (column c, group of 4 output rows)
  const int c = threadIdx.x & 31, r0 = (threadIdx.x >> 5) * 4;
  float acc[5][4];
#pragma unroll
  for (int m = 0; m < 5; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[m][j] = 0.f;
#pragma unroll
  for (int i = 0; i < 14; ++i) {
    float v[5];
#pragma unroll
    for (int m = 0; m < 5; ++m) v[m] = sh[m][(r0 + i) * kLO + c];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i - j >= 0 && i - j <= 2 * kLR) {
        const float w = a.window[i - j];
#pragma unroll
        for (int m = 0; m < 5; ++m) acc[m][j] += w * v[m];
      }
    }
  }
  float l1 = 0.f, ss_sum = 0.f;
  const int x = x0 + c;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int y = y0 + r0 + j;
    if (y < a.H && x < a.W) {
      const float mu1 = acc[0][j], mu2 = acc[1][j], x2 = acc[2][j], y2 = acc[3][j], xy = acc[4][j];
      const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
      const float A = 2.f * mu12 + kC1, B = 2.f * (xy - mu12) + kC2;
      const float C = mu1_sq + mu2_sq + kC1, D = (x2 - mu1_sq) + (y2 - mu2_sq) + kC2;
      const float inv_cd = 1.f / (C * D);
      const float ss = A * B * inv_cd;
      ss_sum += ss;
      l1 += fabsf(sx[(r0 + j + kLR) * kLS + c + kLR] - sy[(r0 + j + kLR) * kLS + c + kLR]);
      if (a.d_mu1) {
        const size_t o = poff + (size_t)y * a.W + x;
        a.d_mu1[o] = 2.f * mu2 * (B - A) * inv_cd - 2.f * mu1 * ss * (D - C) * inv_cd;
        a.d_x2[o] = -ss / D;
        a.d_xy[o] = 2.f * A * inv_cd;
      }
    }
  }
  l1 = warp_sum(l1);
  ss_sum = warp_sum(ss_sum);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = l1; red[1][threadIdx.x >> 5] = ss_sum; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kLThreads / 32; ++w) s += (double)red[threadIdx.x][w];
    atomicAdd(a.sums + threadIdx.x, s);
  }
}

__global__ void __launch_bounds__(kLThreads) photo_loss_bwd_kernel(const __grid_constant__ MobgsPhotoLossBwd a) {
  __shared__ float sm[3][kLH * kLS];            // d_mu1, d_x2, d_xy with halo
  __shared__ float sh[3][kLH * kLO];
  const int plane = blockIdx.z, y0 = blockIdx.y * kLT, x0 = blockIdx.x * kLT;
  const size_t poff = (size_t)plane * a.H * a.W;
  load_halo(a.d_mu1 + poff, a.H, a.W, y0, x0, sm[0]);
  load_halo(a.d_x2 + poff, a.H, a.W, y0, x0, sm[1]);
  load_halo(a.d_xy + poff, a.H, a.W, y0, x0, sm[2]);
  __syncthreads();
  for (int item = threadIdx.x; item < kLH * (kLT / 4); item += kLThreads) {
    const int r = item >> 3, c0 = (item & 7) * 4;
    float acc[3][4];
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[m][j] = 0.f;
#pragma unroll
    for (int i = 0; i < 14; ++i) {
      float v[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) v[m] = sm[m][r * kLS + c0 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i - j >= 0 && i - j <= 2 * kLR) {
          const float w = a.window[i - j];
#pragma unroll
          for (int m = 0; m < 3; ++m) acc[m][j] += w * v[m];
        }
      }
    }
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) sh[m][r * kLO + c0 + j] = acc[m][j];
  }
  __syncthreads();
  const int c = threadIdx.x & 31, r0 = (threadIdx.x >> 5) * 4;
  float acc[3][4];
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[m][j] = 0.f;
#pragma unroll
  for (int i = 0; i < 14; ++i) {
    float v[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) v[m] = sh[m][(r0 + i) * kLO + c];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i - j >= 0 && i - j <= 2 * kLR) {
        const float w = a.window[i - j];
#pragma unroll
        for (int m = 0; m < 3; ++m) acc[m][j] += w * v[m];
      }
    }
  }
  const int x = x0 + c;
  const float v = a.v_loss ? __ldg(a.v_loss) : 1.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int y = y0 + r0 + j;
    if (y < a.H && x < a.W) {
      const size_t o = poff + (size_t)y * a.W + x;
      const float xv = __ldg(a.img + o), yv = __ldg(a.gt + o);
      const float d = xv - yv;
      const float sgn = (d > 0.f ? 1.f : 0.f) - (d < 0.f ? 1.f : 0.f);
      a.v_img[o] = v * (a.scale_l1 * sgn + a.scale_ssim * (acc[0][j] + 2.f * xv * acc[1][j] + yv * acc[2][j]));
    }
  }
}

}  // namespace mobgs

using namespace mobgs;

extern "C" int mobgs_photo_loss_fwd(const MobgsPhotoLossFwd* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->planes >= 1 && a->planes <= 65535 && a->H > 0 && a->W > 0, "bad extents");
  MOBGS_REQUIRE(a->img && a->gt && a->sums, "NULL pointer");
  MOBGS_REQUIRE((a->d_mu1 == nullptr) == (a->d_x2 == nullptr) && (a->d_mu1 == nullptr) == (a->d_xy == nullptr),
                "d_mu1 / d_x2 / d_xy must be all set or all NULL");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(a->sums, 0, 2 * sizeof(double), s);
  const dim3 grid((a->W + kLT - 1) / kLT, (a->H + kLT - 1) / kLT, a->planes);
  photo_loss_fwd_kernel<<<grid, kLThreads, 0, s>>>(*a);
  return check_launch("photo_loss_fwd");
}

extern "C" int mobgs_photo_loss_bwd(const MobgsPhotoLossBwd* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->planes >= 1 && a->planes <= 65535 && a->H > 0 && a->W > 0, "bad extents");
  MOBGS_REQUIRE(a->img && a->gt && a->d_mu1 && a->d_x2 && a->d_xy && a->v_img, "NULL pointer");
  const dim3 grid((a->W + kLT - 1) / kLT, (a->H + kLT - 1) / kLT, a->planes);
  photo_loss_bwd_kernel<<<grid, kLThreads, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("photo_loss_bwd");
}
