// f2: photometric loss of the training step, fused forward + backward:
//     photo_loss = l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))        train.py:621-628
// l1_loss = mean |x - y| (utils/loss_utils.py:233-239, mask=None); ssim = mean of the SSIM map built from
// five 11x11 Gaussian-window (sigma 1.5, zero padded, per channel) filterings of x, y, x^2, y^2, x y
// (utils/loss_utils.py:351-382).  The reference runs 5 grouped conv2d + ~20 elementwise launches forward
// and as many backward; here one forward kernel (separable window out of shared memory, both
// reductions, and the three per-pixel partials the backward needs) and one backward kernel.
//
// Forward, per pixel q:   mu1, mu2, X2 = w*x^2, Y2 = w*y^2, XY = w*xy
//   A = 2 mu1 mu2 + C1, B = 2 (XY - mu1 mu2) + C2, C = mu1^2 + mu2^2 + C1, D = (X2 - mu1^2) + (Y2 - mu2^2) + C2
//   ssim = A B / (C D);  partials w.r.t. the filter outputs that depend on x:
//   d_mu1 = 2 mu2 (B - A) / (C D) - 2 mu1 ssim (D - C) / (C D),  d_X2 = -ssim / D,  d_XY = 2 A / (C D)
// Backward, per pixel p:  v_x = v * [ s_l1 sign(x - y) + s_ssim ( w*d_mu1 + 2 x w*d_X2 + y w*d_XY ) ]
// (the window is symmetric, so the adjoint of the filtering is the same filtering).
#include "common.cuh"

namespace mobgs {

constexpr int kLT = 16;                  // output tile (pixels)
constexpr int kLR = 5;                   // window radius (11 taps)
constexpr int kLH = kLT + 2 * kLR;       // 26: tile + halo
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

// plane tile (with halo, zero outside the image) -> shared memory
__device__ __forceinline__ void load_halo(const float* __restrict__ src, int H, int W, int y0, int x0,
                                          float (*dst)[kLH + 1]) {
  for (int i = threadIdx.x; i < kLH * kLH; i += kLT * kLT) {
    const int r = i / kLH, c = i - r * kLH;
    const int y = y0 + r - kLR, x = x0 + c - kLR;
    dst[r][c] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(src + (size_t)y * W + x) : 0.f;
  }
}

__global__ void __launch_bounds__(kLT * kLT) photo_loss_fwd_kernel(const __grid_constant__ MobgsPhotoLossFwd a) {
  __shared__ float sx[kLH][kLH + 1], sy[kLH][kLH + 1];
  __shared__ float sh[5][kLH][kLT + 1];           // horizontally filtered x, y, x^2, y^2, xy
  __shared__ float red[2][kLT * kLT / 32];
  const int plane = blockIdx.z, y0 = blockIdx.y * kLT, x0 = blockIdx.x * kLT;
  const size_t poff = (size_t)plane * a.H * a.W;
  load_halo(a.img + poff, a.H, a.W, y0, x0, sx);
  load_halo(a.gt + poff, a.H, a.W, y0, x0, sy);
  __syncthreads();
  for (int i = threadIdx.x; i < kLH * kLT; i += kLT * kLT) {
    const int r = i / kLT, c = i - r * kLT;
    float hx = 0.f, hy = 0.f, hxx = 0.f, hyy = 0.f, hxy = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * kLR + 1; ++k) {
      const float w = a.window[k], x = sx[r][c + k], y = sy[r][c + k];
      hx += w * x; hy += w * y; hxx += w * x * x; hyy += w * y * y; hxy += w * x * y;
    }
    sh[0][r][c] = hx; sh[1][r][c] = hy; sh[2][r][c] = hxx; sh[3][r][c] = hyy; sh[4][r][c] = hxy;
  }
  __syncthreads();
  const int ty = threadIdx.x / kLT, tx = threadIdx.x - ty * kLT;
  const int y = y0 + ty, x = x0 + tx;
  float l1 = 0.f, ss = 0.f;
  if (y < a.H && x < a.W) {
    float mu1 = 0.f, mu2 = 0.f, x2 = 0.f, y2 = 0.f, xy = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * kLR + 1; ++k) {
      const float w = a.window[k];
      mu1 += w * sh[0][ty + k][tx]; mu2 += w * sh[1][ty + k][tx];
      x2 += w * sh[2][ty + k][tx]; y2 += w * sh[3][ty + k][tx]; xy += w * sh[4][ty + k][tx];
    }
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float A = 2.f * mu12 + kC1, B = 2.f * (xy - mu12) + kC2;
    const float C = mu1_sq + mu2_sq + kC1, D = (x2 - mu1_sq) + (y2 - mu2_sq) + kC2;
    const float inv_cd = 1.f / (C * D);
    ss = A * B * inv_cd;
    l1 = fabsf(sx[ty + kLR][tx + kLR] - sy[ty + kLR][tx + kLR]);
    if (a.d_mu1) {
      const size_t o = poff + (size_t)y * a.W + x;
      a.d_mu1[o] = 2.f * mu2 * (B - A) * inv_cd - 2.f * mu1 * ss * (D - C) * inv_cd;
      a.d_x2[o] = -ss / D;
      a.d_xy[o] = 2.f * A * inv_cd;
    }
  }
  l1 = warp_sum(l1);
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = l1; red[1][threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kLT * kLT / 32; ++w) s += (double)red[threadIdx.x][w];
    atomicAdd(a.sums + threadIdx.x, s);
  }
}

__global__ void __launch_bounds__(kLT * kLT) photo_loss_bwd_kernel(const __grid_constant__ MobgsPhotoLossBwd a) {
  __shared__ float sm[3][kLH][kLH + 1];           // d_mu1, d_x2, d_xy with halo
  __shared__ float sh[3][kLH][kLT + 1];
  const int plane = blockIdx.z, y0 = blockIdx.y * kLT, x0 = blockIdx.x * kLT;
  const size_t poff = (size_t)plane * a.H * a.W;
  load_halo(a.d_mu1 + poff, a.H, a.W, y0, x0, sm[0]);
  load_halo(a.d_x2 + poff, a.H, a.W, y0, x0, sm[1]);
  load_halo(a.d_xy + poff, a.H, a.W, y0, x0, sm[2]);
  __syncthreads();
  for (int i = threadIdx.x; i < kLH * kLT; i += kLT * kLT) {
    const int r = i / kLT, c = i - r * kLT;
    float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * kLR + 1; ++k) {
      const float w = a.window[k];
      h0 += w * sm[0][r][c + k]; h1 += w * sm[1][r][c + k]; h2 += w * sm[2][r][c + k];
    }
    sh[0][r][c] = h0; sh[1][r][c] = h1; sh[2][r][c] = h2;
  }
  __syncthreads();
  const int ty = threadIdx.x / kLT, tx = threadIdx.x - ty * kLT;
  const int y = y0 + ty, x = x0 + tx;
  if (y >= a.H || x >= a.W) return;
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
  for (int k = 0; k < 2 * kLR + 1; ++k) {
    const float w = a.window[k];
    g0 += w * sh[0][ty + k][tx]; g1 += w * sh[1][ty + k][tx]; g2 += w * sh[2][ty + k][tx];
  }
  const size_t o = poff + (size_t)y * a.W + x;
  const float xv = __ldg(a.img + o), yv = __ldg(a.gt + o);
  const float d = xv - yv;
  const float sgn = (d > 0.f ? 1.f : 0.f) - (d < 0.f ? 1.f : 0.f);
  const float v = a.v_loss ? __ldg(a.v_loss) : 1.f;
  a.v_img[o] = v * (a.scale_l1 * sgn + a.scale_ssim * (g0 + 2.f * xv * g1 + yv * g2));
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
  photo_loss_fwd_kernel<<<grid, kLT * kLT, 0, s>>>(*a);
  return check_launch("photo_loss_fwd");
}

extern "C" int mobgs_photo_loss_bwd(const MobgsPhotoLossBwd* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->planes >= 1 && a->planes <= 65535 && a->H > 0 && a->W > 0, "bad extents");
  MOBGS_REQUIRE(a->img && a->gt && a->d_mu1 && a->d_x2 && a->d_xy && a->v_img, "NULL pointer");
  const dim3 grid((a->W + kLT - 1) / kLT, (a->H + kLT - 1) / kLT, a->planes);
  photo_loss_bwd_kernel<<<grid, kLT * kLT, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("photo_loss_bwd");
}
