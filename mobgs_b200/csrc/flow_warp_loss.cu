// f1, second half: the flow-warp loss of train.py:656-676, forward + backward fused.
//   exp2mid:  warped1 = grid_sample(ori_image (expanded over the K exposures), norm(exp2mid_coord))
//             term1   = l1_loss(warped1, latent_img, mask = latent_alpha)
//   mid2exp:  warped2 = grid_sample(latent_img, norm(mid2exp_coord))
//             term2   = l1_loss(warped2, ori_image (expanded), mask = d_alpha (expanded))
//   flow_loss = lambda_flow_loss * (term1 + term2)            (the caller applies lambda)
// with  norm(c) = 2 c / (size - 1) - 1 (train.py:659-662, 667-670),  F.grid_sample(mode='bilinear',
// padding_mode='border') at its default align_corners=False, i.e. source position
// s = ((norm + 1) size - 1) / 2 = c size / (size - 1) - 0.5 clipped to [0, size - 1] (no coordinate
// gradient where clipped), and the masked l1_loss of utils/loss_utils.py:233-237:
// sum |(a - b) mask| / (sum(mask expanded to the 3 channels) + 1e-8).
// The reference runs 2 grid_sample + ~30 elementwise launches forward and as many backward over
// [B*K,3,H,W] tensors; here one forward kernel (four sums) and one backward kernel.  One thread = one
// (view b, pixel), looping over the K exposures.
#include "common.cuh"

namespace mobgs {

constexpr int kFwThreads = 256;

struct Bilin {
  int x0, y0;            // top-left neighbour
  float wx, wy;          // weights of the right / lower neighbours
  float gx, gy;          // d s / d c (0 where the border clip is active)
};

__device__ __forceinline__ Bilin bilin_setup(float cx, float cy, int W, int H) {
  Bilin b;
  // 2 c / (size - 1) - 1 as the reference writes it, then grid_sample's align_corners=False unnormalisation
  const float nx = 2.f * (cx / (float)(W - 1)) - 1.f, ny = 2.f * (cy / (float)(H - 1)) - 1.f;
  float sx = ((nx + 1.f) * (float)W - 1.f) * 0.5f, sy = ((ny + 1.f) * (float)H - 1.f) * 0.5f;
  b.gx = (sx > 0.f && sx < (float)(W - 1)) ? (float)W / (float)(W - 1) : 0.f;     // clip_coordinates_set_grad
  b.gy = (sy > 0.f && sy < (float)(H - 1)) ? (float)H / (float)(H - 1) : 0.f;
  sx = fminf(fmaxf(sx, 0.f), (float)(W - 1));
  sy = fminf(fmaxf(sy, 0.f), (float)(H - 1));
  const float fx = floorf(sx), fy = floorf(sy);
  b.x0 = (int)fx; b.y0 = (int)fy;
  b.wx = sx - fx; b.wy = sy - fy;
  return b;
}

// value and (d/dsx, d/dsy) of the bilinear sample of one plane; neighbours outside the image contribute
// nothing (they only occur with weight 0, at the clipped border)
__device__ __forceinline__ float bilin_sample(const float* __restrict__ plane, const Bilin& b, int W, int H,
                                              float& dsx, float& dsy) {
  const bool xin = b.x0 + 1 < W, yin = b.y0 + 1 < H;
  const float v00 = __ldg(plane + (size_t)b.y0 * W + b.x0);
  const float v01 = xin ? __ldg(plane + (size_t)b.y0 * W + b.x0 + 1) : 0.f;
  const float v10 = yin ? __ldg(plane + (size_t)(b.y0 + 1) * W + b.x0) : 0.f;
  const float v11 = (xin && yin) ? __ldg(plane + (size_t)(b.y0 + 1) * W + b.x0 + 1) : 0.f;
  const float ux = 1.f - b.wx, uy = 1.f - b.wy;
  dsx = (v01 - v00) * uy + (v11 - v10) * b.wy;
  dsy = (v10 - v00) * ux + (v11 - v01) * b.wx;
  return v00 * ux * uy + v01 * b.wx * uy + v10 * ux * b.wy + v11 * b.wx * b.wy;
}

__device__ __forceinline__ float sgn(float d) { return (d > 0.f ? 1.f : 0.f) - (d < 0.f ? 1.f : 0.f); }

__global__ void __launch_bounds__(kFwThreads) flow_warp_fwd_kernel(const __grid_constant__ MobgsFlowWarp a) {
  __shared__ double red[4][kFwThreads / 32];
  const int P = a.H * a.W;
  const int b = blockIdx.y;
  const int p = blockIdx.x * kFwThreads + threadIdx.x;
  float s[4] = {0.f, 0.f, 0.f, 0.f};      // num1, den1, num2, den2 of this pixel over the K exposures
  if (p < P) {
    const float* ori = a.ori + (size_t)b * 3 * P;
    const float m2 = a.d_alpha[(size_t)b * P + p];
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = ori[(size_t)c * P + p];
    for (int k = 0; k < a.K; ++k) {
      const size_t bk = (size_t)b * a.K + k;
      const float* lat = a.latent + bk * 3 * P;
      const float2 c1 = *reinterpret_cast<const float2*>(a.exp2mid + (bk * P + p) * 2);
      const float2 c2 = *reinterpret_cast<const float2*>(a.mid2exp + (bk * P + p) * 2);
      const Bilin b1 = bilin_setup(c1.x, c1.y, a.W, a.H), b2 = bilin_setup(c2.x, c2.y, a.W, a.H);
      const float m1 = a.latent_alpha[bk * P + p];
      float dx, dy;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float w1 = bilin_sample(ori + (size_t)c * P, b1, a.W, a.H, dx, dy);
        s[0] += fabsf((w1 - lat[(size_t)c * P + p]) * m1);
        const float w2 = bilin_sample(lat + (size_t)c * P, b2, a.W, a.H, dx, dy);
        s[2] += fabsf((w2 - o[c]) * m2);
      }
      s[1] += 3.f * m1;
      s[3] += 3.f * m2;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float w = warp_sum(s[i]);
    if ((threadIdx.x & 31) == 0) red[i][threadIdx.x >> 5] = (double)w;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kFwThreads / 32; ++w) t += red[threadIdx.x][w];
    atomicAdd(a.sums + threadIdx.x, t);
  }
}

__global__ void __launch_bounds__(kFwThreads) flow_warp_bwd_kernel(const __grid_constant__ MobgsFlowWarp a) {
  const int P = a.H * a.W;
  const int b = blockIdx.y;
  const int p = blockIdx.x * kFwThreads + threadIdx.x;
  if (p >= P) return;
  const float g = a.v_loss ? __ldg(a.v_loss) : 1.f;
  const double den1 = a.sums[1] + 1e-8, den2 = a.sums[3] + 1e-8;
  const float i1 = (float)(1.0 / den1), i2 = (float)(1.0 / den2);
  const float q1 = (float)(3.0 * a.sums[0] / (den1 * den1)), q2 = (float)(3.0 * a.sums[2] / (den2 * den2));   // d term / d mask via the denominator
  const float* ori = a.ori + (size_t)b * 3 * P;
  const float m2 = a.d_alpha[(size_t)b * P + p];
  float o[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = ori[(size_t)c * P + p];
  float v_m2 = 0.f;
  float v_o[3] = {0.f, 0.f, 0.f};          // d loss / d ori[b, :, p] as the L1 target of the mid2exp term
  float* v_ori = a.v_ori ? a.v_ori + (size_t)b * 3 * P : nullptr;
  for (int k = 0; k < a.K; ++k) {
    const size_t bk = (size_t)b * a.K + k;
    const float* lat = a.latent + bk * 3 * P;
    float* v_lat = a.v_latent + bk * 3 * P;
    const float2 c1 = *reinterpret_cast<const float2*>(a.exp2mid + (bk * P + p) * 2);
    const float2 c2 = *reinterpret_cast<const float2*>(a.mid2exp + (bk * P + p) * 2);
    const Bilin b1 = bilin_setup(c1.x, c1.y, a.W, a.H), b2 = bilin_setup(c2.x, c2.y, a.W, a.H);
    const float m1 = a.latent_alpha[bk * P + p];
    float v1x = 0.f, v1y = 0.f, v2x = 0.f, v2y = 0.f, v_m1 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float dx, dy;
      // exp2mid: source = ori image (data), target = latent image, mask = latent alpha
      const float w1 = bilin_sample(ori + (size_t)c * P, b1, a.W, a.H, dx, dy);
      const float r1 = w1 - lat[(size_t)c * P + p];
      const float s1 = sgn(r1 * m1);
      const float gw1 = g * s1 * m1 * i1;                 // d loss / d warped1[c]
      v1x += gw1 * dx; v1y += gw1 * dy;
      if (gw1 != 0.f) {
        atomicAdd(v_lat + (size_t)c * P + p, -gw1);
        if (v_ori) {                                       // the sampled source: bilinear scatter
          float* vp = v_ori + (size_t)c * P;
          const bool xin = b1.x0 + 1 < a.W, yin = b1.y0 + 1 < a.H;
          const float ux = 1.f - b1.wx, uy = 1.f - b1.wy;
          atomicAdd(vp + (size_t)b1.y0 * a.W + b1.x0, gw1 * ux * uy);
          if (xin) atomicAdd(vp + (size_t)b1.y0 * a.W + b1.x0 + 1, gw1 * b1.wx * uy);
          if (yin) atomicAdd(vp + (size_t)(b1.y0 + 1) * a.W + b1.x0, gw1 * ux * b1.wy);
          if (xin && yin) atomicAdd(vp + (size_t)(b1.y0 + 1) * a.W + b1.x0 + 1, gw1 * b1.wx * b1.wy);
        }
      }
      v_m1 += g * s1 * r1 * i1;
      // mid2exp: source = latent image (scatter), target = ori image, mask = dynamic alpha
      const float w2 = bilin_sample(lat + (size_t)c * P, b2, a.W, a.H, dx, dy);
      const float r2 = w2 - o[c];
      const float s2 = sgn(r2 * m2);
      const float gw2 = g * s2 * m2 * i2;                 // d loss / d warped2[c]
      v2x += gw2 * dx; v2y += gw2 * dy;
      v_m2 += g * s2 * r2 * i2;
      v_o[c] -= gw2;
      if (gw2 != 0.f) {
        float* vp = v_lat + (size_t)c * P;
        const bool xin = b2.x0 + 1 < a.W, yin = b2.y0 + 1 < a.H;
        const float ux = 1.f - b2.wx, uy = 1.f - b2.wy;
        atomicAdd(vp + (size_t)b2.y0 * a.W + b2.x0, gw2 * ux * uy);
        if (xin) atomicAdd(vp + (size_t)b2.y0 * a.W + b2.x0 + 1, gw2 * b2.wx * uy);
        if (yin) atomicAdd(vp + (size_t)(b2.y0 + 1) * a.W + b2.x0, gw2 * ux * b2.wy);
        if (xin && yin) atomicAdd(vp + (size_t)(b2.y0 + 1) * a.W + b2.x0 + 1, gw2 * b2.wx * b2.wy);
      }
    }
    *reinterpret_cast<float2*>(a.v_exp2mid + (bk * P + p) * 2) = make_float2(v1x * b1.gx, v1y * b1.gy);
    *reinterpret_cast<float2*>(a.v_mid2exp + (bk * P + p) * 2) = make_float2(v2x * b2.gx, v2y * b2.gy);
    a.v_latent_alpha[bk * P + p] = v_m1 - g * q1;
    v_m2 -= g * q2;
  }
  a.v_d_alpha[(size_t)b * P + p] = v_m2;
  if (v_ori) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (v_o[c] != 0.f) atomicAdd(v_ori + (size_t)c * P + p, v_o[c]);
  }
}

}  // namespace mobgs

using namespace mobgs;

static int check_flow_warp(const MobgsFlowWarp* a) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->B >= 1 && a->B <= 65535 && a->K >= 1 && a->H > 1 && a->W > 1, "bad extents");
  MOBGS_REQUIRE(a->ori && a->latent && a->exp2mid && a->mid2exp && a->latent_alpha && a->d_alpha && a->sums,
                "NULL pointer");
  return MOBGS_OK;
}

extern "C" int mobgs_flow_warp_loss_fwd(const MobgsFlowWarp* a, void* stream) {
  if (int e = check_flow_warp(a)) return e;
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(a->sums, 0, 4 * sizeof(double), s);
  const dim3 grid((a->H * a->W + kFwThreads - 1) / kFwThreads, a->B);
  flow_warp_fwd_kernel<<<grid, kFwThreads, 0, s>>>(*a);
  return check_launch("flow_warp_loss_fwd");
}

extern "C" int mobgs_flow_warp_loss_bwd(const MobgsFlowWarp* a, void* stream) {
  if (int e = check_flow_warp(a)) return e;
  MOBGS_REQUIRE(a->v_latent && a->v_exp2mid && a->v_mid2exp && a->v_latent_alpha && a->v_d_alpha, "NULL gradient pointer");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(a->v_latent, 0, sizeof(float) * (size_t)a->B * a->K * 3 * a->H * a->W, s);
  if (a->v_ori) cudaMemsetAsync(a->v_ori, 0, sizeof(float) * (size_t)a->B * 3 * a->H * a->W, s);
  const dim3 grid((a->H * a->W + kFwThreads - 1) / kFwThreads, a->B);
  flow_warp_bwd_kernel<<<grid, kFwThreads, 0, s>>>(*a);
  return check_launch("flow_warp_loss_bwd");
}
