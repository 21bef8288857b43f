// f2 (remainder): the regularisers of train.py:651-655 in one pass,
//     reg = w_depth * l1_loss(depth, gt_depth) + w_mask * (entropy_loss(d_alpha) + sparsity_loss(d_alpha))
// with l1_loss = mean |a - b| (utils/loss_utils.py:233-239, mask=None), entropy_loss = -sum(a log(a + 1e-6) +
// (1 - a) log(1 - a + 1e-6)) (:264-276) and sparsity_loss = sum(a^2) (:285-295); train.py uses w_depth = 0.2,
// w_mask = 1e-7.  The kernel writes the three sums (double) and the UNSCALED gradient maps
// sign(depth - gt) and d(entropy + sparsity)/d alpha, so the backward is a scalar multiply.
// Streaming, HBM-bound: 12 B read + 8 B written per pixel.
#include "common.cuh"

namespace mobgs {

constexpr int kRegThreads = 256;

__global__ void __launch_bounds__(kRegThreads) reg_loss_kernel(const __grid_constant__ MobgsRegLoss a) {
  __shared__ double red[3][kRegThreads / 32];
  float s[3] = {0.f, 0.f, 0.f};
  const int64_t stride = (int64_t)gridDim.x * kRegThreads;
  for (int64_t i = (int64_t)blockIdx.x * kRegThreads + threadIdx.x; i < a.n_depth; i += stride) {
    const float d = a.depth[i] - a.gt_depth[i];
    s[0] += fabsf(d);
    if (a.g_depth) a.g_depth[i] = (d > 0.f ? 1.f : 0.f) - (d < 0.f ? 1.f : 0.f);
  }
  for (int64_t i = (int64_t)blockIdx.x * kRegThreads + threadIdx.x; i < a.n_alpha; i += stride) {
    const float al = a.alpha[i];
    const float p = al + 1e-6f, q = 1.f - al + 1e-6f;
    const float lp = logf(p), lq = logf(q);
    s[1] -= al * lp + (1.f - al) * lq;
    s[2] += al * al;
    if (a.g_alpha) a.g_alpha[i] = -(lp + al / p - lq - (1.f - al) / q) + 2.f * al;
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float w = warp_sum(s[j]);
    if ((threadIdx.x & 31) == 0) red[j][threadIdx.x >> 5] = (double)w;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kRegThreads / 32; ++w) t += red[threadIdx.x][w];
    atomicAdd(a.sums + threadIdx.x, t);
  }
}

}  // namespace mobgs

using namespace mobgs;

extern "C" int mobgs_reg_loss_fwd(const MobgsRegLoss* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->n_depth >= 0 && a->n_alpha >= 0 && a->sums, "bad arguments");
  MOBGS_REQUIRE(a->n_depth == 0 || (a->depth && a->gt_depth), "NULL depth");
  MOBGS_REQUIRE(a->n_alpha == 0 || a->alpha, "NULL alpha");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(a->sums, 0, 3 * sizeof(double), s);
  const int64_t n = a->n_depth > a->n_alpha ? a->n_depth : a->n_alpha;
  if (n == 0) return MOBGS_OK;
  static int max_ctas = 0;
  if (!max_ctas) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    max_ctas = 8 * sms;
  }
  const int64_t want = (n + kRegThreads - 1) / kRegThreads;
  const int grid = want < (int64_t)max_ctas ? (int)want : max_ctas;
  reg_loss_kernel<<<grid, kRegThreads, 0, s>>>(*a);
  return check_launch("reg_loss_fwd");
}
