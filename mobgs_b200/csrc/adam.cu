// f4: multi-tensor Adam step — every parameter tensor of both Gaussian models in ONE
// launch.  The reference builds `torch.optim.Adam(l, lr=0.0, eps=1e-15)` over 17 parameter groups per
// model (scene/gaussian_model.py:598-641) and steps two of them per iteration (train.py:796-800).
// Streaming kernel, HBM-bound: per element it reads param, grad, exp_avg, exp_avg_sq (16 B) and writes
// param, exp_avg, exp_avg_sq (12 B) = 28 algorithmic bytes.
//
// Arithmetic = torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False), per element:
//   m  = m + (g - m) (1 - beta1)                         (exp_avg.lerp_(grad, 1 - beta1))
//   v  = v beta2 + (1 - beta2) g g                       (exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2))
//   p -= step_size * m / (sqrt(v) / bc2_sqrt + eps)      (step_size = lr / (1 - beta1^t), bc2_sqrt = sqrt(1 - beta2^t))
// step_size and bc2_sqrt are per tensor (each parameter keeps its own step count, and densification
// re-creates state), computed on the host in double and passed in the argument struct.
#include "common.cuh"

namespace mobgs {

constexpr int kAdamThreads = 256;
constexpr int kAdamChunk = kAdamThreads * 16;   // elements per (CTA, iteration): 4 x float4 per thread and array

struct AdamConst { float w1, beta2, w2, eps; };   // w1 = 1 - beta1, w2 = 1 - beta2 (rounded from double, as torch passes them)

// Operation order and roundings of torch's foreach Adam: lerp = fma, mul_ rounds before addcmul's fma,
// sqrt / div / add are separate IEEE operations, addcdiv = fma.
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamConst& c,
                                            float step_size, float bc2_sqrt) {
  m = fmaf(c.w1, g - m, m);
  v = fmaf(c.w2 * g, g, __fmul_rn(v, c.beta2));
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), bc2_sqrt), c.eps);
  p = fmaf(-step_size, __fdiv_rn(m, denom), p);
}

__global__ void __launch_bounds__(kAdamThreads) adam_kernel(const __grid_constant__ MobgsAdam a, int total_chunks) {
  __shared__ int s_begin[MOBGS_ADAM_MAX_TENSORS + 1];
  for (int i = threadIdx.x; i <= a.n_tensors; i += kAdamThreads) s_begin[i] = a.chunk_begin[i];
  __syncthreads();
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    // which tensor owns this chunk (n_tensors <= 64: a short binary search over shared memory)
    int lo = 0, hi = a.n_tensors;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_begin[mid] <= chunk) lo = mid; else hi = mid;
    }
    const int t = lo;
    const int64_t n = a.numel[t];
    const int64_t base = (int64_t)(chunk - s_begin[t]) * kAdamChunk;
    float* __restrict__ p = a.param[t];
    const float* __restrict__ g = a.grad[t];
    float* __restrict__ m = a.exp_avg[t];
    float* __restrict__ v = a.exp_avg_sq[t];
    const float ss = a.step_size[t], bc = a.bc2_sqrt[t];
    const AdamConst c = {a.one_minus_beta1, a.beta2, a.one_minus_beta2, a.eps};
    const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0 && base + kAdamChunk <= n;
    if (vec) {
      float4 P[4], G[4], M[4], V[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t i = base / 4 + j * kAdamThreads + threadIdx.x;
        P[j] = reinterpret_cast<const float4*>(p)[i];
        G[j] = __ldcs(reinterpret_cast<const float4*>(g) + i);      // gradients are read once: streaming load
        M[j] = reinterpret_cast<const float4*>(m)[i];
        V[j] = reinterpret_cast<const float4*>(v)[i];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        adam_update(P[j].x, G[j].x, M[j].x, V[j].x, c, ss, bc);
        adam_update(P[j].y, G[j].y, M[j].y, V[j].y, c, ss, bc);
        adam_update(P[j].z, G[j].z, M[j].z, V[j].z, c, ss, bc);
        adam_update(P[j].w, G[j].w, M[j].w, V[j].w, c, ss, bc);
        const int64_t i = base / 4 + j * kAdamThreads + threadIdx.x;
        reinterpret_cast<float4*>(p)[i] = P[j];
        reinterpret_cast<float4*>(m)[i] = M[j];
        reinterpret_cast<float4*>(v)[i] = V[j];
      }
    } else {
      const int64_t end = min(n, base + kAdamChunk);
      for (int64_t i = base + threadIdx.x; i < end; i += kAdamThreads) {
        float pp = p[i], mm = m[i], vv = v[i];
        adam_update(pp, g[i], mm, vv, c, ss, bc);
        p[i] = pp; m[i] = mm; v[i] = vv;
      }
    }
  }
}

}  // namespace mobgs

using namespace mobgs;

extern "C" int mobgs_adam_chunk_elems(void) { return kAdamChunk; }

extern "C" int mobgs_adam_step(const MobgsAdam* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->n_tensors >= 0 && a->n_tensors <= MOBGS_ADAM_MAX_TENSORS, "n_tensors=%d out of range", a->n_tensors);
  if (a->n_tensors == 0) return MOBGS_OK;
  int chunks = 0;
  for (int t = 0; t < a->n_tensors; ++t) {
    MOBGS_REQUIRE(a->numel[t] >= 0, "negative numel");
    MOBGS_REQUIRE(a->numel[t] == 0 || (a->param[t] && a->grad[t] && a->exp_avg[t] && a->exp_avg_sq[t]), "NULL tensor %d", t);
    MOBGS_REQUIRE(a->chunk_begin[t] == chunks, "chunk_begin[%d] must be the running sum of ceil(numel / chunk)", t);
    chunks += (int)((a->numel[t] + kAdamChunk - 1) / kAdamChunk);
  }
  MOBGS_REQUIRE(a->chunk_begin[a->n_tensors] == chunks, "chunk_begin[n_tensors] must be the total chunk count");
  if (chunks == 0) return MOBGS_OK;
  static int max_ctas = 0;       // a multiple of the SM count: 8 resident CTAs of 256 threads per SM
  if (!max_ctas) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    max_ctas = 8 * sms;
  }
  adam_kernel<<<min(chunks, max_ctas), kAdamThreads, 0, (cudaStream_t)stream>>>(*a, chunks);
  return check_launch("adam_step");
}
