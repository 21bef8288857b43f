// Per-pixel Sandwich decoder arithmetic shared by decode.cu and the fused blend epilogue/prologue
// (reference helper_model.py:19-28).
#pragma once
#include "common.cuh"

namespace mobgs {

constexpr float kEdFloor = 1e-10f;
constexpr float kMeanEps = 1e-10f;

// The 90 decoder weights stay in shared memory (broadcast LDS.128 reads): the backward already
// needs 90 registers per thread for the weight-gradient accumulators.
struct DecW { const float* w1; const float* w2; };

__device__ __forceinline__ void load_w(DecW& w, const float* w1, const float* w2, float* smem) {
  for (int i = threadIdx.x; i < 90; i += blockDim.x) smem[i] = i < 72 ? w1[i] : w2[i - 72];
  __syncthreads();
  w.w1 = smem;
  w.w2 = smem + 72;
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


// VJP of sandwich_fwd for one pixel.  g_out[3] = d loss / d rgb.  Returns d loss / d img[0..8] in gv and
// d loss / d rays in g_rays; the weight-gradient terms are exposed through ghpre / gpre
// (d W1[j][i] += ghpre[j] * x[i],  d W2[c][j] += gpre[c] * relu(hpre[j])).
__device__ __forceinline__ void sandwich_bwd(const DecW& w, const float hpre[6], const float out[3],
                                             const float g_out[3], float gv[9], float g_rays[6],
                                             float gpre[3], float ghpre[6]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) { gpre[c] = g_out[c] * out[c] * (1.f - out[c]); gv[c] = gpre[c]; }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    float gh = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) gh += w.w2[6 * c + j] * gpre[c];
    ghpre[j] = hpre[j] > 0.f ? gh : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) s += w.w1[12 * j + i] * ghpre[j];
    if (i < 6) gv[3 + i] = s; else g_rays[i - 6] = s;
  }
}

}  // namespace mobgs
