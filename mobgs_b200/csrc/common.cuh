// Shared helpers for the mobgs_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mobgs_b200.h"
#include "gs_math.cuh"

namespace mobgs {

constexpr int kMaxK = MOBGS_MAX_K;   // sub-frames per launch (reference num_warp = 9)
constexpr int kTile = MOBGS_TILE;
constexpr int kTilePix = kTile * kTile;

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define MOBGS_REQUIRE(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::mobgs::set_error(__VA_ARGS__);      \
      return MOBGS_EINVAL;                  \
    }                                       \
  } while (0)

__device__ __forceinline__ Cam load_cam(const float* __restrict__ viewmats,
                                        const float* __restrict__ Ks, int k) {
  Cam c;
  const float* v = viewmats + 16 * k;
  c.r[0] = v[0]; c.r[1] = v[1]; c.r[2] = v[2]; c.t[0] = v[3];
  c.r[3] = v[4]; c.r[4] = v[5]; c.r[5] = v[6]; c.t[1] = v[7];
  c.r[6] = v[8]; c.r[7] = v[9]; c.r[8] = v[10]; c.t[2] = v[11];
  const float* kk = Ks + 9 * k;
  c.fx = kk[0]; c.cx = kk[2]; c.fy = kk[4]; c.cy = kk[5];
  return c;
}

__device__ __forceinline__ ProjCfg make_cfg(const MobgsCameras& c) {
  ProjCfg cfg;
  cfg.width = c.width; cfg.height = c.height;
  cfg.eps2d = c.eps2d; cfg.near_plane = c.near_plane; cfg.far_plane = c.far_plane;
  cfg.radius_clip = c.radius_clip;
  return cfg;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum of 16 per-lane values over the warp with a transposing butterfly: every stage halves the values a lane holds while
// doubling the lanes summed — 16 shuffles instead of 80.  Afterwards v[0] of every EVEN lane holds the complete warp sum
// of value index `return` (lanes 2 i and 2 i + 1 hold the same one); each of the 16 values has exactly one even owner.
__device__ __forceinline__ int warp_transpose_sum16(float (&v)[16], int lane) {
  int vidx = 0;
#pragma unroll
  for (int o = 16, n = 8; o >= 2; o >>= 1, n >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < n) {
        const float send = up ? v[i] : v[i + n];
        const float keep = up ? v[i + n] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    vidx += up ? n : 0;
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return vidx;
}

// 16-byte vector reduction to global memory (sm_90+): one L2 atomic op for four floats.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

}  // namespace mobgs
