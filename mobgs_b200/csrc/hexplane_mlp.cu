// a11: fused HexPlane grid-sample + deformation MLP, forward.
//
// Reference: deform_network.forward -> Deformation.forward_dynamic2 (scene/deformation.py:158-199,
// 252-253, 285-290) over HexPlaneField (scene/hexplane.py:19-108, 165-187): 6 planes x L levels of
// bilinear grid_sample (align_corners=True, border), product over planes, concat over levels,
// Linear(32L -> 128), three heads ReLU-Linear(128,128)-ReLU-Linear(128,{7,3,4}), then the
// dx / se(3)-quirk / ds-clamp / quaternion-product post-processing.  The reference runs this as
// 18 grid_sample launches + ~12 cuBLAS/elementwise launches with N x 96 / N x 128 intermediates in
// HBM; here one CTA owns 128 points end to end and nothing but the inputs, the L2-resident planes
// and the 40-byte result per point touches global memory.
//
// Dense layers run on the 5th-gen tensor cores: tcgen05.mma kind::tf32, M=128 (one point per TMEM
// lane), operands in shared memory in the no-swizzle K-major core-matrix layout, fp32 accumulators
// in TMEM, read back with tcgen05.ld for bias/ReLU.  A single TF32 pass (10-bit mantissa) cannot
// hold the 1e-4 fp32 parity bar, so every product is issued as the error-compensated 3xTF32 split
//     a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo,   x_hi = tf32(x), x_lo = x - x_hi
// (three MMAs accumulate into the same TMEM tile; relative error ~2^-21).
// Weight tiles arrive by 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) from a
// pre-tiled hi/lo image the host prepares once per call.
#include "hexplane_tc.cuh"

namespace mobgs {

__global__ void __launch_bounds__(kHexThreads, 1) hexplane_mlp_fwd_kernel(const __grid_constant__ MobgsHexMlpFwd a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  HexSmem& sm = *reinterpret_cast<HexSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row0 = blockIdx.x * kHexRows;
  const int L = a.levels, K0 = L * kHexC;

  if (tid == 0) {
    mbar_init(&sm.bar_w, 1);
    mbar_init(&sm.bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t tmem_lane = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 lanes
  uint32_t ph_w = 0, ph_mma = 0;

  // ---- 1. HexPlane features: 8 lanes per point, 4 channels each -------------------------------
  {
    const int c4 = tid & 7;
    for (int pass = 0; pass < kHexRows / 16; ++pass) {
      const int row = pass * 16 + (tid >> 3);
      const int g = min(row0 + row, a.N - 1);
      float c[4];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float v = (a.pts[3 * g + i] - a.aabb[i]) * (2.0f / (a.aabb[3 + i] - a.aabb[i])) - 1.0f;
        c[i] = fminf(fmaxf(v, -1.f), 1.f);
      }
      c[3] = a.times[g];
      for (int l = 0; l < L; ++l) {
        float4 prod = make_float4(1.f, 1.f, 1.f, 1.f);
        int pi = 0;
#pragma unroll
        for (int ca = 0; ca < 4; ++ca)
#pragma unroll
          for (int cb = ca + 1; cb < 4; ++cb, ++pi) {
            const int id = l * 6 + pi;
            const PlaneSample s = plane_sample(c[ca], c[cb], a.plane_w[id], a.plane_h[id]);
            const float* p = a.planes[id] + 4 * c4;
            const float4 v00 = __ldg(reinterpret_cast<const float4*>(p + s.o00));
            const float4 v01 = __ldg(reinterpret_cast<const float4*>(p + s.o01));
            const float4 v10 = __ldg(reinterpret_cast<const float4*>(p + s.o10));
            const float4 v11 = __ldg(reinterpret_cast<const float4*>(p + s.o11));
            prod.x *= v00.x * s.w00 + v01.x * s.w01 + v10.x * s.w10 + v11.x * s.w11;
            prod.y *= v00.y * s.w00 + v01.y * s.w01 + v10.y * s.w10 + v11.y * s.w11;
            prod.z *= v00.z * s.w00 + v01.z * s.w01 + v10.z * s.w10 + v11.z * s.w11;
            prod.w *= v00.w * s.w00 + v01.w * s.w01 + v10.w * s.w10 + v11.w * s.w11;
          }
        store_a(sm, row, l * 8 + c4, prod);
      }
    }
  }

  // generic-proxy writes of A must be visible to the tensor core (async proxy)
  auto publish_a = [&]() {
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
  };
  // load one pre-tiled weight block (hi then lo, `bytes` each) and run the GEMM pass on it
  auto gemm_pass = [&](const float* w_tile, uint32_t bytes, uint32_t tmem_d, int K, int n) {
    if (tid == 0) {
      mbar_expect_tx(&sm.bar_w, 2 * bytes);
      bulk_g2s(sm.b_hi, w_tile, bytes, &sm.bar_w);
      bulk_g2s(sm.b_lo, reinterpret_cast<const char*>(w_tile) + bytes, bytes, &sm.bar_w);
      mbar_wait(&sm.bar_w, ph_w);
      tc_fence_after();
      issue_gemm(sm, tmem_d, K, n);
      umma_commit(&sm.bar_mma);
    }
    ph_w ^= 1;
    mbar_wait(&sm.bar_mma, ph_mma);   // all threads: accumulators complete, B tile free again
    ph_mma ^= 1;
    tc_fence_after();
  };

  // ---- 2. hidden = feat @ W0^T (bias added when read) -----------------------------------------
  publish_a();
  {
    const uint32_t bytes = (uint32_t)(K0 / 4) * 64 * 16;
    for (int half = 0; half < 2; ++half)
      gemm_pass(a.w0 + (size_t)half * 2 * (bytes / 4), bytes, tmem + half * 64, K0, 64);
  }

  float out[3][16];
  // ---- 3. three heads --------------------------------------------------------------------------
  for (int h = 0; h < 3; ++h) {
    // A <- relu(hidden + b0)   (hidden stays in TMEM columns [0,128) for the next head)
#pragma unroll 1
    for (int cb = 0; cb < kHexW / 32; ++cb) {
      float v[32];
      tmem_ld32(tmem_lane + cb * 32, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = cb * 32 + 4 * j;
        store_a(sm, tid, k >> 2, make_float4(fmaxf(v[4 * j] + a.b0[k], 0.f), fmaxf(v[4 * j + 1] + a.b0[k + 1], 0.f),
                                             fmaxf(v[4 * j + 2] + a.b0[k + 2], 0.f), fmaxf(v[4 * j + 3] + a.b0[k + 3], 0.f)));
      }
    }
    publish_a();
    {
      const uint32_t bytes = (kHexW / 4) * 64 * 16;
      const float* wa = a.wa + (size_t)h * 4 * (bytes / 4);
      for (int half = 0; half < 2; ++half)
        gemm_pass(wa + (size_t)half * 2 * (bytes / 4), bytes, tmem + kHexW + half * 64, kHexW, 64);
    }
    // A <- relu(D + ba[h])
#pragma unroll 1
    for (int cb = 0; cb < kHexW / 32; ++cb) {
      float v[32];
      tmem_ld32(tmem_lane + kHexW + cb * 32, v);
      const float* b = a.ba + h * kHexW;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = cb * 32 + 4 * j;
        store_a(sm, tid, k >> 2, make_float4(fmaxf(v[4 * j] + b[k], 0.f), fmaxf(v[4 * j + 1] + b[k + 1], 0.f),
                                             fmaxf(v[4 * j + 2] + b[k + 2], 0.f), fmaxf(v[4 * j + 3] + b[k + 3], 0.f)));
      }
    }
    publish_a();
    {
      const uint32_t bytes = (kHexW / 4) * 16 * 16;      // 16 (zero-padded) output rows
      gemm_pass(a.wb + (size_t)h * 2 * (bytes / 4), bytes, tmem + kHexW, kHexW, 16);
    }
    tmem_ld16(tmem_lane + kHexW, out[h]);
#pragma unroll
    for (int j = 0; j < 16; ++j) out[h][j] += a.bb[h * 16 + j];
    tc_fence_before();
    __syncthreads();     // every thread has read D before the next head overwrites it
    tc_fence_after();
  }

  // ---- 4. forward_dynamic2 post-processing (scene/deformation.py:177-197) ---------------------
  const int g = row0 + tid;
  if (g < a.N) {
    const float* dx = out[0];
    // quat2mat quirk (:417-438): [1, q0..q3] normalised by its 5-vector norm, first four used
    const float n5 = 1.0f / sqrtf(1.f + dx[3] * dx[3] + dx[4] * dx[4] + dx[5] * dx[5] + dx[6] * dx[6]);
    const float qw = n5, qx = dx[3] * n5, qy = dx[4] * n5, qz = dx[5] * n5;
    const float p0 = a.pts[3 * g] + dx[0], p1 = a.pts[3 * g + 1] + dx[1], p2 = a.pts[3 * g + 2] + dx[2];
    const float w2 = qw * qw, x2 = qx * qx, y2 = qy * qy, z2 = qz * qz;
    const float wx = qw * qx, wy = qw * qy, wz = qw * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz;
    a.out_pts[3 * g + 0] = (w2 + x2 - y2 - z2) * p0 + (2 * xy - 2 * wz) * p1 + (2 * wy + 2 * xz) * p2;
    a.out_pts[3 * g + 1] = (2 * wz + 2 * xy) * p0 + (w2 - x2 + y2 - z2) * p1 + (2 * yz - 2 * wx) * p2;
    a.out_pts[3 * g + 2] = (2 * xz - 2 * wy) * p0 + (2 * wx + 2 * yz) * p1 + (w2 - x2 - y2 + z2) * p2;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      a.out_scales[3 * g + i] = a.scales[3 * g + i] + fminf(fmaxf(out[1][i], -kLog100), kLog100);
    float q1[4], q2[4] = {dx[3], dx[4], dx[5], dx[6]};
#pragma unroll
    for (int i = 0; i < 4; ++i) q1[i] = a.rots[4 * g + i] + out[2][i];
    float r[4];
    r[0] = q1[0] * q2[0] - q1[1] * q2[1] - q1[2] * q2[2] - q1[3] * q2[3];
    r[1] = q1[0] * q2[1] + q1[1] * q2[0] + q1[2] * q2[3] - q1[3] * q2[2];
    r[2] = q1[0] * q2[2] - q1[1] * q2[3] + q1[2] * q2[0] + q1[3] * q2[1];
    r[3] = q1[0] * q2[3] + q1[1] * q2[2] - q1[2] * q2[1] + q1[3] * q2[0];
    const float rn = 1.0f / sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) a.out_rots[4 * g + i] = r[i] * rn;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace mobgs

using namespace mobgs;

extern "C" int mobgs_hexplane_mlp_fwd(const MobgsHexMlpFwd* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->N >= 0, "N < 0");
  MOBGS_REQUIRE(a->levels >= 1 && a->levels <= kMaxLevels, "levels=%d out of range [1,%d]", a->levels, kMaxLevels);
  MOBGS_REQUIRE(a->net_width == kHexW && a->plane_features == kHexC,
                "this build covers net_width=%d, %d features per plane (the stereo configs)", kHexW, kHexC);
  if (a->N == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->pts && a->scales && a->rots && a->times && a->w0 && a->b0 && a->wa && a->ba && a->wb && a->bb,
                "NULL input");
  MOBGS_REQUIRE(a->out_pts && a->out_scales && a->out_rots, "NULL output");
  for (int i = 0; i < a->levels * 6; ++i)
    MOBGS_REQUIRE(a->planes[i] && a->plane_w[i] >= 1 && a->plane_h[i] >= 1, "bad plane %d", i);
  const size_t smem = sizeof(HexSmem) + 128;
  cudaFuncSetAttribute(hexplane_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = (a->N + kHexRows - 1) / kHexRows;
  hexplane_mlp_fwd_kernel<<<grid, kHexThreads, smem, (cudaStream_t)stream>>>(*a);
  return check_launch("hexplane_mlp_fwd");
}
