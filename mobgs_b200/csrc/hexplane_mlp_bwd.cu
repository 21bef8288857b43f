// a11 backward on tcgen05 (reference: autograd through scene/deformation.py:158-199 and
// scene/hexplane.py:19-108 — 18 grid_sample_backward + ~16 cuBLAS SGEMM launches with N x 96 / N x 128
// intermediates in HBM).
//
// hexplane_mlp_bwd_kernel (one CTA = 128 points = 128 TMEM lanes, the forward's structure):
//   1. recompute the forward on the tensor cores (features -> h0 -> per head z1 -> o), keeping h0 in TMEM
//      columns [0,128) and the relu masks of z1 as 3 x 128 bits per thread;
//   2. per-point VJP of the post-processing (quat2mat quirk, scale clamp, quaternion product) -> g_o;
//   3. DATA gradients as 3xTF32 GEMMs against pre-tiled transposed weights:
//        g_a2 = g_o Wb            (K = 16)     -> g_z1 = g_a2 * (z1 > 0)
//        S   += g_z1 Wa           (K = 128)    accumulated over the three heads IN TMEM (the relu mask of h0
//                                               is head independent, so it is applied once, after the sum)
//        g_h0 = S * (h0 > 0);  g_feat = g_h0 W0 (K = 128)
//   4. every operand the weight gradients need leaves feature-major ([row][point]) so that the contraction
//      over the points is K-major for the second kernel.
// hexplane_wgrad_kernel: C[m][n] += sum_k A[m][k] B[n][k], k = points; each CTA owns a K range, stages
// 32-point chunks of both operands (hi/lo split) and keeps the 128 x n accumulator in TMEM until its range is
// done; one vector-atomic flush per CTA.  Row sums of A or B (bias gradients) ride the staging loads.
#include "hexplane_tc.cuh"

namespace mobgs {

constexpr int kBwdTmemCols = 512;      // [0,128) h0 | [128,256) layer outputs | [256,384) sum_h g_z1 Wa

__device__ __forceinline__ void issue_gemm_acc(HexSmem& sm, uint32_t tmem_d, int K, int n, uint32_t acc) {
  const uint32_t idesc = make_idesc(n);
  const uint32_t a_lbo = kHexRows * 16, b_lbo = (uint32_t)n * 16;
  const uint32_t a_hi = smem_u32(sm.a_hi), a_lo = smem_u32(sm.a_lo);
  const uint32_t b_hi = smem_u32(sm.b_hi), b_lo = smem_u32(sm.b_lo);
  for (int k8 = 0; k8 < K / 8; ++k8) {
    const uint32_t ao = (uint32_t)k8 * 2 * a_lbo, bo = (uint32_t)k8 * 2 * b_lbo;
    umma_tf32(tmem_d, make_desc(a_hi + ao, a_lbo, 128), make_desc(b_hi + bo, b_lbo, 128), idesc, acc);
    acc = 1;
    umma_tf32(tmem_d, make_desc(a_lo + ao, a_lbo, 128), make_desc(b_hi + bo, b_lbo, 128), idesc, 1);
    umma_tf32(tmem_d, make_desc(a_hi + ao, a_lbo, 128), make_desc(b_lo + bo, b_lbo, 128), idesc, 1);
  }
}

__device__ __forceinline__ float4 load_a(const HexSmem& sm, int row, int k4) {
  const float4 hi = reinterpret_cast<const float4*>(sm.a_hi)[k4 * kHexRows + row];
  const float4 lo = reinterpret_cast<const float4*>(sm.a_lo)[k4 * kHexRows + row];
  return make_float4(hi.x + lo.x, hi.y + lo.y, hi.z + lo.z, hi.w + lo.w);
}

__global__ void __launch_bounds__(kHexThreads, 1) hexplane_mlp_bwd_kernel(const __grid_constant__ MobgsHexMlpBwd a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  HexSmem& sm = *reinterpret_cast<HexSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row0 = blockIdx.x * kHexRows;
  const int L = a.levels, K0 = L * kHexC;
  const int n = row0 + tid;                       // this thread's point
  const bool live = n < a.N;
  const size_t ld = (size_t)a.ld;

  if (tid == 0) {
    mbar_init(&sm.bar_w, 1);
    mbar_init(&sm.bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, kBwdTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t tmem_lane = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t ph_w = 0, ph_mma = 0;

  auto publish_a = [&]() {
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
  };
  auto gemm_pass = [&](const float* w_tile, uint32_t bytes, uint32_t tmem_d, int K, int ncol, uint32_t acc) {
    if (tid == 0) {
      mbar_expect_tx(&sm.bar_w, 2 * bytes);
      bulk_g2s(sm.b_hi, w_tile, bytes, &sm.bar_w);
      bulk_g2s(sm.b_lo, reinterpret_cast<const char*>(w_tile) + bytes, bytes, &sm.bar_w);
      mbar_wait(&sm.bar_w, ph_w);
      tc_fence_after();
      issue_gemm_acc(sm, tmem_d, K, ncol, acc);
      umma_commit(&sm.bar_mma);
    }
    ph_w ^= 1;
    mbar_wait(&sm.bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
  };
  auto sync_tmem_readers = [&]() {     // every thread has read D before a later GEMM overwrites it
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  };

  // ---- 1a. HexPlane features (forward kernel's gather) ------------------------------------------------
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
  __syncthreads();
  // features, feature-major (thread = point: coalesced rows), zero beyond N
  for (int k4 = 0; k4 < K0 / 4; ++k4) {
    const float4 f = load_a(sm, tid, k4);
    float* o = a.featT + (size_t)(4 * k4) * ld + n;
    o[0] = live ? f.x : 0.f; o[ld] = live ? f.y : 0.f; o[2 * ld] = live ? f.z : 0.f; o[3 * ld] = live ? f.w : 0.f;
  }

  // ---- 1b. h0 = feat W0^T -> TMEM [0,128) ----------------------------------------------------------------
  publish_a();
  {
    const uint32_t bytes = (uint32_t)(K0 / 4) * 64 * 16;
    for (int half = 0; half < 2; ++half)
      gemm_pass(a.w0 + (size_t)half * 2 * (bytes / 4), bytes, tmem + half * 64, K0, 64, 0);
  }

  // ---- 1c. the three heads; keep the outputs and the relu masks of z1 -----------------------------------
  float o0[7], o1[3], o2[4];
  uint32_t m2[3][4];
#pragma unroll 1
  for (int h = 0; h < 3; ++h) {
#pragma unroll 1
    for (int cb = 0; cb < kHexW / 32; ++cb) {
      float v[32];
      tmem_ld32(tmem_lane + cb * 32, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = cb * 32 + 4 * j;
        const float4 r = make_float4(fmaxf(v[4 * j] + a.b0[k], 0.f), fmaxf(v[4 * j + 1] + a.b0[k + 1], 0.f),
                                     fmaxf(v[4 * j + 2] + a.b0[k + 2], 0.f), fmaxf(v[4 * j + 3] + a.b0[k + 3], 0.f));
        store_a(sm, tid, k >> 2, r);
        if (h == 0) {
          float* o = a.a1T + (size_t)k * ld + n;
          o[0] = live ? r.x : 0.f; o[ld] = live ? r.y : 0.f; o[2 * ld] = live ? r.z : 0.f; o[3 * ld] = live ? r.w : 0.f;
        }
      }
    }
    publish_a();
    {
      const uint32_t bytes = (kHexW / 4) * 64 * 16;
      const float* wa = a.wa + (size_t)h * 4 * (bytes / 4);
      for (int half = 0; half < 2; ++half)
        gemm_pass(wa + (size_t)half * 2 * (bytes / 4), bytes, tmem + kHexW + half * 64, kHexW, 64, 0);
    }
#pragma unroll 1
    for (int cb = 0; cb < kHexW / 32; ++cb) {
      float v[32];
      tmem_ld32(tmem_lane + kHexW + cb * 32, v);
      const float* b = a.ba + h * kHexW;
      uint32_t bits = 0u;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = cb * 32 + 4 * j;
        const float z0 = v[4 * j] + b[k], z1 = v[4 * j + 1] + b[k + 1], z2 = v[4 * j + 2] + b[k + 2], z3 = v[4 * j + 3] + b[k + 3];
        bits |= (z0 > 0.f ? 1u : 0u) << (4 * j) | (z1 > 0.f ? 1u : 0u) << (4 * j + 1) | (z2 > 0.f ? 1u : 0u) << (4 * j + 2) |
                (z3 > 0.f ? 1u : 0u) << (4 * j + 3);
        const float4 r = make_float4(fmaxf(z0, 0.f), fmaxf(z1, 0.f), fmaxf(z2, 0.f), fmaxf(z3, 0.f));
        store_a(sm, tid, k >> 2, r);
        float* o = a.a2T + ((size_t)h * kHexW + k) * ld + n;
        o[0] = live ? r.x : 0.f; o[ld] = live ? r.y : 0.f; o[2 * ld] = live ? r.z : 0.f; o[3 * ld] = live ? r.w : 0.f;
      }
      m2[h][cb] = bits;
    }
    publish_a();
    {
      const uint32_t bytes = (kHexW / 4) * 16 * 16;
      gemm_pass(a.wb + (size_t)h * 2 * (bytes / 4), bytes, tmem + kHexW, kHexW, 16, 0);
    }
    {
      float v[16];
      tmem_ld16(tmem_lane + kHexW, v);
      if (h == 0) {
#pragma unroll
        for (int j = 0; j < 7; ++j) o0[j] = v[j] + a.bb[j];
      } else if (h == 1) {
#pragma unroll
        for (int j = 0; j < 3; ++j) o1[j] = v[j] + a.bb[16 + j];
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) o2[j] = v[j] + a.bb[32 + j];
      }
    }
    sync_tmem_readers();
  }

  // ---- 2. VJP of forward_dynamic2's post-processing (scene/deformation.py:177-197, :417-438) ----------------
  float go0[7] = {0, 0, 0, 0, 0, 0, 0}, go1[3] = {0, 0, 0}, go2[4] = {0, 0, 0, 0};
  {
    const int g = min(n, a.N - 1);
    float gp[3] = {0, 0, 0}, gs[3] = {0, 0, 0}, gr[4] = {0, 0, 0, 0};
    if (live) {
      if (a.g_out_pts) { gp[0] = a.g_out_pts[3 * g]; gp[1] = a.g_out_pts[3 * g + 1]; gp[2] = a.g_out_pts[3 * g + 2]; }
      if (a.g_out_scales) { gs[0] = a.g_out_scales[3 * g]; gs[1] = a.g_out_scales[3 * g + 1]; gs[2] = a.g_out_scales[3 * g + 2]; }
      if (a.g_out_rots) {
        const float4 t = *reinterpret_cast<const float4*>(a.g_out_rots + 4 * g);
        gr[0] = t.x; gr[1] = t.y; gr[2] = t.z; gr[3] = t.w;
      }
    }
    // out_scales = scales + clamp(ds, -log 100, log 100)
#pragma unroll
    for (int i = 0; i < 3; ++i) go1[i] = (o1[i] >= -kLog100 && o1[i] <= kLog100) ? gs[i] : 0.f;
    // out_rots = normalize((rots + dr) (x) dx[3:7])
    float q1[4], q2[4] = {o0[3], o0[4], o0[5], o0[6]};
    {
      const float4 t = *reinterpret_cast<const float4*>(a.rots + 4 * g);
      q1[0] = t.x + o2[0]; q1[1] = t.y + o2[1]; q1[2] = t.z + o2[2]; q1[3] = t.w + o2[3];
    }
    float r[4];
    r[0] = q1[0] * q2[0] - q1[1] * q2[1] - q1[2] * q2[2] - q1[3] * q2[3];
    r[1] = q1[0] * q2[1] + q1[1] * q2[0] + q1[2] * q2[3] - q1[3] * q2[2];
    r[2] = q1[0] * q2[2] - q1[1] * q2[3] + q1[2] * q2[0] + q1[3] * q2[1];
    r[3] = q1[0] * q2[3] + q1[1] * q2[2] - q1[2] * q2[1] + q1[3] * q2[0];
    const float rn = 1.0f / sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
    const float dot = (gr[0] * r[0] + gr[1] * r[1] + gr[2] * r[2] + gr[3] * r[3]) * rn * rn;
    float g4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) g4[i] = (gr[i] - dot * r[i]) * rn;
    float ga[4], gb[4];
    ga[0] = g4[0] * q2[0] + g4[1] * q2[1] + g4[2] * q2[2] + g4[3] * q2[3];
    ga[1] = -g4[0] * q2[1] + g4[1] * q2[0] - g4[2] * q2[3] + g4[3] * q2[2];
    ga[2] = -g4[0] * q2[2] + g4[1] * q2[3] + g4[2] * q2[0] - g4[3] * q2[1];
    ga[3] = -g4[0] * q2[3] - g4[1] * q2[2] + g4[2] * q2[1] + g4[3] * q2[0];
    gb[0] = g4[0] * q1[0] + g4[1] * q1[1] + g4[2] * q1[2] + g4[3] * q1[3];
    gb[1] = -g4[0] * q1[1] + g4[1] * q1[0] + g4[2] * q1[3] - g4[3] * q1[2];
    gb[2] = -g4[0] * q1[2] - g4[1] * q1[3] + g4[2] * q1[0] + g4[3] * q1[1];
    gb[3] = -g4[0] * q1[3] + g4[1] * q1[2] - g4[2] * q1[1] + g4[3] * q1[0];
#pragma unroll
    for (int i = 0; i < 4; ++i) { go2[i] = ga[i]; go0[3 + i] = gb[i]; }
    // out_pts = R(w,x,y,z) (pts + dx[0:3]),  (w,x,y,z) = (1, dx3, dx4, dx5) / |(1, dx3, dx4, dx5, dx6)|
    const float n5 = 1.0f / sqrtf(1.f + o0[3] * o0[3] + o0[4] * o0[4] + o0[5] * o0[5] + o0[6] * o0[6]);
    const float w = n5, x = o0[3] * n5, y = o0[4] * n5, z = o0[5] * n5;
    const float p0 = a.pts[3 * g] + o0[0], p1 = a.pts[3 * g + 1] + o0[1], p2 = a.pts[3 * g + 2] + o0[2];
    const float R00 = w * w + x * x - y * y - z * z, R01 = 2 * x * y - 2 * w * z, R02 = 2 * w * y + 2 * x * z;
    const float R10 = 2 * w * z + 2 * x * y, R11 = w * w - x * x + y * y - z * z, R12 = 2 * y * z - 2 * w * x;
    const float R20 = 2 * x * z - 2 * w * y, R21 = 2 * w * x + 2 * y * z, R22 = w * w - x * x - y * y + z * z;
    const float gpx = R00 * gp[0] + R10 * gp[1] + R20 * gp[2];       // R^T gp
    const float gpy = R01 * gp[0] + R11 * gp[1] + R21 * gp[2];
    const float gpz = R02 * gp[0] + R12 * gp[1] + R22 * gp[2];
    go0[0] = gpx; go0[1] = gpy; go0[2] = gpz;
    const float G00 = gp[0] * p0, G01 = gp[0] * p1, G02 = gp[0] * p2, G10 = gp[1] * p0, G11 = gp[1] * p1, G12 = gp[1] * p2,
                G20 = gp[2] * p0, G21 = gp[2] * p1, G22 = gp[2] * p2;
    const float gw = 2.f * (w * (G00 + G11 + G22) + (-z * G01 + y * G02 + z * G10 - x * G12 - y * G20 + x * G21));
    const float gx = 2.f * (x * (G00 - G11 - G22) + (y * G01 + z * G02 + y * G10 - w * G12 + z * G20 + w * G21));
    const float gy = 2.f * (y * (-G00 + G11 - G22) + (x * G01 + w * G02 + x * G10 + z * G12 - w * G20 + z * G21));
    const float gz = 2.f * (z * (-G00 - G11 + G22) + (-w * G01 + x * G02 + w * G10 + y * G12 + x * G20 + y * G21));
    const float gn5 = gw + gx * o0[3] + gy * o0[4] + gz * o0[5];
    const float n53 = n5 * n5 * n5;
    go0[3] += gx * n5 - gn5 * o0[3] * n53;
    go0[4] += gy * n5 - gn5 * o0[4] * n53;
    go0[5] += gz * n5 - gn5 * o0[5] * n53;
    go0[6] += -gn5 * o0[6] * n53;
    if (live) {
      a.g_pts[3 * g] = gpx; a.g_pts[3 * g + 1] = gpy; a.g_pts[3 * g + 2] = gpz;
      a.g_scales[3 * g] = gs[0]; a.g_scales[3 * g + 1] = gs[1]; a.g_scales[3 * g + 2] = gs[2];
      *reinterpret_cast<float4*>(a.g_rots + 4 * g) = make_float4(ga[0], ga[1], ga[2], ga[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 7; ++i) go0[i] = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) go1[i] = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) go2[i] = 0.f;
    }
  }

  // ---- 3. data gradients ----------------------------------------------------------------------------------
#pragma unroll 1
  for (int h = 0; h < 3; ++h) {
    // A <- g_o[h], zero-padded to K = 16
    float g16[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) g16[j] = 0.f;
    if (h == 0) {
#pragma unroll
      for (int j = 0; j < 7; ++j) g16[j] = go0[j];
    } else if (h == 1) {
#pragma unroll
      for (int j = 0; j < 3; ++j) g16[j] = go1[j];
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) g16[j] = go2[j];
    }
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      store_a(sm, tid, k4, make_float4(g16[4 * k4], g16[4 * k4 + 1], g16[4 * k4 + 2], g16[4 * k4 + 3]));
#pragma unroll
      for (int j = 0; j < 4; ++j) a.goT[((size_t)h * 16 + 4 * k4 + j) * ld + n] = g16[4 * k4 + j];
    }
    publish_a();
    {
      const uint32_t bytes = (16 / 4) * 64 * 16;                    // Wb^T: 64 rows (inputs) x K = 16
      const float* wt = a.wb_t + (size_t)h * 4 * (bytes / 4);
      for (int half = 0; half < 2; ++half)
        gemm_pass(wt + (size_t)half * 2 * (bytes / 4), bytes, tmem + kHexW + half * 64, 16, 64, 0);
    }
    // g_z1 = g_a2 * (z1 > 0) -> A, and feature-major for the weight gradient of Wa
#pragma unroll 1
    for (int cb = 0; cb < kHexW / 32; ++cb) {
      float v[32];
      tmem_ld32(tmem_lane + kHexW + cb * 32, v);
      const uint32_t bits = m2[h][cb];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = cb * 32 + 4 * j;
        const float4 r = make_float4((bits >> (4 * j)) & 1u ? v[4 * j] : 0.f, (bits >> (4 * j + 1)) & 1u ? v[4 * j + 1] : 0.f,
                                     (bits >> (4 * j + 2)) & 1u ? v[4 * j + 2] : 0.f, (bits >> (4 * j + 3)) & 1u ? v[4 * j + 3] : 0.f);
        store_a(sm, tid, k >> 2, r);
        float* o = a.gz1T + ((size_t)h * kHexW + k) * ld + n;
        o[0] = r.x; o[ld] = r.y; o[2 * ld] = r.z; o[3 * ld] = r.w;
      }
    }
    publish_a();
    {
      const uint32_t bytes = (kHexW / 4) * 64 * 16;                 // Wa^T: 64 rows (inputs) x K = 128
      const float* wt = a.wa_t + (size_t)h * 4 * (bytes / 4);
      for (int half = 0; half < 2; ++half)
        gemm_pass(wt + (size_t)half * 2 * (bytes / 4), bytes, tmem + 2 * kHexW + half * 64, kHexW, 64, h > 0 ? 1u : 0u);
    }
    sync_tmem_readers();
  }
  // g_h0 = (sum_h g_a1) * (h0 + b0 > 0) -> A, feature-major for the weight gradient of W0
#pragma unroll 1
  for (int cb = 0; cb < kHexW / 32; ++cb) {
    float v[32], hv[32];
    tmem_ld32(tmem_lane + 2 * kHexW + cb * 32, v);
    tmem_ld32(tmem_lane + cb * 32, hv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = cb * 32 + 4 * j;
      const float4 r = make_float4(hv[4 * j] + a.b0[k] > 0.f ? v[4 * j] : 0.f, hv[4 * j + 1] + a.b0[k + 1] > 0.f ? v[4 * j + 1] : 0.f,
                                   hv[4 * j + 2] + a.b0[k + 2] > 0.f ? v[4 * j + 2] : 0.f,
                                   hv[4 * j + 3] + a.b0[k + 3] > 0.f ? v[4 * j + 3] : 0.f);
      store_a(sm, tid, k >> 2, r);
      float* o = a.gh0T + (size_t)k * ld + n;
      o[0] = r.x; o[ld] = r.y; o[2 * ld] = r.z; o[3 * ld] = r.w;
    }
  }
  publish_a();
  {
    // g_feat = g_h0 W0: W0^T tiles of min(64, K0) and K0 - 64 rows, K = 128
    const int r0 = K0 < 64 ? K0 : 64, r1 = K0 - r0;
    const uint32_t bytes0 = (kHexW / 4) * (uint32_t)r0 * 16;
    gemm_pass(a.w0_t, bytes0, tmem + kHexW, kHexW, r0, 0);
    if (r1 > 0) {
      const uint32_t bytes1 = (kHexW / 4) * (uint32_t)r1 * 16;
      gemm_pass(a.w0_t + 2 * (bytes0 / 4), bytes1, tmem + kHexW + 64, kHexW, r1, 0);
    }
  }
#pragma unroll 1
  for (int cb = 0; cb < K0 / 32; ++cb) {
    float v[32];
    tmem_ld32(tmem_lane + kHexW + cb * 32, v);
    if (live) {
      float4* o = reinterpret_cast<float4*>(a.g_feat + (size_t)n * K0 + cb * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kBwdTmemCols);
}

// ---------------------------------------------------------------------------------------------------------
// weight gradients: split-K (over the points) 3xTF32 GEMMs with TMEM-resident accumulators
// ---------------------------------------------------------------------------------------------------------
constexpr int kWgChunk = 32;                       // points per staged chunk
struct WgSmem {
  float a_hi[kWgChunk / 4 * 128 * 4];              // 16 KB each
  float a_lo[kWgChunk / 4 * 128 * 4];
  float b_hi[kWgChunk / 4 * 128 * 4];
  float b_lo[kWgChunk / 4 * 128 * 4];
  uint64_t bar_mma;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 2) hexplane_wgrad_kernel(const __grid_constant__ MobgsHexWgrad a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WgSmem& sm = *reinterpret_cast<WgSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int p = blockIdx.y;
  const int chunks = a.ld / kWgChunk;
  const int per = (chunks + gridDim.x - 1) / gridDim.x;
  const int c0 = blockIdx.x * per, c1 = min(chunks, c0 + per);
  if (c0 >= c1) return;
  const int ncol = a.n_cols[p];
  const float* __restrict__ A = a.A[p];
  const float* __restrict__ B = a.B[p];
  const size_t ld = (size_t)a.ld;

  if (tid == 0) {
    mbar_init(&sm.bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t tmem_lane = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t ph = 0;
  const uint32_t idesc = make_idesc(ncol);
  const uint32_t a_lbo = 128 * 16, b_lbo = (uint32_t)ncol * 16;
  float sum_a = 0.f, sum_b = 0.f;

  for (int c = c0; c < c1; ++c) {
    const float4* ar = reinterpret_cast<const float4*>(A + (size_t)tid * ld + (size_t)c * kWgChunk);
#pragma unroll
    for (int k4 = 0; k4 < kWgChunk / 4; ++k4) {
      const float4 v = __ldg(ar + k4);
      sum_a += (v.x + v.y) + (v.z + v.w);
      const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
      reinterpret_cast<float4*>(sm.a_hi)[k4 * 128 + tid] = hi;
      reinterpret_cast<float4*>(sm.a_lo)[k4 * 128 + tid] = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
    }
    if (tid < ncol) {
      const float4* br = reinterpret_cast<const float4*>(B + (size_t)tid * ld + (size_t)c * kWgChunk);
#pragma unroll
      for (int k4 = 0; k4 < kWgChunk / 4; ++k4) {
        const float4 v = __ldg(br + k4);
        sum_b += (v.x + v.y) + (v.z + v.w);
        const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        reinterpret_cast<float4*>(sm.b_hi)[k4 * ncol + tid] = hi;
        reinterpret_cast<float4*>(sm.b_lo)[k4 * ncol + tid] = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t ahi = smem_u32(sm.a_hi), alo = smem_u32(sm.a_lo), bhi = smem_u32(sm.b_hi), blo = smem_u32(sm.b_lo);
#pragma unroll
      for (int k8 = 0; k8 < kWgChunk / 8; ++k8) {
        const uint32_t ao = (uint32_t)k8 * 2 * a_lbo, bo = (uint32_t)k8 * 2 * b_lbo;
        umma_tf32(tmem, make_desc(ahi + ao, a_lbo, 128), make_desc(bhi + bo, b_lbo, 128), idesc, (c > c0 || k8 > 0) ? 1u : 0u);
        umma_tf32(tmem, make_desc(alo + ao, a_lbo, 128), make_desc(bhi + bo, b_lbo, 128), idesc, 1);
        umma_tf32(tmem, make_desc(ahi + ao, a_lbo, 128), make_desc(blo + bo, b_lbo, 128), idesc, 1);
      }
      umma_commit(&sm.bar_mma);
    }
    mbar_wait(&sm.bar_mma, ph);        // operands consumed: the staging buffers may be overwritten
    ph ^= 1;
    tc_fence_after();
  }

  // flush: lane = output row m (0..127)
  float* C = a.C[p];
  const int ldc = a.ldc[p];
  for (int j0 = 0; j0 < ncol; j0 += 16) {
    float v[16];
    tmem_ld16(tmem_lane + j0, v);
    if (a.transpose_out[p]) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (v[j] != 0.f) atomicAdd(C + (size_t)(j0 + j) * ldc + tid, v[j]);
    } else {
      float* dst = C + (size_t)tid * ldc + j0;
#pragma unroll
      for (int j = 0; j < 16; j += 4) red_add_v4(dst + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  if (a.bias[p]) {
    if (a.bias_from[p] == 1 && sum_a != 0.f) atomicAdd(a.bias[p] + tid, sum_a);
    if (a.bias_from[p] == 2 && tid < ncol && sum_b != 0.f) atomicAdd(a.bias[p] + tid, sum_b);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace mobgs

using namespace mobgs;

extern "C" int mobgs_hexplane_mlp_bwd(const MobgsHexMlpBwd* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->N >= 0, "N < 0");
  MOBGS_REQUIRE(a->levels >= 1 && a->levels <= kMaxLevels, "levels=%d out of range [1,%d]", a->levels, kMaxLevels);
  MOBGS_REQUIRE(a->net_width == kHexW && a->plane_features == kHexC,
                "this build covers net_width=%d, %d features per plane (the stereo configs)", kHexW, kHexC);
  if (a->N == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->ld >= a->N && a->ld % 128 == 0, "ld=%d must be a multiple of 128 and >= N=%d", a->ld, a->N);
  MOBGS_REQUIRE(a->pts && a->rots && a->times && a->w0 && a->b0 && a->wa && a->ba && a->wb && a->bb, "NULL input");
  MOBGS_REQUIRE(a->w0_t && a->wa_t && a->wb_t, "NULL transposed weight tiles");
  MOBGS_REQUIRE(a->g_pts && a->g_scales && a->g_rots && a->g_feat, "NULL gradient output");
  MOBGS_REQUIRE(a->featT && a->a1T && a->a2T && a->gz1T && a->goT && a->gh0T, "NULL feature-major scratch");
  for (int i = 0; i < a->levels * 6; ++i)
    MOBGS_REQUIRE(a->planes[i] && a->plane_w[i] >= 1 && a->plane_h[i] >= 1, "bad plane %d", i);
  const size_t smem = sizeof(HexSmem) + 128;
  cudaFuncSetAttribute(hexplane_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = a->ld / kHexRows;          // the padded tail is written as zeros
  hexplane_mlp_bwd_kernel<<<grid, kHexThreads, smem, (cudaStream_t)stream>>>(*a);
  return check_launch("hexplane_mlp_bwd");
}

extern "C" int mobgs_hexplane_wgrad(const MobgsHexWgrad* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->n_problems >= 0 && a->n_problems <= MOBGS_WGRAD_MAX, "n_problems out of range");
  if (a->n_problems == 0 || a->ld == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->ld > 0 && a->ld % 128 == 0, "ld=%d must be a positive multiple of 128", a->ld);
  for (int p = 0; p < a->n_problems; ++p) {
    MOBGS_REQUIRE(a->A[p] && a->B[p] && a->C[p], "problem %d: NULL operand", p);
    MOBGS_REQUIRE(a->n_cols[p] >= 16 && a->n_cols[p] <= 128 && a->n_cols[p] % 16 == 0, "problem %d: n_cols=%d", p, a->n_cols[p]);
    MOBGS_REQUIRE(a->ldc[p] >= (a->transpose_out[p] ? 128 : a->n_cols[p]), "problem %d: ldc too small", p);
  }
  const size_t smem = sizeof(WgSmem) + 128;
  cudaFuncSetAttribute(hexplane_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int chunks = a->ld / kWgChunk;
  const int gx = chunks < 96 ? chunks : 96;
  hexplane_wgrad_kernel<<<dim3(gx, a->n_problems), 128, smem, (cudaStream_t)stream>>>(*a);
  return check_launch("hexplane_wgrad");
}
