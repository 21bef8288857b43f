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
#include "common.cuh"
#include "tma.cuh"

namespace mobgs {

constexpr int kHexThreads = 128;       // one thread = one point = one TMEM lane
constexpr int kHexRows = 128;
constexpr int kHexW = 128;             // net_width (arguments/stereo/default.py:9)
constexpr int kHexC = 32;              // features per plane (output_coordinate_dim)
constexpr int kMaxLevels = 4;
constexpr int kTmemCols = 256;         // [0,128) hidden, [128,256) layer outputs
constexpr float kLog100 = 4.605170185988092f;

// ---- PTX wrappers (mbarrier / TMA bulk copy helpers live in tma.cuh) ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// element (row r, k) lives at  start + (r%8)*16 + (r/8)*SBO + (k/4)*LBO + (k%4)*4  bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128 (cute::UMMA::InstrDescriptor)
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kHexRows >> 4) << 24);
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// ---- shared-memory plan ---------------------------------------------------------------------
// operand tiles are [K/4][rows][4] floats (= core matrices of 8 rows x 16 B, SBO = 128 B,
// LBO = rows * 16 B)
struct HexSmem {
  float a_hi[kHexW / 4 * kHexRows * 4];   // 64 KB
  float a_lo[kHexW / 4 * kHexRows * 4];   // 64 KB
  float b_hi[kHexW / 4 * 64 * 4];         // 32 KB (64 output rows per pass)
  float b_lo[kHexW / 4 * 64 * 4];         // 32 KB
  uint64_t bar_w;                         // weight tile landed (TMA complete_tx)
  uint64_t bar_mma;                       // MMAs retired (tcgen05.commit)
  uint32_t tmem_base;
};

__device__ __forceinline__ void store_a(HexSmem& sm, int row, int k4, float4 v) {
  const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
  const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
  reinterpret_cast<float4*>(sm.a_hi)[k4 * kHexRows + row] = hi;
  reinterpret_cast<float4*>(sm.a_lo)[k4 * kHexRows + row] = lo;
}

struct PlaneSample { int o00, o01, o10, o11; float w00, w01, w10, w11; };

// F.grid_sample(bilinear, align_corners=True, padding_mode='border') on a channels-last plane
__device__ __forceinline__ PlaneSample plane_sample(float x, float y, int Wd, int Hd) {
  float ix = (x + 1.f) * 0.5f * (float)(Wd - 1), iy = (y + 1.f) * 0.5f * (float)(Hd - 1);
  ix = fminf(fmaxf(ix, 0.f), (float)(Wd - 1));
  iy = fminf(fmaxf(iy, 0.f), (float)(Hd - 1));
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const int x1 = min(x0 + 1, Wd - 1), y1 = min(y0 + 1, Hd - 1);
  const float tx = ix - fx, ty = iy - fy;
  PlaneSample s;
  s.o00 = (y0 * Wd + x0) * kHexC; s.o01 = (y0 * Wd + x1) * kHexC;
  s.o10 = (y1 * Wd + x0) * kHexC; s.o11 = (y1 * Wd + x1) * kHexC;
  s.w00 = (1.f - tx) * (1.f - ty); s.w01 = tx * (1.f - ty);
  s.w10 = (1.f - tx) * ty;         s.w11 = tx * ty;
  return s;
}

// One logical GEMM pass: D[:, n0:n0+n) (+)= A(128 x K) * Btile(n x K)^T with the 3xTF32 split.
__device__ __forceinline__ void issue_gemm(HexSmem& sm, uint32_t tmem_d, int K, int n) {
  const uint32_t idesc = make_idesc(n);
  const uint32_t a_lbo = kHexRows * 16, b_lbo = (uint32_t)n * 16;
  const uint32_t a_hi = smem_u32(sm.a_hi), a_lo = smem_u32(sm.a_lo);
  const uint32_t b_hi = smem_u32(sm.b_hi), b_lo = smem_u32(sm.b_lo);
  uint32_t acc = 0;
  for (int k8 = 0; k8 < K / 8; ++k8) {          // one MMA = K 8 = two 16-byte k-chunks
    const uint32_t ao = (uint32_t)k8 * 2 * a_lbo, bo = (uint32_t)k8 * 2 * b_lbo;
    umma_tf32(tmem_d, make_desc(a_hi + ao, a_lbo, 128), make_desc(b_hi + bo, b_lbo, 128), idesc, acc);
    acc = 1;
    umma_tf32(tmem_d, make_desc(a_lo + ao, a_lbo, 128), make_desc(b_hi + bo, b_lbo, 128), idesc, 1);
    umma_tf32(tmem_d, make_desc(a_hi + ao, a_lbo, 128), make_desc(b_lo + bo, b_lbo, 128), idesc, 1);
  }
}

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
