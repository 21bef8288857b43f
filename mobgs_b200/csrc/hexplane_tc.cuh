// tcgen05 / TMEM building blocks shared by the HexPlane + MLP forward (hexplane_mlp.cu) and backward
// (hexplane_mlp_bwd.cu) kernels: PTX wrappers, the K-major no-swizzle operand layout, the 3xTF32 GEMM pass.
#pragma once
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


}  // namespace mobgs
