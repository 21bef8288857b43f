// Flow records for the K-batched get_flow() (reference
// gaussian_renderer/__init__.py:411-471): one pass builds, for all K exposure offsets, the record
// sets whose two colour channels are the screen-space displacement between the mid-time and the
// exposure-time projection of every Gaussian.  One thread = one Gaussian, looping over k, so the
// mid-time record is read once and (in the VJP) its gradient is summed in registers — no atomics.
#include "common.cuh"

namespace mobgs {

__global__ void __launch_bounds__(256) flow_records_fwd_kernel(MobgsFlowRecFwd a) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.N) return;
  const float4* rec = reinterpret_cast<const float4*>(a.records);
  float4* out = reinterpret_cast<float4*>(a.flow_records);
  const size_t N4 = (size_t)a.N * 4;
  const float4 m0 = rec[(size_t)g * 4], m1 = rec[(size_t)g * 4 + 1];
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < a.K; ++k) {
    const float4 e0 = rec[(size_t)(k + 1) * N4 + (size_t)g * 4], e1 = rec[(size_t)(k + 1) * N4 + (size_t)g * 4 + 1];
    const float fx = m0.x - e0.x, fy = m0.y - e0.y;          // mid - exp
    float4* oe = out + (size_t)(2 * k) * N4 + (size_t)g * 4;
    float4* om = out + (size_t)(2 * k + 1) * N4 + (size_t)g * 4;
    oe[0] = e0; oe[1] = make_float4(e1.x, e1.y, fx, fy); oe[2] = z; oe[3] = z;
    om[0] = m0; om[1] = make_float4(m1.x, m1.y, -fx, -fy); om[2] = z; om[3] = z;
  }
}

__global__ void __launch_bounds__(256) flow_records_bwd_kernel(MobgsFlowRecBwd a) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.N) return;
  const float4* v = reinterpret_cast<const float4*>(a.v_flow_records);
  float4* out = reinterpret_cast<float4*>(a.v_records);
  const size_t N4 = (size_t)a.N * 4;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 gm0 = z, gm1 = z;
  for (int k = 0; k < a.K; ++k) {
    const float4* ve = v + (size_t)(2 * k) * N4 + (size_t)g * 4;
    const float4* vm = v + (size_t)(2 * k + 1) * N4 + (size_t)g * 4;
    const float4 e0 = ve[0], e1 = ve[1], q0 = vm[0], q1 = vm[1];
    // d loss / d (mid - exp) from both flow renders
    const float dfx = e1.z - q1.z, dfy = e1.w - q1.w;
    float4* oe = out + (size_t)(k + 1) * N4 + (size_t)g * 4;
    oe[0] = make_float4(e0.x - dfx, e0.y - dfy, e0.z, e0.w);
    oe[1] = make_float4(e1.x, e1.y, 0.f, 0.f);
    oe[2] = z; oe[3] = z;
    gm0.x += q0.x + dfx; gm0.y += q0.y + dfy; gm0.z += q0.z; gm0.w += q0.w;
    gm1.x += q1.x; gm1.y += q1.y;
  }
  float4* om = out + (size_t)g * 4;
  om[0] = gm0; om[1] = gm1; om[2] = z; om[3] = z;
}

}  // namespace mobgs

using namespace mobgs;

extern "C" int mobgs_flow_records_fwd(const MobgsFlowRecFwd* a, void* stream) {
  MOBGS_REQUIRE(a && a->K >= 1 && a->N >= 0, "bad arguments");
  if (a->N == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->records && a->flow_records, "NULL pointer");
  flow_records_fwd_kernel<<<(a->N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("flow_records_fwd");
}

extern "C" int mobgs_flow_records_bwd(const MobgsFlowRecBwd* a, void* stream) {
  MOBGS_REQUIRE(a && a->K >= 1 && a->N >= 0, "bad arguments");
  if (a->N == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->v_flow_records && a->v_records, "NULL pointer");
  flow_records_bwd_kernel<<<(a->N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("flow_records_bwd");
}
