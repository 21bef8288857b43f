// a11: operand packing of the HexPlane + MLP kernels ON THE DEVICE, one launch.
//
// The tcgen05 forward / backward (hexplane_mlp.cu, hexplane_mlp_bwd.cu) read their weights pre-tiled into the
// shared-memory operand layout ({tf32 hi, fp32 remainder lo} x [K/4][rows][4]) and the planes channels-last
// ([H][W][32]).  The parameters change every optimiser step, so the packing runs once per step; done with torch ops
// it was ~200 tiny launches (masking, subtracting, reshaping, concatenating 25 weight tiles and permuting 18 planes:
// about 1 ms of host time per training step).  Here every tile, bias vector and plane is one JOB of a single
// table-driven launch (chunk table as in compact.cu / adam.cu).  Bit-exact with the torch formulation it replaces
// (hi = bits & 0xffffe000, lo = w - hi).
#include "common.cuh"

namespace mobgs {

constexpr int kPackThreads = 256;
constexpr int kPackChunk = 2048;   // logical elements per chunk

__device__ __forceinline__ float pack_src(const MobgsPackJob& j, int r, int c) {
  return (r < j.valid_rows && c < j.valid_cols) ? __ldg(j.src + (int64_t)r * j.row_stride + (int64_t)c * j.col_stride) : 0.f;
}

__global__ void __launch_bounds__(kPackThreads) pack_operands_kernel(const MobgsPackJob* __restrict__ jobs,
                                                                     const int32_t* __restrict__ chunk_begin, int n_jobs) {
  __shared__ int s_begin[MOBGS_PACK_MAX_JOBS + 1];
  for (int i = threadIdx.x; i <= n_jobs; i += blockDim.x) s_begin[i] = chunk_begin[i];
  __syncthreads();
  const int total = s_begin[n_jobs];
  for (int chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_begin[mid] <= chunk) lo = mid; else hi = mid - 1;
    }
    const MobgsPackJob j = jobs[lo];
    const int64_t n = (int64_t)j.rows * j.cols;
    const int64_t e0 = (int64_t)(chunk - s_begin[lo]) * kPackChunk, e1 = min(n, e0 + kPackChunk);
    for (int64_t e = e0 + threadIdx.x; e < e1; e += kPackThreads) {
      if (j.kind == MOBGS_PACK_TILED) {
        // out[(k4 * rows + r) * 4 + q] = m[r][4 k4 + q];  hi block, then lo block
        const int q = (int)(e & 3);
        const int64_t t = e >> 2;
        const int r = (int)(t % j.rows), k4 = (int)(t / j.rows);
        const float v = pack_src(j, r, 4 * k4 + q);
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);     // the 10 mantissa bits tf32 keeps
        j.dst[e] = h;
        j.dst[n + e] = v - h;
      } else if (j.kind == MOBGS_PACK_PLAIN) {
        const int r = (int)(e / j.cols), c = (int)(e - (int64_t)r * j.cols);
        j.dst[e] = pack_src(j, r, c);
      } else {                       // MOBGS_PACK_TRANSPOSE: dst [cols][rows] = src [rows][cols] (channels-last planes)
        const int c = (int)(e / j.rows), r = (int)(e - (int64_t)c * j.rows);
        j.dst[e] = pack_src(j, r, c);
      }
    }
  }
}

}  // namespace mobgs

using namespace mobgs;

extern "C" int mobgs_pack_chunk_elems(void) { return kPackChunk; }

extern "C" int mobgs_pack_operands(const MobgsPackOperands* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->n_jobs >= 0 && a->n_jobs <= MOBGS_PACK_MAX_JOBS, "n_jobs out of range");
  if (a->n_jobs == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->jobs && a->chunk_begin && a->n_chunks > 0, "NULL job table / no chunks");
  const int grid = a->n_chunks < 148 * 8 ? a->n_chunks : 148 * 8;
  pack_operands_kernel<<<grid, kPackThreads, 0, (cudaStream_t)stream>>>(a->jobs, a->chunk_begin, a->n_jobs);
  return check_launch("pack_operands");
}
