// f4: densify / prune gather-compaction of every per-Gaussian tensor in ONE launch.
//
// The reference rebuilds each parameter and both Adam moments with a separate boolean-mask index
// (`_prune_optimizer`, scene/gaussian_model.py:1044-1069: param[mask], exp_avg[mask], exp_avg_sq[mask] per group,
// ~45 launches + 45 nonzero() syncs) or torch.cat (`cat_tensors_to_optimizer`, :1094-1123).  Both are row
// gathers with one shared index list:  out[i] = src[idx[i]]            if idx[i] <  n_old
//                                              = ext[idx[i] - n_old]    if idx[i] >= n_old   (ext NULL -> 0)
// prune: idx = the kept rows in ascending order; append: idx = 0..n_old+n_new-1 (ext = the new rows, NULL for
// the Adam moments, which the reference extends with zeros).  Rows are copied as 32-bit words, so fp32 rows
// and the int64 `current_control_num` rows (2 words) go through the same path, bit-exact.
// HBM-bound: (row bytes read + row bytes written) per kept row; chunk table as in adam.cu.
#include "common.cuh"

namespace mobgs {

constexpr int kCompactThreads = 256;
constexpr int kCompactChunk = 4096;   // output words per chunk

__global__ void __launch_bounds__(kCompactThreads) compact_rows_kernel(const __grid_constant__ MobgsCompactRows a) {
  __shared__ int s_begin[MOBGS_COMPACT_MAX_TENSORS + 1];
  for (int i = threadIdx.x; i <= a.n_tensors; i += blockDim.x) s_begin[i] = a.chunk_begin[i];
  __syncthreads();
  const int total = s_begin[a.n_tensors];
  for (int chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
    int lo = 0, hi = a.n_tensors - 1;              // tensor owning this chunk
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_begin[mid] <= chunk) lo = mid; else hi = mid - 1;
    }
    const int t = lo;
    const int rw = a.row_words[t];
    const uint32_t* __restrict__ src = reinterpret_cast<const uint32_t*>(a.src[t]);
    const uint32_t* __restrict__ ext = reinterpret_cast<const uint32_t*>(a.ext[t]);
    uint32_t* __restrict__ dst = reinterpret_cast<uint32_t*>(a.dst[t]);
    const int64_t words = (int64_t)a.n_out * rw;
    const int64_t w0 = (int64_t)(chunk - s_begin[t]) * kCompactChunk;
    const int64_t w1 = min(words, w0 + kCompactChunk);
    for (int64_t w = w0 + threadIdx.x; w < w1; w += kCompactThreads) {
      const int64_t i = w / rw;
      const int c = (int)(w - i * rw);
      const int64_t r = a.idx[i];
      uint32_t v = 0u;
      if (r < a.n_old) v = __ldg(src + r * rw + c);
      else if (ext) v = __ldg(ext + (r - a.n_old) * rw + c);
      dst[w] = v;
    }
  }
}

__global__ void __launch_bounds__(kCompactThreads) copy_segments_kernel(const __grid_constant__ MobgsCopySegments a) {
  __shared__ int s_begin[MOBGS_COPY_MAX_SEGMENTS + 1];
  for (int i = threadIdx.x; i <= a.n_segments; i += blockDim.x) s_begin[i] = a.chunk_begin[i];
  __syncthreads();
  const int total = s_begin[a.n_segments];
  for (int chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
    int lo = 0, hi = a.n_segments - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_begin[mid] <= chunk) lo = mid; else hi = mid - 1;
    }
    const uint32_t* __restrict__ src = reinterpret_cast<const uint32_t*>(a.src[lo]);
    uint32_t* __restrict__ dst = reinterpret_cast<uint32_t*>(a.dst[lo]);
    const int64_t w0 = (int64_t)(chunk - s_begin[lo]) * kCompactChunk;
    const int64_t w1 = min(a.n_words[lo], w0 + kCompactChunk);
    for (int64_t w = w0 + threadIdx.x; w < w1; w += kCompactThreads) dst[w] = src[w];
  }
}

}  // namespace mobgs

using namespace mobgs;

extern "C" int mobgs_copy_segments(const MobgsCopySegments* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->n_segments >= 0 && a->n_segments <= MOBGS_COPY_MAX_SEGMENTS, "n_segments out of range");
  if (a->n_segments == 0) return MOBGS_OK;
  for (int i = 0; i < a->n_segments; ++i)
    MOBGS_REQUIRE(a->n_words[i] == 0 || (a->src[i] && a->dst[i]), "segment %d: NULL pointer", i);
  const int total = a->chunk_begin[a->n_segments];
  if (total <= 0) return MOBGS_OK;
  const int grid = total < 148 * 8 ? total : 148 * 8;
  copy_segments_kernel<<<grid, kCompactThreads, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("copy_segments");
}

extern "C" int mobgs_compact_chunk_words(void) { return kCompactChunk; }

extern "C" int mobgs_compact_rows(const MobgsCompactRows* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->n_tensors >= 0 && a->n_tensors <= MOBGS_COMPACT_MAX_TENSORS, "n_tensors out of range");
  MOBGS_REQUIRE(a->n_out >= 0 && a->n_old >= 0, "bad row counts");
  if (a->n_tensors == 0 || a->n_out == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->idx, "NULL index list");
  for (int t = 0; t < a->n_tensors; ++t) {
    MOBGS_REQUIRE(a->dst[t] && a->row_words[t] > 0, "tensor %d: NULL destination or empty rows", t);
    MOBGS_REQUIRE(a->src[t] || a->n_old == 0, "tensor %d: NULL source", t);
  }
  const int total = a->chunk_begin[a->n_tensors];
  if (total <= 0) return MOBGS_OK;
  const int grid = total < 148 * 8 ? total : 148 * 8;
  compact_rows_kernel<<<grid, kCompactThreads, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("compact_rows");
}
