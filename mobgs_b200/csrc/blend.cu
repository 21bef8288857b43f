// K6'/K7': front-to-back alpha compositing, forward and VJP.
//
// One CTA = one 16x16 tile of one sub-frame, one thread = one pixel.  The tile's depth-sorted list is
// consumed in batches: every thread stages one packed 64-byte record with a TMA bulk copy (cp.async.bulk
// -> mbarrier) into shared memory — geometry *and* colours (gsplat re-reads colours from global memory
// per (pixel, Gaussian) pair and pads D=10 to 16 channels; here the channel count is a template
// parameter) — and computes which of the tile's sixteen 4x4-pixel *units* the Gaussian can reach with
// alpha >= 1/255 (blend_units.cuh, exact).  Each warp compacts the batch into one list per unit, so its two
// half-warps walk different Gaussians at the same time and skip everything that cannot touch them.
//
// Forward (blend_fwd_kernel): per-lane loop over the unit's list, LDS.128 broadcast of the record, ex2.approx;
// optional fused epilogue: expected depth + Sandwich decoder while the pixel is in registers.
//
// Backward (blend_bwd_tr_kernel): back to front with the scalar suffix S instead of gsplat's per-channel
// buffer.  The sum over pixels is done by changing ownership: Phase A (lane = pixel) parks fac = alpha T and
// v_sigma of 8 list entries in a per-warp matrix, Phase B (lane = entry x half of the unit's pixels) turns a
// row of it into the 16 gradient sums of that (unit, Gaussian) and sends them to global memory as 2 x
// RED.128 per lane — gsplat issues 16 scalar atomics per (warp, Gaussian) after 5-step shuffle reductions.
// blend_bwd_kernel (transposing-butterfly reductions + shared accumulator) is the earlier variant, kept for
// the ablation build (-DMOBGS_BWD_TRANSPOSE=0) and for unit sizes other than 16 lanes.
#include "common.cuh"
#include "blend_units.cuh"
#include "decode_math.cuh"
#include "tma.cuh"
#include "ray_math.cuh"

namespace mobgs {

constexpr int kBlendThreads = kTilePix;   // 256
#ifndef MOBGS_FWD_MIN_CTAS
#define MOBGS_FWD_MIN_CTAS 6
#endif
// 1: stage each batch of records with per-record TMA bulk copies (cp.async.bulk -> UBLKCP) that
// complete on an mbarrier; 0: LDG.128 -> STS.128 through registers.
#ifndef MOBGS_TMA_STAGE
#define MOBGS_TMA_STAGE 1
#endif
// 1: branch-free forward body in blocks of 8 entries (measured 9 % SLOWER than the divergent loop — the
// per-lane `continue` skips the colour half of most iterations, and the variant spills at 40 registers);
// kept for the ablation build only.
#ifndef MOBGS_FWD_PREDICATED
#define MOBGS_FWD_PREDICATED 0
#endif
#ifndef MOBGS_BWD_PREDICATED
#define MOBGS_BWD_PREDICATED 1
#endif
#ifndef MOBGS_BWD_MIN_CTAS
#define MOBGS_BWD_MIN_CTAS 4
#endif
// 1: decoder weight gradients of the backward prologue as 2 x 2 register blocks joined by shuffles (see
// bwd_pixel_prologue); 0: one output per thread + shared-memory atomics (ablation)
#ifndef MOBGS_BWD_WGRAD_BLOCKED
#define MOBGS_BWD_WGRAD_BLOCKED 1
#endif

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {   // one MUFU.RCP (|rel err| <= 2^-22); x in [1e-3, 1] at its call sites
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kNegLog2e = -1.4426950408889634f;
// The staged conic is rescaled once per entry to (A, B, C) = -log2(e) (a/2, b, c/2), so that the exponent of
// alpha = opac 2^p is p = dx (A dx + B dy) + C dy^2: five instructions per (pixel, Gaussian) pair instead of eight.
constexpr float kConicDiag = 0.5f * kNegLog2e, kConicOff = kNegLog2e;
__device__ __forceinline__ void rescale_conic(float* rec) {   // rec: x y opac ca | cb cc ...
  rec[3] *= kConicDiag;
  *reinterpret_cast<float2*>(rec + 4) = make_float2(rec[4] * kConicOff, rec[5] * kConicDiag);
}
__device__ __forceinline__ float pair_exponent(const float4& r0, const float4& r1, float dx, float dy) {
  return dx * (r0.w * dx + r1.x * dy) + (r1.y * dy) * dy;
}

// Sub-warp units (blend_units.cuh): MOBGS_UNIT_LANES = 32 / 16 / 8 lanes per unit.
#ifndef MOBGS_UNIT_LANES
#define MOBGS_UNIT_LANES 16
#endif
constexpr int kUL = MOBGS_UNIT_LANES;              // lanes (= pixels) per unit
using UG = UnitGeom<kUL>;
constexpr int kUW = UG::kUW, kUH = UG::kUH, kUX = UG::kUX, kUnits = UG::kUnits, kUPW = UG::kUPW;

// Ablation switches (tools/ablation.sh): rebuild with -DMOBGS_ABL_NO_STRIP_MASK / -DMOBGS_ABL_NAIVE_REDUCE
// to measure what the unit culling and the transposing butterfly buy.  Never set in the product build.
__device__ __forceinline__ unsigned unit_mask(const float4& r0, const float4& r1, float tile_x0, float tile_y0) {
#ifdef MOBGS_ABL_NO_STRIP_MASK
  return UG::kAll;
#else
  return unit_mask<kUL>(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, tile_x0, tile_y0);
#endif
}

__device__ __forceinline__ float rec_color(const float4& r1, const float4& r2, const float4& r3, int c) {
  switch (c) {
    case 0: return r1.z; case 1: return r1.w;
    case 2: return r2.x; case 3: return r2.y; case 4: return r2.z; case 5: return r2.w;
    case 6: return r3.x; case 7: return r3.y; case 8: return r3.z; default: return r3.w;
  }
}

// Each warp compacts, for each of its kUPW units, the batch entries whose mask includes the unit
// (ascending order kept) into the unit's byte list, so a unit's inner loop visits only Gaussians that
// can reach its pixels.  `wl` = the list of the warp's first unit (lists are STRIDE bytes
// apart), `t_begin` = first batch slot this lane's unit still needs.  Returns this lane's unit's count.
template <int STRIDE = kBlendThreads>
__device__ __forceinline__ int build_unit_lists(const unsigned* smask, unsigned char* wl, int warp,
                                                int lane, int t_begin, int bn) {
  int cnt[kUPW], tb[kUPW];
#pragma unroll
  for (int s = 0; s < kUPW; ++s) {
    cnt[s] = 0;
    tb[s] = kUPW == 1 ? t_begin : __shfl_sync(0xffffffffu, t_begin, s * kUL);
  }
  const unsigned lt = (1u << lane) - 1u;
  for (int t0 = 0; t0 < bn; t0 += 32) {
    const int t = t0 + lane;
    const unsigned m = t < bn ? smask[t] >> (warp * kUPW) : 0u;
#pragma unroll
    for (int s = 0; s < kUPW; ++s) {
      const bool hit = ((m >> s) & 1u) && t >= tb[s];
      const unsigned b = __ballot_sync(0xffffffffu, hit);
      if (hit) wl[s * STRIDE + cnt[s] + __popc(b & lt)] = (unsigned char)t;
      cnt[s] += __popc(b);
    }
  }
  __syncwarp();
  int mine = cnt[0];
#pragma unroll
  for (int s = 1; s < kUPW; ++s) mine = (lane / kUL) == s ? cnt[s] : mine;
  return mine;
}

// Gaussian index of list entry `idx`, clamped to [0, N): when a speculative list capacity overflows, the attempt's lists
// hold unwritten entries (the caller discards it and redoes the chain with the exact size), and this kernel — already
// enqueued on them — must still dereference valid records.
__device__ __forceinline__ int list_gid(const int32_t* sorted_ids, int idx, int N) {
  return (int)min((unsigned)sorted_ids[idx], (unsigned)(N - 1));
}

// first / one-past-last entry of the tile lists walked by CTA (list k, tile): lists may share a binning
__device__ __forceinline__ int tile_segment(const MobgsLists& l, int k, int tiles, int tile) {
  return l.tile_list[k] * tiles + tile;
}

__device__ __forceinline__ int lane_id() {
  int l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

// Transposing butterfly over NV (8 or 16) per-lane values within each unit of kUL lanes: every
// stage halves the number of live values per lane while doubling the lanes summed, so NV values cost
// about NV shuffles in total (not log2(kUL) * NV).  Afterwards g[0 .. R-1] (R = max(1, NV / kUL)) of
// lane l hold the complete unit sums of value indices vidx .. vidx + R - 1; when NV < kUL only lanes
// with (l & (kUL / NV - 1)) == 0 own distinct values.
template <int NV>
__device__ __forceinline__ int butterfly_reduce(float (&g)[NV], int lane) {
  int vidx = 0;
  int n = NV / 2;
#pragma unroll
  for (int o = kUL / 2; o >= 1; o >>= 1) {
    if (n >= 1) {
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) {
        if (i < n) {
          const float send = up ? g[i] : g[i + n];
          const float keep = up ? g[i + n] : g[i];
          g[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      vidx += up ? n : 0;
      n >>= 1;
    } else {
      g[0] += __shfl_xor_sync(0xffffffffu, g[0], o);
    }
  }
  return vidx;
}

// tile-local pixel of thread `tid`: unit u = tid / kUL at (u % kUX, u / kUX), lane q = tid % kUL inside it
__device__ __forceinline__ void unit_pixel(int tid, int& lx, int& ly) {
  const int u = tid / kUL, q = tid % kUL;
  lx = (u % kUX) * kUW + (q % kUW);
  ly = (u / kUX) * kUH + (q / kUW);
}

// POSE: camera rays of the fused epilogue generated in registers from a.dec_pose (ray_math.cuh) instead of read
// from a.dec_rays — a template parameter so that neither variant carries the other's code and registers.
template <int D, bool DEC, bool FLOW = false, bool POSE = false>
__global__ void __launch_bounds__(kBlendThreads, FLOW ? MOBGS_FWD_MIN_CTAS - 1 : MOBGS_FWD_MIN_CTAS) blend_fwd_kernel(const __grid_constant__ MobgsBlendFwd a, int tiles_x, int tiles_y) {
  static_assert(!FLOW || (DEC && !MOBGS_FWD_PREDICATED), "fused flow channels ride the decode launch of the divergent body");
  static_assert(!POSE || DEC, "in-kernel rays belong to the fused decode epilogue");
  __shared__ __align__(128) float4 srec[kBlendThreads][4];
  __shared__ __align__(8) float2 sflow[FLOW ? kBlendThreads : 1];   // per staged entry: records[flow_ref][g].xy - own xy
  __shared__ unsigned smask[kBlendThreads];
  __shared__ __align__(8) unsigned char swl[kUnits][kBlendThreads + 8];   // +8: 8-byte loads of a warp's units differ in bank
  __shared__ __align__(16) float sdec[DEC ? (POSE ? 112 : 96) : 4];   // decoder weights [0,90) | camera pose [96,108)
  __shared__ __align__(8) uint64_t sbar;
  const int tiles = tiles_x * tiles_y;
  const int k = blockIdx.x / tiles, tile = blockIdx.x - k * tiles;
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int tid = threadIdx.x;
  const int lane = lane_id();
  const unsigned char* ulist = swl[tid / kUL];          // this lane's unit's list
  unsigned char* wl0 = swl[(tid >> 5) * kUPW];          // list of the warp's first unit
  uint32_t bar_phase = 0;
  if (MOBGS_TMA_STAGE && tid == 0) {
    mbar_init(&sbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // made visible by the first barrier below
  }
  if (DEC && tid < 90) sdec[tid] = tid < 72 ? a.dec_w1[tid] : a.dec_w2[tid - 72];   // visible after the first barrier
  if (POSE && tid >= 96 && tid < 108) {
    const int rk = a.dec_rays_per_k == 2 ? a.lists.rec_k[k] : (a.dec_rays_per_k ? k : 0);
    sdec[tid] = a.dec_pose[12 * rk + (tid - 96)];
  }
  int lx, ly;
  unit_pixel(tid, lx, ly);
  const int ix = tx * kTile + lx, iy = ty * kTile + ly;
  const bool inside = ix < a.width && iy < a.height;
  const float px = ix + 0.5f, py = iy + 0.5f;
  const int cap = (int)min(a.list_capacity, (int64_t)0x7fffffff);
  const int seg = tile_segment(a.lists, k, tiles, tile);
  const int beg = min(a.tile_offsets[seg], cap), end = min(a.tile_offsets[seg + 1], cap);
  const float4* recs = reinterpret_cast<const float4*>(a.records) + (size_t)a.lists.rec_k[k] * a.N * 4;
  const float4* recs_ref = FLOW ? reinterpret_cast<const float4*>(a.records) + (size_t)a.flow_ref * a.N * 4 : nullptr;

  float T = 1.f;
  float fl0 = 0.f, fl1 = 0.f;
  float pix[D];
#pragma unroll
  for (int c = 0; c < D; ++c) pix[c] = 0.f;
  int last = -1;
  bool done = !inside;

  for (int b0 = beg; b0 < end; b0 += kBlendThreads) {
    if (__syncthreads_count(done) == kBlendThreads) break;
    const int idx = b0 + tid;
    const int bn = min(kBlendThreads, end - b0);
#if MOBGS_TMA_STAGE
    constexpr uint32_t kRecBytes = D > 6 ? 64 : (D > 2 ? 48 : 32);
    if (tid == 0) mbar_expect_tx(&sbar, (uint32_t)bn * kRecBytes);
    float2 mref = make_float2(0.f, 0.f);
    if (idx < end) {
      const int g = list_gid(a.sorted_ids, idx, a.N);
      bulk_g2s(&srec[tid][0], recs + (size_t)g * 4, kRecBytes, &sbar);
      if (FLOW) mref = __ldg(reinterpret_cast<const float2*>(recs_ref + (size_t)g * 4));
    }
    mbar_wait(&sbar, bar_phase);
    bar_phase ^= 1;
    if (idx < end) {
      smask[tid] = unit_mask(srec[tid][0], srec[tid][1], (float)(tx * kTile), (float)(ty * kTile));
      if (kUnits <= 16 && a.list_masks) a.list_masks[idx] = (uint16_t)smask[tid];       // kept for the backward
      if (FLOW) sflow[tid] = make_float2(mref.x - srec[tid][0].x, mref.y - srec[tid][0].y);
      rescale_conic(reinterpret_cast<float*>(&srec[tid][0]));
    }
#else
    if (idx < end) {
      const int g = list_gid(a.sorted_ids, idx, a.N);
      const float4* r = recs + (size_t)g * 4;
      const float4 q0 = __ldg(r), q1 = __ldg(r + 1);
      srec[tid][0] = q0; srec[tid][1] = q1;
      if (D > 2) srec[tid][2] = __ldg(r + 2);
      if (D > 6) srec[tid][3] = __ldg(r + 3);
      smask[tid] = unit_mask(q0, q1, (float)(tx * kTile), (float)(ty * kTile));
      if (kUnits <= 16 && a.list_masks) a.list_masks[idx] = (uint16_t)smask[tid];
      if (FLOW) {
        const float2 mref = __ldg(reinterpret_cast<const float2*>(recs_ref + (size_t)g * 4));
        sflow[tid] = make_float2(mref.x - q0.x, mref.y - q0.y);
      }
      rescale_conic(reinterpret_cast<float*>(&srec[tid][0]));
    }
#endif
    __syncthreads();
    const int cnt = build_unit_lists<kBlendThreads + 8>(smask, wl0, tid >> 5, lane, 0, bn);
    int last_t = -1;
#if MOBGS_FWD_PREDICATED
    // Branch-free body, blocks of 8 list entries (one 8-byte list load): a lane whose pixel does not
    // blend the entry (culled, alpha < 1/255, already terminated, list exhausted) runs the same
    // arithmetic with weight 0; the warp leaves when all its lanes have terminated.
    int cnt_w = cnt;
#pragma unroll
    for (int o = 16; o >= kUL; o >>= 1) cnt_w = max(cnt_w, __shfl_xor_sync(0xffffffffu, cnt_w, o));
    for (int base = 0; base < cnt_w; base += 8) {
      if (__all_sync(0xffffffffu, done)) break;
      const uint2 tl = *reinterpret_cast<const uint2*>(ulist + base);   // stale past cnt: masked by `act`
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool act = base + i < cnt && !done;
        const int t = act ? (int)(((i < 4 ? tl.x : tl.y) >> (8 * (i & 3))) & 0xffu) : 0;
        const float4 r0 = srec[t][0], r1 = srec[t][1];
        const float dx = r0.x - px, dy = r0.y - py;
        const float pw = pair_exponent(r0, r1, dx, dy);            // = -sigma log2(e)
        const float alpha = fminf(kAlphaMax, r0.z * ex2_approx(pw));
        const bool valid = act && pw <= 0.f && alpha >= kAlphaMin;
        const float next_T = T * (1.f - alpha);
        const bool stop = valid && next_T <= kTStop;               // this entry is not blended
        done = done || stop;
        const bool use = valid && !stop;
        const float w = use ? alpha * T : 0.f;
        T = use ? next_T : T;
        last_t = use ? t : last_t;
        pix[0] += r1.z * w;
        if (D > 1) pix[1 % D] += r1.w * w;
        if (D > 2) {
          const float4 r2 = srec[t][2];
          pix[2 % D] += r2.x * w;
          if (D > 3) pix[3 % D] += r2.y * w;
          if (D > 4) pix[4 % D] += r2.z * w;
          if (D > 5) pix[5 % D] += r2.w * w;
        }
        if (D > 6) {
          const float4 r3 = srec[t][3];
          pix[6 % D] += r3.x * w;
          if (D > 7) pix[7 % D] += r3.y * w;
          if (D > 8) pix[8 % D] += r3.z * w;
          if (D > 9) pix[9 % D] += r3.w * w;
        }
      }
    }
#else
    // (a rolled loop: unrolling it over 8-byte list loads, as the backward does, was 12 % slower here —
    // the per-lane continue / break paths multiply)
    {
      for (int i = 0; i < cnt && !done; ++i) {
      const int t = ulist[i];
      const float4 r0 = srec[t][0], r1 = srec[t][1];
      const float dx = r0.x - px, dy = r0.y - py;
      const float pw = pair_exponent(r0, r1, dx, dy);            // = -sigma log2(e)
      const float alpha = fminf(kAlphaMax, r0.z * ex2_approx(pw));
      if (pw > 0.f || alpha < kAlphaMin) continue;
      const float next_T = T * (1.f - alpha);
      if (next_T <= kTStop) { done = true; break; }
      const float w = alpha * T;
      pix[0] += r1.z * w;
      if (D > 1) pix[1 % D] += r1.w * w;
      if (D > 2) {
        const float4 r2 = srec[t][2];
        pix[2 % D] += r2.x * w;
        if (D > 3) pix[3 % D] += r2.y * w;
        if (D > 4) pix[4 % D] += r2.z * w;
        if (D > 5) pix[5 % D] += r2.w * w;
      }
      if (D > 6) {
        const float4 r3 = srec[t][3];
        pix[6 % D] += r3.x * w;
        if (D > 7) pix[7 % D] += r3.y * w;
        if (D > 8) pix[8 % D] += r3.z * w;
        if (D > 9) pix[9 % D] += r3.w * w;
      }
      if (FLOW) {
        const float2 f = sflow[t];
        fl0 += f.x * w;
        fl1 += f.y * w;
      }
      last_t = t;
      T = next_T;
      }
    }
#endif
    if (last_t >= 0) last = b0 + last_t;
  }
  if (inside) {
    const size_t p = ((size_t)k * a.height + iy) * a.width + ix;
    a.out_alphas[p] = 1.f - T;
    a.last_idx[p] = last;
    float* oc = a.out_colors + p * D;
    const float* bg = a.backgrounds ? a.backgrounds + (size_t)k * D : nullptr;
#pragma unroll
    for (int c = 0; c < D; ++c) { pix[c] = bg ? pix[c] + T * bg[c] : pix[c]; oc[c] = pix[c]; }
  }
  if (DEC) {
    // fused epilogue: expected depth + Sandwich decoder on the pixel still held in registers
    __syncthreads();                       // sdec is complete even for tiles with an empty list
    if (inside) {
      const size_t P = (size_t)a.width * a.height, pp = (size_t)iy * a.width + ix;
      DecW w; w.w1 = sdec; w.w2 = sdec + (DEC ? 72 : 0);
      float v[10], rays[6], x[12], hpre[6], out[3];
#pragma unroll
      for (int c = 0; c < 10; ++c) v[c] = pix[c % D];
      if (POSE) {                  // the ray of this pixel from 12 pose floats: no [.,6,H,W] image is read
        const RayIntr in = {a.dec_ppx, a.dec_ppy, a.dec_sfx, a.dec_sfy};
        pixel_ray<true>(sdec + (POSE ? 96 : 0), in, ix, iy, rays);
      } else {
        const int rk = a.dec_rays_per_k == 2 ? a.lists.rec_k[k] : (a.dec_rays_per_k ? k : 0);
        const float* rp = a.dec_rays + (size_t)rk * 6 * P + pp;
#pragma unroll
        for (int i = 0; i < 6; ++i) rays[i] = __ldg(rp + i * P);
      }
      sandwich_fwd(w, v, rays, x, hpre, out);
#pragma unroll
      for (int c = 0; c < 3; ++c) a.out_rgb[((size_t)k * 3 + c) * P + pp] = out[c];
      if (a.out_depth) a.out_depth[(size_t)k * P + pp] = v[9] / fmaxf(1.f - T, kEdFloor);
      if (FLOW) *reinterpret_cast<float2*>(a.out_flow + ((size_t)k * P + pp) * 2) = make_float2(fl0, fl1);
    }
  }
}

// Start state of one pixel's backward walk: T_final, the last blended list entry, the gradient of the
// loss w.r.t. the pixel's D blended channels (v_c) and alpha (v_a).  Plain path: read from
// v_out_colors / v_out_alphas.  DEC: VJP of (expected depth, Sandwich decoder, sub-frame mean)
// evaluated here, with the decoder weight gradients of the whole tile summed into swg[90].
// sx [12][kProPad] and sg [15][kProPad] (ghpre 0..5 | gpre 6..8 | relu(h) 9..14), 16-byte aligned,
// are scratch that may alias buffers which are idle until the list walk starts.  Called by all
// threads of the CTA (contains barriers).
constexpr int kProPad = kBlendThreads + 4;
template <int D, bool DEC, bool POSE = false>
__device__ __forceinline__ void bwd_pixel_prologue(const MobgsBlendBwd& a, int k, int tid, bool inside, int ix, int iy,
                                                   float* sx, float* sg, float* sdec, float* swg,
                                                   float& T_final, int& last, float (&v_c)[D], float& v_a,
                                                   float* spose = nullptr) {   // [8 warps][12] scratch (POSE)
  T_final = 1.f; v_a = 0.f; last = -1;
#pragma unroll
  for (int c = 0; c < D; ++c) v_c[c] = 0.f;
  if (DEC) {
    if (tid < 96) { sdec[tid] = tid < 72 ? a.dec_w1[tid] : (tid < 90 ? a.dec_w2[tid - 72] : 0.f); swg[tid] = 0.f; }
    else if (POSE && tid < 112) {  // camera pose [96,108) and its gradient accumulator
      const int rk = a.dec_rays_per_k == 2 ? a.lists.rec_k[k] : (a.dec_rays_per_k ? k : 0);
      sdec[tid] = tid < 108 ? a.dec_pose[12 * rk + (tid - 96)] : 0.f;
      swg[tid] = 0.f;
    }
    __syncthreads();
  }
  if (inside) {
    const size_t p = ((size_t)k * a.height + iy) * a.width + ix;
    T_final = 1.f - a.out_alphas[p];
    last = a.last_idx[p];
    if (!DEC) {
      const float* vc = a.v_out_colors + p * D;
#pragma unroll
      for (int c = 0; c < D; ++c) v_c[c] = vc[c];
      if (a.v_out_alphas) v_a = a.v_out_alphas[p];
    }
  }
  if (DEC) {
    // fused prologue: VJP of (expected depth, Sandwich decoder, sub-frame mean) for this pixel.
    // The 90 decoder weight-gradient terms are products of per-pixel factors (x[12]; ghpre[6],
    // gpre[3], relu(h)[6]); every pixel parks its factors in shared memory as soon as they exist —
    // the record / accumulator buffers are idle during the prologue — and 180 threads then sum one
    // product each over half of the tile's pixels.  Live ranges are kept short on purpose: holding
    // the factors in registers spilled ~460 B per thread (4.9 GB of local-memory DRAM writes).
    constexpr int kPad = kProPad;                        // row stride: rows stay 16-byte aligned and start 4 banks apart
    const size_t P = (size_t)a.width * a.height, pp = (size_t)iy * a.width + ix;
    const float* w1 = sdec;
    const float* w2 = sdec + (DEC ? 72 : 0);
    const RayIntr rin = {a.dec_ppx, a.dec_ppy, a.dec_sfx, a.dec_sfy};
    const bool want_pose = POSE && a.v_pose_partial;
    float depth_acc = 0.f;
    float gray[6];                 // d loss / d rays of this pixel (pose mode)
#pragma unroll
    for (int i = 0; i < 6; ++i) gray[i] = 0.f;
    if (inside) {
      const size_t p = (size_t)k * P + pp;
      const int rk = a.dec_rays_per_k == 2 ? a.lists.rec_k[k] : (a.dec_rays_per_k ? k : 0);
      const float2* vp2 = reinterpret_cast<const float2*>(a.out_colors + p * 10);
      float albedo[3], hpre[6];
      {
        float x[12];
        const float2 t0 = __ldg(vp2), t1 = __ldg(vp2 + 1), t2 = __ldg(vp2 + 2), t3 = __ldg(vp2 + 3), t4 = __ldg(vp2 + 4);
        albedo[0] = t0.x; albedo[1] = t0.y; albedo[2] = t1.x;
        x[0] = t1.y; x[1] = t2.x; x[2] = t2.y; x[3] = t3.x; x[4] = t3.y; x[5] = t4.x;
        depth_acc = t4.y;
        if (POSE) {
          float r6[6];
          pixel_ray<true>(sdec + (POSE ? 96 : 0), rin, ix, iy, r6);
#pragma unroll
          for (int i = 0; i < 6; ++i) x[6 + i] = r6[i];
        } else {
          const float* rp = a.dec_rays + (size_t)rk * 6 * P + pp;
#pragma unroll
          for (int i = 0; i < 6; ++i) x[6 + i] = __ldg(rp + i * P);
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) sx[i * kPad + tid] = x[i];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          float h = 0.f;
#pragma unroll
          for (int i = 0; i < 12; ++i) h += w1[12 * j + i] * x[i];
          hpre[j] = h;
          sg[(9 + j) * kPad + tid] = fmaxf(h, 0.f);
        }
      }
      const int meanK = a.mean_K > 0 ? a.mean_K : a.K;
      const float invK = 1.0f / (float)meanK;
      float gpre[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float sc = albedo[c];
#pragma unroll
        for (int j = 0; j < 6; ++j) sc += w2[6 * c + j] * fmaxf(hpre[j], 0.f);
        const float o = 1.0f / (1.0f + expf(-sc));
        float go = (a.g_mean && k < meanK) ? __ldg(a.g_mean + c * P + pp) * invK : 0.f;
        if (a.g_rgb) go += __ldg(a.g_rgb + ((size_t)k * 3 + c) * P + pp);
        gpre[c] = go * o * (1.f - o);
        v_c[c % D] = gpre[c];
        sg[(6 + c) * kPad + tid] = gpre[c];
      }
      float ghpre[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const float gh = w2[j] * gpre[0] + w2[6 + j] * gpre[1] + w2[12 + j] * gpre[2];
        ghpre[j] = hpre[j] > 0.f ? gh : 0.f;
        sg[j * kPad + tid] = ghpre[j];
      }
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        float gx = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) gx += w1[12 * j + i] * ghpre[j];
        if (i < 6) {
          v_c[(3 + i) % D] = gx;
        } else if (POSE) {
          gray[i - 6] = gx;
        } else if (a.v_rays) {
          if (a.dec_rays_per_k == 1) a.v_rays[((size_t)k * 6 + (i - 6)) * P + pp] = gx;
          else atomicAdd(a.v_rays + ((size_t)rk * 6 + (i - 6)) * P + pp, gx);   // rays shared between lists
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 12; ++i) sx[i * kPad + tid] = 0.f;
#pragma unroll
      for (int i = 0; i < 15; ++i) sg[i * kPad + tid] = 0.f;
    }
    if (POSE && want_pose) {
      // pose gradient (what mobgs_camera_rays_bwd reduces from a [.,6,H,W] image): 12 sums over the tile's pixels,
      // warp shuffles -> CTA accumulator swg[96..107] -> one atomicAdd per CTA and value (after the barrier below)
      // (transposing butterfly: 16 shuffles per warp instead of 60, the 12 sums land in 12 different lanes and are parked
      // per warp without atomics; 12 threads join the 8 warps after the barrier below)
      float v12[12], v16[16];
      pixel_ray_vjp<true>(sdec + (POSE ? 96 : 0), rin, ix, iy, gray, v12);
#pragma unroll
      for (int i = 0; i < 16; ++i) v16[i] = i < 12 ? v12[i % 12] : 0.f;
      const int vidx = warp_transpose_sum16(v16, tid & 31);
      if ((tid & 1) == 0 && vidx < 12) spose[(tid >> 5) * 12 + vidx] = v16[0];
    }
    if (inside) {
      const size_t p = (size_t)k * P + pp;
      const float al = 1.f - T_final, den = fmaxf(al, kEdFloor);
      const float gd = a.g_depth ? __ldg(a.g_depth + p) : 0.f;
      v_c[9 % D] = gd / den;
      v_a = (a.g_alpha ? __ldg(a.g_alpha + p) : 0.f) + (al > kEdFloor ? -gd * depth_acc / (den * den) : 0.f);
    }
    __syncthreads();
    if (POSE && want_pose && tid >= 96 && tid < 108) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < kBlendThreads / 32; ++w) sum += spose[w * 12 + (tid - 96)];
      swg[tid] = sum;
    }
#if MOBGS_BWD_WGRAD_BLOCKED
    // The 90 weight-gradient sums over the tile's 256 pixels as 24 register blocks of 2 x 2 outputs (two factor rows a,
    // two factor rows b: 4 LDS.128 feed 16 multiply-adds, where one output per thread needed 8), each block summed by
    // the 8 lanes of a quarter-warp over interleaved pixel quads — the 8 lanes read 128 contiguous bytes of a row, one
    // conflict-free wavefront — then joined by three shuffle stages and STORED by one lane: every output has exactly one
    // owner, so the shared-memory atomics (CAS loops) of the one-output-per-thread form are gone too.
    if (tid < 192) {
      const int b = tid >> 3, s = tid & 7;
      const float *fa0, *fa1, *fb0, *fb1;
      int o0, o1;
      if (b < 18) {                       // v_w1[j][i] = sum ghpre_j x_i: rows j = 2 jp, 2 jp + 1; columns i = 2 ip, 2 ip + 1
        const int jp = b / 6, ip = b - 6 * jp;
        fa0 = sg + (2 * jp) * kPad; fa1 = fa0 + kPad;
        fb0 = sx + (2 * ip) * kPad; fb1 = fb0 + kPad;
        o0 = (2 * jp) * 12 + 2 * ip; o1 = o0 + 12;
      } else if (b < 21) {                // v_w2[c][h] = sum gpre_c relu(h)_h: c = 0, 1
        const int hp = b - 18;
        fa0 = sg + 6 * kPad; fa1 = fa0 + kPad;
        fb0 = sg + (9 + 2 * hp) * kPad; fb1 = fb0 + kPad;
        o0 = 72 + 2 * hp; o1 = o0 + 6;
      } else {                            // c = 2 (second row of the block unused)
        const int hp = b - 21;
        fa0 = fa1 = sg + 8 * kPad;
        fb0 = sg + (9 + 2 * hp) * kPad; fb1 = fb0 + kPad;
        o0 = 84 + 2 * hp; o1 = -1;
      }
      float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
#pragma unroll 2
      for (int it = 0; it < kBlendThreads / 32; ++it) {
        const int p = (it * 8 + s) * 4;
        const float4 x0 = *reinterpret_cast<const float4*>(fa0 + p), x1 = *reinterpret_cast<const float4*>(fa1 + p);
        const float4 y0 = *reinterpret_cast<const float4*>(fb0 + p), y1 = *reinterpret_cast<const float4*>(fb1 + p);
        a00 += x0.x * y0.x + x0.y * y0.y + x0.z * y0.z + x0.w * y0.w;
        a01 += x0.x * y1.x + x0.y * y1.y + x0.z * y1.z + x0.w * y1.w;
        a10 += x1.x * y0.x + x1.y * y0.y + x1.z * y0.z + x1.w * y0.w;
        a11 += x1.x * y1.x + x1.y * y1.y + x1.z * y1.z + x1.w * y1.w;
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        a00 += __shfl_xor_sync(0xffffffffu, a00, o);
        a01 += __shfl_xor_sync(0xffffffffu, a01, o);
        a10 += __shfl_xor_sync(0xffffffffu, a10, o);
        a11 += __shfl_xor_sync(0xffffffffu, a11, o);
      }
      if (s == 0) {
        swg[o0] = a00; swg[o0 + 1] = a01;
        if (o1 >= 0) { swg[o1] = a10; swg[o1 + 1] = a11; }
      }
    }
#else
    if (tid < 180) {
      const int o = tid % 90, p0 = (tid / 90) * (kBlendThreads / 2);
      const float* fa = o < 72 ? sg + (o / 12) * kPad : sg + (6 + (o - 72) / 6) * kPad;
      const float* fb = o < 72 ? sx + (o % 12) * kPad : sg + (9 + (o - 72) % 6) * kPad;
      float acc = 0.f;
#pragma unroll 4
      for (int p = p0; p < p0 + kBlendThreads / 2; p += 4) {
        const float4 x4 = *reinterpret_cast<const float4*>(fa + p), y4 = *reinterpret_cast<const float4*>(fb + p);
        acc += x4.x * y4.x + x4.y * y4.y + x4.z * y4.z + x4.w * y4.w;
      }
      if (acc != 0.f) atomicAdd(&swg[o], acc);
    }
#endif
  }
}

template <int D, bool DEC>
__global__ void __launch_bounds__(kBlendThreads, MOBGS_BWD_MIN_CTAS) blend_bwd_kernel(const __grid_constant__ MobgsBlendBwd a, int tiles_x, int tiles_y) {
  constexpr bool POSE = false;             // in-kernel rays exist in the transposing kernel only
  __shared__ __align__(128) float4 srec[kBlendThreads][4];
  __shared__ __align__(8) uint64_t sbar;
  __shared__ __align__(16) float sacc[kBlendThreads][kRecFloats];
  __shared__ int sid[kBlendThreads];
  __shared__ unsigned smask[kBlendThreads];
  __shared__ unsigned char swl[kUnits][kBlendThreads];
  __shared__ int warp_max[kBlendThreads / 32];
  __shared__ __align__(16) float sdec[DEC ? 96 : 4];
  __shared__ float swg[DEC ? 96 : 4];      // CTA-level decoder weight-gradient accumulator
  const int tiles = tiles_x * tiles_y;
  const int k = blockIdx.x / tiles, tile = blockIdx.x - k * tiles;
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int tid = threadIdx.x, lane = lane_id();
  const unsigned char* ulist = swl[tid / kUL];          // this lane's unit's list
  unsigned char* wl0 = swl[(tid >> 5) * kUPW];          // list of the warp's first unit
  uint32_t bar_phase = 0;
  if (MOBGS_TMA_STAGE && tid == 0) {
    mbar_init(&sbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // visible after the barrier that follows
  }
  int lx, ly;
  unit_pixel(tid, lx, ly);
  const int ix = tx * kTile + lx, iy = ty * kTile + ly;
  const bool inside = ix < a.width && iy < a.height;
  const float px = ix + 0.5f, py = iy + 0.5f;
  const int beg = min(a.tile_offsets[tile_segment(a.lists, k, tiles, tile)], (int)min(a.list_capacity, (int64_t)0x7fffffff));
  const float4* recs = reinterpret_cast<const float4*>(a.records) + (size_t)a.lists.rec_k[k] * a.N * 4;
  float* v_recs = a.v_records + (size_t)a.lists.rec_k[k] * a.N * kRecFloats;

  float T_final, v_a, bg_dot = 0.f;
  float v_c[D];
  int last;
  bwd_pixel_prologue<D, DEC>(a, k, tid, inside, ix, iy, reinterpret_cast<float*>(&srec[0][0]), &sacc[0][0], sdec, swg,
                             T_final, last, v_c, v_a);
  if (inside) {
    if (a.backgrounds) {
      const float* bg = a.backgrounds + (size_t)k * D;
#pragma unroll
      for (int c = 0; c < D; ++c) bg_dot += bg[c] * v_c[c];
    }
  }
  float T = T_final;
  // S = sum_c v_c[c] * (colour accumulated behind the current Gaussian): the only thing the VJP needs
  // from gsplat's per-channel `buffer`, kept as one scalar.
  float S = 0.f;
  const float tf_term = T_final * (v_a - bg_dot);
  // last list entry any pixel of this tile blended
  int wmax = last;
#pragma unroll
  for (int o = kUL / 2; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  const int umax = wmax;                   // furthest entry any pixel of this lane's unit blended
#pragma unroll
  for (int o = 16; o >= kUL; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  if (lane == 0) warp_max[tid >> 5] = wmax;
  __syncthreads();
  if (DEC && tid < 90 && swg[tid] != 0.f)
    atomicAdd(a.v_w_partial + (size_t)(blockIdx.x % MOBGS_DEC_SLOTS) * 90 + tid, swg[tid]);
  if (POSE && a.v_pose_partial && tid >= 96 && tid < 108 && swg[tid] != 0.f) {
    const int rk = a.dec_rays_per_k == 2 ? a.lists.rec_k[k] : (a.dec_rays_per_k ? k : 0);
    atomicAdd(a.v_pose_partial + ((size_t)rk * MOBGS_POSE_SLOTS + tile % MOBGS_POSE_SLOTS) * 12 + (tid - 96), swg[tid]);
  }
  int tile_last = -1;
#pragma unroll
  for (int w = 0; w < kBlendThreads / 32; ++w) tile_last = max(tile_last, warp_max[w]);
  if (tile_last < beg) return;

  // walk the list back to front in batches: batch b covers [hi - 255, hi]
  for (int hi = tile_last; hi >= beg; hi -= kBlendThreads) {
    const int lo = max(beg, hi - kBlendThreads + 1);
    const int bn = hi - lo + 1;
    __syncthreads();   // previous batch fully flushed
#if MOBGS_TMA_STAGE
    constexpr uint32_t kRecBytes = D > 6 ? 64 : (D > 2 ? 48 : 32);
    if (tid == 0) mbar_expect_tx(&sbar, (uint32_t)bn * kRecBytes);
    if (tid < bn) {
      const int g = list_gid(a.sorted_ids, hi - tid, a.N);   // slot t holds list entry hi - t
      sid[tid] = g;
      bulk_g2s(&srec[tid][0], recs + (size_t)g * 4, kRecBytes, &sbar);
    }
#pragma unroll
    for (int c = 0; c < kRecFloats; ++c) sacc[tid][c] = 0.f;
    mbar_wait(&sbar, bar_phase);
    bar_phase ^= 1;
    if (tid < bn) smask[tid] = unit_mask(srec[tid][0], srec[tid][1], (float)(tx * kTile), (float)(ty * kTile));
#else
    if (tid < bn) {
      const int g = list_gid(a.sorted_ids, hi - tid, a.N);   // slot t holds list entry hi - t
      sid[tid] = g;
      const float4* r = recs + (size_t)g * 4;
      const float4 q0 = __ldg(r), q1 = __ldg(r + 1);
      srec[tid][0] = q0; srec[tid][1] = q1;
      if (D > 2) srec[tid][2] = __ldg(r + 2);
      if (D > 6) srec[tid][3] = __ldg(r + 3);
      smask[tid] = unit_mask(q0, q1, (float)(tx * kTile), (float)(ty * kTile));
    }
#pragma unroll
    for (int c = 0; c < kRecFloats; ++c) sacc[tid][c] = 0.f;
#endif
    __syncthreads();
    // entries above this unit's furthest pixel contribute nothing: they never enter its list
    const int cnt = build_unit_lists(smask, wl0, tid >> 5, lane, max(0, hi - umax), bn);
    int cnt_warp = cnt;                    // the warp runs until its longest unit list is exhausted;
#pragma unroll                             // a unit that is through idles on slot 0 with valid = false
    for (int o = 16; o >= kUL; o >>= 1) cnt_warp = max(cnt_warp, __shfl_xor_sync(0xffffffffu, cnt_warp, o));
    for (int i = 0; i < cnt_warp; ++i) {
      const bool act = i < cnt;
      const int t = act ? ulist[i] : 0;
      const int idx = hi - t;
      const float4 r0 = srec[t][0], r1 = srec[t][1];
      const float dx = r0.x - px, dy = r0.y - py;
      const float sigma = 0.5f * (r0.w * dx * dx + r1.y * dy * dy) + r1.x * dx * dy;
      const float vis = ex2_approx(sigma * kNegLog2e);
      const float alpha = fminf(kAlphaMax, r0.z * vis);
      const bool valid = act && inside && idx <= last && sigma >= 0.f && alpha >= kAlphaMin;
      if (!__any_sync(0xffffffffu, valid)) continue;
      // v_x v_y v_opac v_ca v_cb v_cc v_col[D], zero-padded to a power of two
      constexpr int NV = (6 + D) <= 8 ? 8 : 16;
      float g[NV];
#if MOBGS_BWD_PREDICATED
      // branch-free: lanes whose pixel does not blend this Gaussian run the same arithmetic with
      // alpha = 0, which leaves T and S unchanged and makes every gradient term exactly zero
      {
        float4 r2 = make_float4(0, 0, 0, 0), r3 = make_float4(0, 0, 0, 0);
        if (D > 2) r2 = srec[t][2];
        if (D > 6) r3 = srec[t][3];
        const float al = valid ? alpha : 0.f;
        const float ra = __fdividef(1.f, 1.f - al);
        T *= ra;
        const float fac = al * T;
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          d += rec_color(r1, r2, r3, c) * v_c[c];
          g[6 + c] = fac * v_c[c];
        }
#pragma unroll
        for (int c = 6 + D; c < NV; ++c) g[c] = 0.f;
        const float v_alpha = d * T + (tf_term - S) * ra;
        S += d * fac;
        const float visv = valid ? vis : 0.f;          // (sigma < 0 lanes may hold vis = inf)
        const float ov = r0.z * visv;
        const float gate = ov <= kAlphaMax ? 1.f : 0.f;
        const float v_sigma = -ov * v_alpha * gate;
        const float sx = v_sigma * dx, sy = v_sigma * dy;
        g[0] = r0.w * sx + r1.x * sy;
        g[1] = r1.x * sx + r1.y * sy;
        g[2] = visv * v_alpha * gate;
        g[3] = sx * dx;            // the 1/2 of d sigma / d conic_a, conic_c is applied once, at the flush
        g[4] = sx * dy;
        g[5] = sy * dy;
      }
#else
#pragma unroll
      for (int c = 0; c < NV; ++c) g[c] = 0.f;
      if (valid) {
        float4 r2 = make_float4(0, 0, 0, 0), r3 = make_float4(0, 0, 0, 0);
        if (D > 2) r2 = srec[t][2];
        if (D > 6) r3 = srec[t][3];
        const float ra = __fdividef(1.f, 1.f - alpha);
        T *= ra;
        const float fac = alpha * T;
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          d += rec_color(r1, r2, r3, c) * v_c[c];
          g[6 + c] = fac * v_c[c];
        }
        const float v_alpha = d * T + (tf_term - S) * ra;
        S += d * fac;
        if (r0.z * vis <= kAlphaMax) {
          const float v_sigma = -r0.z * vis * v_alpha;
          const float sx = v_sigma * dx, sy = v_sigma * dy;
          g[0] = r0.w * sx + r1.x * sy;
          g[1] = r1.x * sx + r1.y * sy;
          g[2] = vis * v_alpha;
          g[3] = sx * dx;
          g[4] = sx * dy;
          g[5] = sy * dy;
        }
      }
#endif
#ifdef MOBGS_ABL_NAIVE_REDUCE
      // gsplat-style: one full 5-step shuffle reduction per value, lane 0 adds them one by one
#pragma unroll
      for (int c = 0; c < 6 + D; ++c) {
#pragma unroll
        for (int o = kUL / 2; o > 0; o >>= 1) g[c] += __shfl_xor_sync(0xffffffffu, g[c], o);
      }
      if ((lane & (kUL - 1)) == 0) {
#pragma unroll
        for (int c = 0; c < 6 + D; ++c) atomicAdd(&sacc[t][c], g[c]);
      }
#else
      const int vidx = butterfly_reduce<NV>(g, lane);
      constexpr int kOwnerMask = NV < kUL ? kUL / NV - 1 : 0;   // lanes whose low bits are 0 own values
      constexpr int kLeft = NV > kUL ? NV / kUL : 1;            // values left per lane
      if ((lane & kOwnerMask) == 0) {
#pragma unroll
        for (int r = 0; r < kLeft; ++r)
          if (g[r] != 0.f) atomicAdd(&sacc[t][vidx + r], g[r]);
      }
#endif
    }
    __syncthreads();
    if (tid < bn) {
      const float4* s4 = reinterpret_cast<const float4*>(sacc[tid]);
      float* dst = v_recs + (size_t)sid[tid] * kRecFloats;
      constexpr int kVec = (6 + D + 3) / 4;
#pragma unroll
      for (int v = 0; v < kVec; ++v) {
        float4 s = s4[v];
        if (v == 0) s.w *= 0.5f;           // conic_a   (record layout: x y opac ca | cb cc ...)
        if (v == 1) s.y *= 0.5f;           // conic_c
        if (s.x != 0.f || s.y != 0.f || s.z != 0.f || s.w != 0.f) red_add_v4(dst + 4 * v, s.x, s.y, s.z, s.w);
        if (v == 0 && a.v_means2d_sep && k == a.sep_list && (s.x != 0.f || s.y != 0.f)) {
          atomicAdd(a.v_means2d_sep + 2 * (size_t)sid[tid], s.x);
          atomicAdd(a.v_means2d_sep + 2 * (size_t)sid[tid] + 1, s.y);
        }
      }
    }
  }
}


#if MOBGS_UNIT_LANES == 16
// ---------------------------------------------------------------------------------------------
// Backward, transposing variant (kUL == 16).  The butterfly above spends ~45 % of the kernel's
// instructions moving per-(pixel, Gaussian) terms between lanes.  Here the reduction over pixels
// is done by changing ownership instead: a unit walks its list in blocks of kBlk = 8 entries.
//   Phase A (lane = pixel): the usual back-to-front recurrence (T, S) per entry, but only the two
//     scalars every gradient of the pair is linear in — fac = alpha T (colours) and v_sigma
//     (geometry, opacity) — are produced, and parked in a per-warp [8 entries][32 pixels] matrix.
//   Phase B (lane = (entry j, half of the unit's pixels)): reads row j of both matrices and sums
//     over its 8 pixels the colour gradients fac * v_c[p] (v_c of the unit's pixels sits in shared
//     memory, written once per tile) and the moments sum v_sigma {dx, dy, 1, dx^2, dx dy, dy^2};
//     one butterfly stage joins the two halves, the 16 sums go to the tile accumulator.
// The flush turns moments into gradients (v_x = a Ax + b Ay, v_y = b Ax + c Ay, v_opac = -A0 / opac,
// conic = (Axx / 2, Axy, Ayy / 2)) — linear, so summing moments over units first is equivalent.
// 1: Phase B lanes turn their (unit, Gaussian) sums into gradients and send them straight to global
// memory (2 x RED.128 per lane); 0: they are first combined across the tile's units in a shared-memory
// accumulator (16 CAS-loop atomics per (unit, Gaussian) — those were 15 % of the instructions and 28 % of
// the stall samples) and flushed once per (tile, Gaussian).
#ifndef MOBGS_BWD_DIRECT_RED
#define MOBGS_BWD_DIRECT_RED 1
#endif
// 1: Phase A without the per-iteration "is this unit's list exhausted" test: the unit lists are pre-filled with the index
// of a dummy record (opacity 0 -> alpha = 0 -> the pair is invalid by the ordinary alpha test), the "blended by this pixel"
// test is one compare against a per-batch constant, and v_sigma is selected straight from opac * vis (6 of 54
// instructions per iteration).
#ifndef MOBGS_BWD_LEAN_A
#define MOBGS_BWD_LEAN_A 1
#endif
// 1: Phase B sums the six moments of v_sigma about the unit's first pixel — sums of vs, c vs, r vs, c^2 vs, c r vs with
// the compile-time pixel offsets c in 0..3, r in 0..1 (27 instructions per lane and block) — and shifts them to the
// Gaussian's mean once per (unit, Gaussian); 0: accumulates them about the mean pixel by pixel (64 instructions).
#ifndef MOBGS_BWD_UNIT_MOMENTS
#define MOBGS_BWD_UNIT_MOMENTS 1
#endif
// 1: L2 prefetch of the list tail's records before the prologue (see blend_bwd_tr_kernel).  Measured 0.5 % SLOWER
// (2.972 vs 2.957 ms): the staging latency is already covered by the other three CTAs of the SM; kept as a switch.
#ifndef MOBGS_BWD_PREFETCH
#define MOBGS_BWD_PREFETCH 0
#endif
constexpr int kBwdBatch = MOBGS_BWD_DIRECT_RED ? 192 : 128;   // list entries staged per batch (shared-memory budget)
constexpr int kDummySlot = kBwdBatch;          // staged row of the dummy record (MOBGS_BWD_LEAN_A)
constexpr int kRecRow = 20;                    // floats between staged records (16 used): the rows of 8 different
                                               // entries start 4 banks apart for the per-entry reads of Phase B
constexpr int kListRow = kBwdBatch + 8;        // bytes between unit lists: 8-byte loads of a warp's two units differ in bank
constexpr int kBlk = 8;                        // entries per Phase A / Phase B block
constexpr int kFRow = 36;                      // row stride of the per-warp F / VS matrices (32 pixels + 4: the
                                               // 8 rows a quarter-warp reads in Phase B start 4 banks apart)
constexpr int kVPix = 12;                      // floats per pixel in sV (10 used; 3 x LDS.128)
constexpr int kVHalf = 8 * kVPix + 4;          // second half of a unit's pixels: +4 banks
constexpr int kVUnit = 2 * kVHalf + 8;         // second unit of the warp: +16 banks -> the four (unit, half)
constexpr int kVWarp = 2 * kVUnit;             //   groups of a warp read four different 16-byte bank groups
constexpr int kAccRow = 17;                    // accumulator row stride: row t starts at bank 17 t
constexpr int kAccFloats = MOBGS_BWD_DIRECT_RED ? 0 : kBwdBatch * kAccRow;
constexpr int kTrScratch = (kBwdBatch + 1) * kRecRow + kAccFloats + 8 * 2 * kBlk * kFRow;   // floats
static_assert(kTrScratch >= 27 * kProPad + 96, "prologue scratch (+ the pose partials) must fit the aliased buffers");
constexpr size_t kTrSmemBytes = (size_t)(kTrScratch + 8 * kVWarp + 112 + 112) * 4 + kBwdBatch * 8 + 16 * kListRow + 8 * 4 + 16;

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// FLOW: two more colour channels per entry, (records[flow_ref][g].xy - own xy), parked in floats 16..17 of the
// staged row (rows are 20 floats, the TMA copy fills 16); their pixel gradients g_flow sit in slots 10..11 of sV.
template <int D, bool DEC, bool FLOW = false, bool POSE = false>
__global__ void __launch_bounds__(kBlendThreads, MOBGS_BWD_MIN_CTAS) blend_bwd_tr_kernel(const __grid_constant__ MobgsBlendBwd a, int tiles_x, int tiles_y) {
  static_assert(kUL == 16, "transposing backward is written for 4x4-pixel units");
  static_assert(!FLOW || (D == 10 && MOBGS_BWD_DIRECT_RED), "fused flow channels need the D = 10 layout and direct reductions");
  extern __shared__ __align__(128) float smem[];
  float* srec = smem;                                            // [kBwdBatch + 1][kRecRow]   (TMA destination + dummy row)
  float* sacc = srec + (kBwdBatch + 1) * kRecRow;                // [kBwdBatch][kAccRow]
  float* sF = sacc + kAccFloats;                                 // [8 warps][F | VS][kBlk][kFRow]
  float* sV = smem + kTrScratch;                                 // [8 warps][kVWarp]
  float* sdec = sV + 8 * kVWarp;                                 // [112] decoder weights | camera pose
  float* swg = sdec + 112;                                       // [112] their gradient accumulators
  int* sid = reinterpret_cast<int*>(swg + 112);                  // [kBwdBatch]
  unsigned* smask = reinterpret_cast<unsigned*>(sid + kBwdBatch);   // [kBwdBatch]
  int* warp_max = reinterpret_cast<int*>(smask + kBwdBatch);     // [8]
  uint64_t* sbar = reinterpret_cast<uint64_t*>(warp_max + 8);    // 8-byte aligned: everything before is a multiple of 8 B
  unsigned char* swl = reinterpret_cast<unsigned char*>(sbar + 1);   // [16 units][kListRow], 8-byte aligned rows

  const int tiles = tiles_x * tiles_y;
  const int k = blockIdx.x / tiles, tile = blockIdx.x - k * tiles;
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  const unsigned char* ulist = swl + (tid / kUL) * kListRow;     // this lane's unit's list
  unsigned char* wl0 = swl + warp * kUPW * kListRow;             // list of the warp's first unit
  uint32_t bar_phase = 0;
  if (tid == 0) {
    mbar_init(sbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // visible after the barrier that follows
  }
  int lx, ly;
  unit_pixel(tid, lx, ly);
  const int ix = tx * kTile + lx, iy = ty * kTile + ly;
  const bool inside = ix < a.width && iy < a.height;
  const float px = ix + 0.5f, py = iy + 0.5f;
  const int beg = min(a.tile_offsets[tile_segment(a.lists, k, tiles, tile)], (int)min(a.list_capacity, (int64_t)0x7fffffff));
  const float4* recs = reinterpret_cast<const float4*>(a.records) + (size_t)a.lists.rec_k[k] * a.N * 4;
  float* v_recs = a.v_records + (size_t)a.lists.rec_k[k] * a.N * kRecFloats;
  const float4* recs_ref = FLOW ? reinterpret_cast<const float4*>(a.records) + (size_t)a.flow_ref * a.N * 4 : nullptr;
  float* v_recs_ref = FLOW ? a.v_records + (size_t)a.flow_ref * a.N * kRecFloats : nullptr;
#if MOBGS_BWD_PREFETCH
  // The records of the first batch are staged only after the per-pixel prologue (they alias its scratch).  Pull the
  // list tail's records towards L2 now, so that the staging copies hit L2 instead of waiting on DRAM: the tail is
  // where the walk starts whenever some pixel of the tile blended the whole list (the usual case).
  {
    const int endl = min(a.tile_offsets[tile_segment(a.lists, k, tiles, tile) + 1], (int)min(a.list_capacity, (int64_t)0x7fffffff));
    const int pf = endl - 1 - tid;
    if (tid < kBwdBatch && pf >= beg) {
      const int g = list_gid(a.sorted_ids, pf, a.N);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(recs + (size_t)g * 4));
    }
  }
#endif

  if (DEC && a.sparse_grads) {
    // MobgsBlendBwd.sparse_grads: most tiles of this launch receive no gradient.  Test the upstream gradients of the
    // tile's pixels before anything else — a tile without any leaves here, before the decoder prologue (img10 reads, ray
    // generation, decoder recompute, 27 scratch rows, the weight-gradient dot products) that the exit below comes after.
    bool nz = false;
    if (inside) {
      const size_t P = (size_t)a.width * a.height, pp = (size_t)iy * a.width + ix, p = (size_t)k * P + pp;
      const int meanK = a.mean_K > 0 ? a.mean_K : a.K;
      if (a.g_mean && k < meanK) nz = __ldg(a.g_mean + pp) != 0.f || __ldg(a.g_mean + P + pp) != 0.f || __ldg(a.g_mean + 2 * P + pp) != 0.f;
      if (a.g_rgb) {
        const float* gr = a.g_rgb + (size_t)k * 3 * P + pp;
        nz = nz || __ldg(gr) != 0.f || __ldg(gr + P) != 0.f || __ldg(gr + 2 * P) != 0.f;
      }
      if (a.g_depth) nz = nz || __ldg(a.g_depth + p) != 0.f;
      if (a.g_alpha) nz = nz || __ldg(a.g_alpha + p) != 0.f;
      if (FLOW) {
        const float2 gf = __ldg(reinterpret_cast<const float2*>(a.g_flow + p * 2));
        nz = nz || gf.x != 0.f || gf.y != 0.f;
      }
    }
    if (!__syncthreads_or(nz)) return;
  }
  float T_final, v_a, bg_dot = 0.f;
  float v_c[D];
  int last;
  bwd_pixel_prologue<D, DEC, POSE>(a, k, tid, inside, ix, iy, smem, smem + 12 * kProPad, sdec, swg,
                             T_final, last, v_c, v_a, smem + 27 * kProPad);
  float v_fl0 = 0.f, v_fl1 = 0.f;
  if (FLOW && inside) {
    const float2 gf = __ldg(reinterpret_cast<const float2*>(a.g_flow + (((size_t)k * a.height + iy) * a.width + ix) * 2));
    v_fl0 = gf.x; v_fl1 = gf.y;
  }
  // a tile none of whose pixels receives a gradient contributes exactly nothing (every term of the VJP is linear
  // in v_c / v_a): it leaves after the prologue.  This is what makes a second backward pass through the same
  // render cheap when the second loss touches few of its images (train.py:629 then :680).
  bool px_zero = v_a == 0.f && v_fl0 == 0.f && v_fl1 == 0.f;
#pragma unroll
  for (int c = 0; c < D; ++c) px_zero = px_zero && v_c[c] == 0.f;
  if (inside && a.backgrounds) {
    const float* bg = a.backgrounds + (size_t)k * D;
#pragma unroll
    for (int c = 0; c < D; ++c) bg_dot += bg[c] * v_c[c];
  }
  // this pixel's v_c -> sV (read by the Phase B lanes of its unit): 12 floats per pixel, zero padded
  const int q = lane & 15, hu = lane >> 4;                        // pixel in unit, unit in warp
  {
    float* vp = sV + warp * kVWarp + hu * kVUnit + (q >> 3) * kVHalf + (q & 7) * kVPix;
    float t12[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) t12[c] = c < D ? v_c[c % D] : 0.f;
    if (FLOW) { t12[10] = v_fl0; t12[11] = v_fl1; }
#pragma unroll
    for (int c = 0; c < 12; c += 4)
      if (c < D) *reinterpret_cast<float4*>(vp + c) = make_float4(t12[c], t12[c + 1], t12[c + 2], t12[c + 3]);
  }
  float T = T_final;
  float S = 0.f;                           // sum_c v_c[c] * (colour accumulated behind the current Gaussian)
  const float tf_term = T_final * (v_a - bg_dot);
  int wmax = last;
#pragma unroll
  for (int o = kUL / 2; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  const int umax = wmax;                   // furthest entry any pixel of this lane's unit blended
  wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, 16));
  if (lane == 0) warp_max[warp] = wmax;
  fence_proxy_async();                     // the prologue's generic-proxy scratch writes precede the TMA writes to srec
  const int tile_zero = __syncthreads_and(px_zero);   // prologue scratch is free, sV / swg / warp_max are complete
  if (DEC && tid < 90 && swg[tid] != 0.f)
    atomicAdd(a.v_w_partial + (size_t)(blockIdx.x % MOBGS_DEC_SLOTS) * 90 + tid, swg[tid]);
  if (POSE && a.v_pose_partial && tid >= 96 && tid < 108 && swg[tid] != 0.f) {
    const int rk = a.dec_rays_per_k == 2 ? a.lists.rec_k[k] : (a.dec_rays_per_k ? k : 0);
    atomicAdd(a.v_pose_partial + ((size_t)rk * MOBGS_POSE_SLOTS + tile % MOBGS_POSE_SLOTS) * 12 + (tid - 96), swg[tid]);
  }
  if (tile_zero) return;
  int tile_last = -1;
#pragma unroll
  for (int w = 0; w < kBlendThreads / 32; ++w) tile_last = max(tile_last, warp_max[w]);
  if (tile_last < beg) return;
#if MOBGS_BWD_LEAN_A
  if (tid < kRecRow) srec[kDummySlot * kRecRow + tid] = 0.f;      // dummy record: opacity 0 (visible after the batch barriers)
#endif

  float* Fm = sF + warp * (2 * kBlk * kFRow);                     // F  [kBlk][kFRow]: fac     of (entry, pixel of the warp)
  float* VSm = Fm + kBlk * kFRow;                                 // VS [kBlk][kFRow]: v_sigma
  const int pj = q & 7, ph = q >> 3;                              // Phase B: entry in block, pixel half
  const float* vrow = sV + warp * kVWarp + hu * kVUnit + ph * kVHalf;
  // tile-local x / y of the Phase B lane's pixels: columns 0..3 of the unit, rows 2 ph, 2 ph + 1
  const float ubx = (float)(tx * kTile + ((tid / kUL) % kUX) * kUW) + 0.5f;
  const float uby = (float)(ty * kTile + ((tid / kUL) / kUX) * kUH + 2 * ph) + 0.5f;
  constexpr uint32_t kRecBytes = D > 6 ? 64 : (D > 2 ? 48 : 32);

  // walk the list back to front in batches: batch b covers [hi - kBwdBatch + 1, hi]
  for (int hi = tile_last; hi >= beg; hi -= kBwdBatch) {
    const int lo = max(beg, hi - kBwdBatch + 1);
    const int bn = hi - lo + 1;
    __syncthreads();   // previous batch fully flushed
    if (tid == 0) mbar_expect_tx(sbar, (uint32_t)bn * kRecBytes);
    float2 mref = make_float2(0.f, 0.f);
    const bool have_masks = kUnits <= 16 && a.list_masks != nullptr;     // the forward's unit masks of these lists
    unsigned fwd_mask = 0u;
    if (tid < bn) {
      const int g = list_gid(a.sorted_ids, hi - tid, a.N);   // slot t holds list entry hi - t
      sid[tid] = g;
      bulk_g2s(srec + tid * kRecRow, recs + (size_t)g * 4, kRecBytes, sbar);
      if (FLOW) mref = __ldg(reinterpret_cast<const float2*>(recs_ref + (size_t)g * 4));
      if (have_masks) fwd_mask = a.list_masks[hi - tid];
    }
    for (int i = tid; i < kAccFloats; i += kBlendThreads) sacc[i] = 0.f;
#if MOBGS_BWD_LEAN_A
    // every list byte names the dummy record until build_unit_lists overwrites it
    for (int i = tid; i < 16 * kListRow / 4; i += kBlendThreads) reinterpret_cast<uint32_t*>(swl)[i] = 0x01010101u * (uint32_t)kDummySlot;
    const int tmin = hi - last;            // pixel blended list entry hi - t  <=>  t >= tmin  (last = -1 outside the image)
#endif
    mbar_wait(sbar, bar_phase);
    bar_phase ^= 1;
    if (tid < bn) {
      const float4* r = reinterpret_cast<const float4*>(srec + tid * kRecRow);
      smask[tid] = have_masks ? fwd_mask : unit_mask(r[0], r[1], (float)(tx * kTile), (float)(ty * kTile));
      if (FLOW) *reinterpret_cast<float2*>(srec + tid * kRecRow + 16) = make_float2(mref.x - r[0].x, mref.y - r[0].y);
      rescale_conic(srec + tid * kRecRow);
    }
    __syncthreads();
    // entries above this unit's furthest pixel contribute nothing: they never enter its list
    const int cnt = build_unit_lists<kListRow>(smask, wl0, warp, lane, max(0, hi - umax), bn);
    const int cnt_warp = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, 16));
    for (int base = 0; base < cnt_warp; base += kBlk) {
      const int nb = min(kBlk, cnt_warp - base);
      const uint2 tl = *reinterpret_cast<const uint2*>(ulist + base);   // the block's 8 list bytes (stale past cnt)
      // ---- Phase A: lane = pixel.  A unit that is through idles on slot 0 with valid = false.
#pragma unroll
      for (int i = 0; i < kBlk; ++i) {
        if (i >= nb) break;
#if MOBGS_BWD_LEAN_A
        const int t = (int)(((i < 4 ? tl.x : tl.y) >> (8 * (i & 3))) & 0xffu);   // past the unit's count: the dummy record
#else
        const bool act = base + i < cnt;
        const int t = act ? (int)(((i < 4 ? tl.x : tl.y) >> (8 * (i & 3))) & 0xffu) : 0;
#endif
        const float4* r = reinterpret_cast<const float4*>(srec + t * kRecRow);
        const float4 r0 = r[0], r1 = r[1];
        float4 r2 = make_float4(0, 0, 0, 0), r3 = make_float4(0, 0, 0, 0);
        if (D > 2) r2 = r[2];
        if (D > 6) r3 = r[3];
        const float dx = r0.x - px, dy = r0.y - py;
        const float pw = pair_exponent(r0, r1, dx, dy);          // = -sigma log2(e)
        const float vis = ex2_approx(pw);
        const float ou = r0.z * vis;
        const float alpha = fminf(kAlphaMax, ou);
#if MOBGS_BWD_LEAN_A
        const bool valid = t >= tmin && pw <= 0.f && alpha >= kAlphaMin;
#else
        const bool valid = act && inside && hi - t <= last && pw <= 0.f && alpha >= kAlphaMin;
#endif
        // lanes whose pixel does not blend this Gaussian run the same arithmetic with alpha = 0, which
        // leaves T and S unchanged and makes both stored terms exactly zero
        const float al = valid ? alpha : 0.f;
        const float ra = rcp_approx(1.f - al);
        T *= ra;
        const float fac = al * T;
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) d += rec_color(r1, r2, r3, c) * v_c[c];
        if (FLOW) {
          const float2 f = *reinterpret_cast<const float2*>(srec + t * kRecRow + 16);
          d += f.x * v_fl0 + f.y * v_fl1;
        }
        const float v_alpha = d * T + (tf_term - S) * ra;
        S += d * fac;
#if MOBGS_BWD_LEAN_A
        const float v_sigma = (valid && ou <= kAlphaMax) ? -ou * v_alpha : 0.f;   // (invalid lanes may hold ou = inf: selected away)
#else
        const float ov = r0.z * (valid ? vis : 0.f);             // (sigma < 0 lanes may hold vis = inf)
        const float v_sigma = ov <= kAlphaMax ? -ov * v_alpha : 0.f;
#endif
        Fm[i * kFRow + lane] = fac;
        VSm[i * kFRow + lane] = v_sigma;
      }
      __syncwarp();
      // ---- Phase B: lane = (entry pj of the block, pixel half ph) of its unit
      float acc[16];
      float accf0 = 0.f, accf1 = 0.f;         // FLOW: the two flow-colour gradient sums
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = 0.f;
      const bool on = pj < nb && base + pj < cnt;
      int tb = 0;
      if (on) {
        tb = (int)(((pj < 4 ? tl.x : tl.y) >> (8 * (pj & 3))) & 0xffu);
        const float2 mean = *reinterpret_cast<const float2*>(srec + tb * kRecRow);
        const float* fr = Fm + pj * kFRow + hu * 16 + ph * 8;
        const float* sr = VSm + pj * kFRow + hu * 16 + ph * 8;
        const float4 f0 = *reinterpret_cast<const float4*>(fr), f1 = *reinterpret_cast<const float4*>(fr + 4);
        const float4 s0 = *reinterpret_cast<const float4*>(sr), s1 = *reinterpret_cast<const float4*>(sr + 4);
        const float fv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
        const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#if MOBGS_BWD_UNIT_MOMENTS
        {
          // this lane's pixels are (c, r), c = 0..3, r = 0..1, at distance (X - c, Y - r) from the mean
          const float X = mean.x - ubx, Y = mean.y - uby;
          const float R0 = (sv[0] + sv[1]) + (sv[2] + sv[3]), R1 = (sv[4] + sv[5]) + (sv[6] + sv[7]);   // sum vs per row
          const float C0 = fmaf(3.f, sv[3], fmaf(2.f, sv[2], sv[1])), C1 = fmaf(3.f, sv[7], fmaf(2.f, sv[6], sv[5]));   // sum c vs
          const float Q0 = fmaf(9.f, sv[3], fmaf(4.f, sv[2], sv[1])), Q1 = fmaf(9.f, sv[7], fmaf(4.f, sv[6], sv[5]));   // sum c^2 vs
          const float S0 = R0 + R1, Sc = C0 + C1, Scc = Q0 + Q1;      // (sum r vs = sum r^2 vs = R1, sum c r vs = C1)
          const float m = X * S0, n = Y * S0;
          acc[0] = m - Sc;                                  // sum vs dx
          acc[1] = n - R1;                                  // sum vs dy
          acc[2] = S0;
          acc[3] = fmaf(X, fmaf(-2.f, Sc, m), Scc);         // sum vs dx^2  = X (X S0 - 2 Sc) + Scc
          acc[4] = fmaf(X, acc[1], fmaf(-Y, Sc, C1));       // sum vs dx dy = X (Y S0 - Sr) - Y Sc + Scr
          acc[5] = fmaf(Y, fmaf(-2.f, R1, n), R1);          // sum vs dy^2  = Y (Y S0 - 2 Sr) + Srr
        }
#else
        float dxs[4], dys[2];
#pragma unroll
        for (int c = 0; c < 4; ++c) dxs[c] = mean.x - (ubx + (float)c);
        dys[0] = mean.y - uby;
        dys[1] = mean.y - (uby + 1.f);
#endif
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
#if !MOBGS_BWD_UNIT_MOMENTS
          const float dx = dxs[pp & 3], dy = dys[pp >> 2];
          const float vs = sv[pp], sx = vs * dx, sy = vs * dy;
          acc[0] += sx;
          acc[1] += sy;
          acc[2] += vs;
          acc[3] += sx * dx;
          acc[4] += sx * dy;
          acc[5] += sy * dy;
#endif
          const float4* vp = reinterpret_cast<const float4*>(vrow + pp * kVPix);
          const float4 v0 = vp[0];
          acc[6] += fv[pp] * v0.x;
          if (D > 1) acc[7] += fv[pp] * v0.y;
          if (D > 2) acc[8] += fv[pp] * v0.z;
          if (D > 3) acc[9] += fv[pp] * v0.w;
          if (D > 4) {
            const float4 v1 = vp[1];
            acc[10] += fv[pp] * v1.x;
            if (D > 5) acc[11] += fv[pp] * v1.y;
            if (D > 6) acc[12] += fv[pp] * v1.z;
            if (D > 7) acc[13] += fv[pp] * v1.w;
          }
          if (D > 8) {
            const float4 v2 = vp[2];
            acc[14] += fv[pp] * v2.x;
            if (D > 9) acc[15] += fv[pp] * v2.y;
            if (FLOW) { accf0 += fv[pp] * v2.z; accf1 += fv[pp] * v2.w; }
          }
        }
      }
      // join the two pixel halves: afterwards acc[i] of lane (pj, ph) is the unit sum of value 8 ph + i
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float send = ph ? acc[i] : acc[i + 8];
        const float keep = ph ? acc[i + 8] : acc[i];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      if (FLOW) {
        accf0 += __shfl_xor_sync(0xffffffffu, accf0, 8);
        accf1 += __shfl_xor_sync(0xffffffffu, accf1, 8);
      }
#if MOBGS_BWD_DIRECT_RED
      if (on) {
        float* dst = v_recs + (size_t)sid[tb] * kRecFloats + ph * 8;
        if (ph == 0) {
          // moments -> gradients: record layout x y opac ca | cb cc c0 c1
          const float4* r = reinterpret_cast<const float4*>(srec + tb * kRecRow);
          const float4 r0 = r[0];
          const float2 r1 = *reinterpret_cast<const float2*>(r + 1);
          const float ax = acc[0], ay = acc[1];
          const float ca = r0.w * (1.f / kConicDiag), cb = r1.x * (1.f / kConicOff), cc = r1.y * (1.f / kConicDiag);
          acc[0] = ca * ax + cb * ay;
          acc[1] = cb * ax + cc * ay;
          acc[2] = acc[2] != 0.f ? -acc[2] * __fdividef(1.f, r0.z) : 0.f;
          acc[3] *= 0.5f;
          acc[5] *= 0.5f;
          if (a.v_means2d_sep && k == a.sep_list && (acc[0] != 0.f || acc[1] != 0.f)) {
            atomicAdd(a.v_means2d_sep + 2 * (size_t)sid[tb], acc[0]);
            atomicAdd(a.v_means2d_sep + 2 * (size_t)sid[tb] + 1, acc[1]);
          }
          if (FLOW) {
            // flow colour = ref.xy - own.xy: its gradient goes to the reference record's mean and, negated, to this one's
            acc[0] -= accf0;
            acc[1] -= accf1;
            if (accf0 != 0.f || accf1 != 0.f) red_add_v2(v_recs_ref + (size_t)sid[tb] * kRecFloats, accf0, accf1);
          }
        }
        constexpr int kVec = (6 + D + 3) / 4;     // 16-byte chunks of the gradient record in use
        if (2 * ph < kVec && (acc[0] != 0.f || acc[1] != 0.f || acc[2] != 0.f || acc[3] != 0.f))
          red_add_v4(dst, acc[0], acc[1], acc[2], acc[3]);
        if (2 * ph + 1 < kVec && (acc[4] != 0.f || acc[5] != 0.f || acc[6] != 0.f || acc[7] != 0.f))
          red_add_v4(dst + 4, acc[4], acc[5], acc[6], acc[7]);
      }
      __syncwarp();   // Phase B reads of F / VS are done before the next block overwrites them
    }
#else
      if (on) {
        float* dst = sacc + tb * kAccRow + ph * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (acc[i] != 0.f) atomicAdd(dst + i, acc[i]);
      }
      __syncwarp();   // Phase B reads of F / VS are done before the next block overwrites them
    }
    __syncthreads();
    if (tid < bn) {
      const float4* r = reinterpret_cast<const float4*>(srec + tid * kRecRow);
      const float4 r0 = r[0], r1 = r[1];
      const float* sa = sacc + tid * kAccRow;
      const float ax = sa[0], ay = sa[1], a0 = sa[2];
      float4 s[4];
      const float ca = r0.w * (1.f / kConicDiag), cb = r1.x * (1.f / kConicOff), cc = r1.y * (1.f / kConicDiag);
      s[0] = make_float4(ca * ax + cb * ay, cb * ax + cc * ay, a0 != 0.f ? -a0 / r0.z : 0.f, 0.5f * sa[3]);
      s[1] = make_float4(sa[4], 0.5f * sa[5], sa[6], sa[7]);
      s[2] = make_float4(sa[8], sa[9], sa[10], sa[11]);
      s[3] = make_float4(sa[12], sa[13], sa[14], sa[15]);
      float* dst = v_recs + (size_t)sid[tid] * kRecFloats;
      constexpr int kVec = (6 + D + 3) / 4;
#pragma unroll
      for (int v = 0; v < kVec; ++v)
        if (s[v].x != 0.f || s[v].y != 0.f || s[v].z != 0.f || s[v].w != 0.f)
          red_add_v4(dst + 4 * v, s[v].x, s[v].y, s[v].z, s[v].w);
      if (a.v_means2d_sep && k == a.sep_list && (s[0].x != 0.f || s[0].y != 0.f)) {
        atomicAdd(a.v_means2d_sep + 2 * (size_t)sid[tid], s[0].x);
        atomicAdd(a.v_means2d_sep + 2 * (size_t)sid[tid] + 1, s[0].y);
      }
    }
#endif
  }
}

#endif  // MOBGS_UNIT_LANES == 16

}  // namespace mobgs

using namespace mobgs;

template <int D>
static void launch_fwd(const MobgsBlendFwd& a, int tiles_x, int tiles_y, cudaStream_t s) {
  blend_fwd_kernel<D, false><<<a.K * tiles_x * tiles_y, kBlendThreads, 0, s>>>(a, tiles_x, tiles_y);
}
// 1: transposing backward (blend_bwd_tr_kernel, needs 4x4 units); 0: butterfly backward (blend_bwd_kernel)
#ifndef MOBGS_BWD_TRANSPOSE
#define MOBGS_BWD_TRANSPOSE 1
#endif
template <int D, bool DEC, bool FLOW = false, bool POSE = false>
static void launch_bwd_kernel(const MobgsBlendBwd& a, int tiles_x, int tiles_y, cudaStream_t s) {
#if MOBGS_BWD_TRANSPOSE && MOBGS_UNIT_LANES == 16
  static bool configured = false;   // per instantiation; the attribute is idempotent, a race only repeats the call
  if (!configured) {
    cudaFuncSetAttribute(blend_bwd_tr_kernel<D, DEC, FLOW, POSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrSmemBytes);
    cudaFuncSetAttribute(blend_bwd_tr_kernel<D, DEC, FLOW, POSE>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    configured = true;
  }
  blend_bwd_tr_kernel<D, DEC, FLOW, POSE><<<a.K * tiles_x * tiles_y, kBlendThreads, kTrSmemBytes, s>>>(a, tiles_x, tiles_y);
#else
  static_assert(!FLOW && !POSE, "fused flow channels / in-kernel rays need the transposing backward");
  blend_bwd_kernel<D, DEC><<<a.K * tiles_x * tiles_y, kBlendThreads, 0, s>>>(a, tiles_x, tiles_y);
#endif
}
template <int D>
static void launch_bwd(const MobgsBlendBwd& a, int tiles_x, int tiles_y, cudaStream_t s) {
  launch_bwd_kernel<D, false>(a, tiles_x, tiles_y, s);
}

#define MOBGS_DISPATCH_D(D, FN, ...)                                   \
  switch (D) {                                                         \
    case 1: FN<1>(__VA_ARGS__); break;                                 \
    case 2: FN<2>(__VA_ARGS__); break;                                 \
    case 3: FN<3>(__VA_ARGS__); break;                                 \
    case 4: FN<4>(__VA_ARGS__); break;                                 \
    case 5: FN<5>(__VA_ARGS__); break;                                 \
    case 6: FN<6>(__VA_ARGS__); break;                                 \
    case 7: FN<7>(__VA_ARGS__); break;                                 \
    case 8: FN<8>(__VA_ARGS__); break;                                 \
    case 9: FN<9>(__VA_ARGS__); break;                                 \
    default: FN<10>(__VA_ARGS__); break;                               \
  }

extern "C" int mobgs_blend_fwd(const MobgsBlendFwd* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->K >= 1 && a->K <= MOBGS_MAX_K && a->width > 0 && a->height > 0, "bad extents");
  MOBGS_REQUIRE(a->D >= 1 && a->D <= MOBGS_MAX_COLORS, "D=%d out of range", a->D);
  MOBGS_REQUIRE(a->tile_offsets && a->out_colors && a->out_alphas && a->last_idx, "NULL pointer");
  MOBGS_REQUIRE(a->N == 0 || (a->records && a->sorted_ids), "NULL records / sorted_ids");
  const int tiles_x = (a->width + kTile - 1) / kTile, tiles_y = (a->height + kTile - 1) / kTile;
  if (a->dec_rays || a->dec_pose) {
    MOBGS_REQUIRE(a->D == 10 && a->dec_w1 && a->dec_w2 && a->out_rgb, "fused decode epilogue needs D=10, w1, w2, out_rgb");
    MOBGS_REQUIRE(!a->dec_pose || (a->dec_sfx != 0.f && a->dec_sfy != 0.f), "dec_pose needs non-zero scale factors");
    if (a->out_flow) {
#if !MOBGS_FWD_PREDICATED
      MOBGS_REQUIRE(a->flow_ref >= 0, "flow_ref must name a record set");
      if (a->dec_pose) blend_fwd_kernel<10, true, true, true><<<a->K * tiles_x * tiles_y, kBlendThreads, 0, (cudaStream_t)stream>>>(*a, tiles_x, tiles_y);
      else blend_fwd_kernel<10, true, true><<<a->K * tiles_x * tiles_y, kBlendThreads, 0, (cudaStream_t)stream>>>(*a, tiles_x, tiles_y);
      return check_launch("blend_decode_flow_fwd");
#else
      MOBGS_REQUIRE(false, "fused flow channels are not built into the predicated ablation build");
#endif
    }
    if (a->dec_pose) blend_fwd_kernel<10, true, false, true><<<a->K * tiles_x * tiles_y, kBlendThreads, 0, (cudaStream_t)stream>>>(*a, tiles_x, tiles_y);
    else blend_fwd_kernel<10, true><<<a->K * tiles_x * tiles_y, kBlendThreads, 0, (cudaStream_t)stream>>>(*a, tiles_x, tiles_y);
    return check_launch("blend_decode_fwd");
  }
  MOBGS_REQUIRE(!a->out_flow, "fused flow channels need the fused decode epilogue (dec_rays)");
  MOBGS_DISPATCH_D(a->D, launch_fwd, *a, tiles_x, tiles_y, (cudaStream_t)stream);
  return check_launch("blend_fwd");
}

extern "C" int mobgs_blend_bwd(const MobgsBlendBwd* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->K >= 1 && a->width > 0 && a->height > 0, "bad extents");
  MOBGS_REQUIRE(a->D >= 1 && a->D <= MOBGS_MAX_COLORS, "D=%d out of range", a->D);
  if (a->N == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->records && a->tile_offsets && a->sorted_ids && a->out_alphas && a->last_idx && a->v_records,
                "NULL pointer");
  const int tiles_x = (a->width + kTile - 1) / kTile, tiles_y = (a->height + kTile - 1) / kTile;
  if (a->dec_rays || a->dec_pose) {
    MOBGS_REQUIRE(a->D == 10 && a->dec_w1 && a->dec_w2 && a->out_colors && a->v_w_partial,
                  "fused decode prologue needs D=10, w1, w2, out_colors, v_w_partial");
    MOBGS_REQUIRE(!a->dec_pose || (a->dec_sfx != 0.f && a->dec_sfy != 0.f), "dec_pose needs non-zero scale factors");
    MOBGS_REQUIRE(a->dec_pose || !a->v_pose_partial, "v_pose_partial needs dec_pose");
    if (a->g_flow) {
#if MOBGS_BWD_TRANSPOSE && MOBGS_UNIT_LANES == 16 && MOBGS_BWD_DIRECT_RED
      MOBGS_REQUIRE(a->flow_ref >= 0, "flow_ref must name a record set");
      if (a->dec_pose) launch_bwd_kernel<10, true, true, true>(*a, tiles_x, tiles_y, (cudaStream_t)stream);
      else launch_bwd_kernel<10, true, true>(*a, tiles_x, tiles_y, (cudaStream_t)stream);
      return check_launch("blend_decode_flow_bwd");
#else
      MOBGS_REQUIRE(false, "fused flow channels need the transposing backward build");
#endif
    }
    if (a->dec_pose) {
#if MOBGS_BWD_TRANSPOSE && MOBGS_UNIT_LANES == 16
      launch_bwd_kernel<10, true, false, true>(*a, tiles_x, tiles_y, (cudaStream_t)stream);
#else
      MOBGS_REQUIRE(false, "in-kernel rays need the transposing backward build");
#endif
    } else {
      launch_bwd_kernel<10, true>(*a, tiles_x, tiles_y, (cudaStream_t)stream);
    }
    return check_launch("blend_decode_bwd");
  }
  MOBGS_REQUIRE(!a->g_flow, "fused flow channels need the fused decode prologue (dec_rays)");
  MOBGS_REQUIRE(a->v_out_colors, "v_out_colors must not be NULL");
  MOBGS_DISPATCH_D(a->D, launch_bwd, *a, tiles_x, tiles_y, (cudaStream_t)stream);
  return check_launch("blend_bwd");
}
