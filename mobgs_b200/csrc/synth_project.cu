// K1'/K2': attribute synthesis + EWA projection, forward and VJP.
//
// One thread owns one Gaussian for *all* K sub-frames of the launch: the parameters (68 B static,
// 240 B dynamic) are read once, the K camera matrices sit in shared memory, and in the backward
// the per-parameter gradient is summed over K in registers and written exactly once — no atomics
// except the 12-float view-matrix gradient (one warp-reduced atomicAdd per warp and sub-frame).
// HBM-bound; see DESIGN.md for the byte model.
#include "common.cuh"
#include "tile_bin.cuh"

namespace mobgs {

constexpr int kProjThreads = 128;

struct CamSmem {
  Cam cam[kMaxK];
  float t_spline[kMaxK];
  float t_poly[kMaxK];
  float v_view[kMaxK][12];   // backward: CTA-level accumulator of the view-matrix gradient
};

__device__ __forceinline__ void load_cams(CamSmem& sm, const MobgsCameras& c, const float* t_spline,
                                          const float* t_poly) {
  for (int k = threadIdx.x; k < c.K; k += blockDim.x) {
    sm.cam[k] = load_cam(c.viewmats, c.Ks, k);
    sm.t_spline[k] = t_spline ? t_spline[k] : 0.f;
    sm.t_poly[k] = t_poly ? t_poly[k] : 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) sm.v_view[k][i] = 0.f;
  }
  __syncthreads();
}

__device__ __forceinline__ void store_record(float* rec, const ProjOut& o, float opac, const float col[10]) {
  float4* r4 = reinterpret_cast<float4*>(rec);
  if (o.radius > 0) {
    r4[0] = make_float4(o.mx, o.my, opac, o.ca);
    r4[1] = make_float4(o.cb, o.cc, col[0], col[1]);
    r4[2] = make_float4(col[2], col[3], col[4], col[5]);
    r4[3] = make_float4(col[6], col[7], col[8], col[9]);
  } else {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    r4[0] = z; r4[1] = z; r4[2] = z; r4[3] = z;
  }
}

// ------------------------------------------------------------------------------------------
// plain gsplat-style projection
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kProjThreads) project_fwd_kernel(MobgsProjectFwd a) {
  __shared__ CamSmem sm;
  load_cams(sm, a.cams, nullptr, nullptr);
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.N) return;
  const ProjCfg cfg = make_cfg(a.cams);
  float p[3], q[4], s[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { p[i] = a.means[3 * g + i]; s[i] = a.scales[3 * g + i]; }
  const float4 q4 = *reinterpret_cast<const float4*>(a.quats + 4 * g);
  q[0] = q4.x; q[1] = q4.y; q[2] = q4.z; q[3] = q4.w;
  for (int k = 0; k < a.cams.K; ++k) {
    ProjState st;
    const ProjOut o = project_fwd(p, q, s, sm.cam[k], cfg, st);
    const size_t i = (size_t)k * a.N + g;
    if (a.radii) a.radii[i] = o.radius;
    if (a.means2d) *reinterpret_cast<float2*>(a.means2d + 2 * i) = make_float2(o.mx, o.my);
    if (a.depths) a.depths[i] = o.depth;
    if (a.conics) { a.conics[3 * i] = o.ca; a.conics[3 * i + 1] = o.cb; a.conics[3 * i + 2] = o.cc; }
  }
}

// View-matrix gradient of one (warp, sub-frame): the 12 per-lane terms (zeros from lanes that did not
// contribute) are summed over the warp with a transposing butterfly — each stage halves the values a
// lane holds while doubling the lanes summed: 16 shuffles for the 12 (padded to 16) values instead of
// 60 — after which 12 different lanes each hold one complete sum and add it to the CTA's shared-memory
// accumulator in parallel (one lane doing 12 CAS-loop atomics in a row was 30 % of this kernel's stall
// samples).  The accumulator leaves the CTA as one atomicAdd per value at the very end: thousands of
// CTAs hammering the same 12*K global addresses would serialise in L2.
__device__ __forceinline__ void reduce_viewmat_grad(float (*v_view)[12], int k, const ProjGrad& g, bool active) {
  if (!__any_sync(0xffffffffu, active)) return;
  float v[16];
#pragma unroll
  for (int i = 0; i < 9; ++i) v[i] = active ? g.r[i] : 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) v[9 + i] = active ? g.t[i] : 0.f;
#pragma unroll
  for (int i = 12; i < 16; ++i) v[i] = 0.f;
  const int lane = threadIdx.x & 31;
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
  if ((lane & 1) == 0 && vidx < 12 && v[0] != 0.f) atomicAdd(&v_view[k][vidx], v[0]);
}

__device__ __forceinline__ void flush_viewmat_grad(float (*v_view)[12], float* v_viewmats, int K) {
  __syncthreads();
  for (int i = threadIdx.x; i < K * 12; i += blockDim.x) {
    const int k = i / 12, j = i - 12 * k;
    const float v = v_view[k][j];
    // v_view row layout: r[0..8] row-major, t[0..2]  ->  viewmat [4,4] row-major
    const int dst = j < 9 ? 4 * (j / 3) + (j % 3) : 4 * (j - 9) + 3;
    if (v != 0.f) atomicAdd(v_viewmats + 16 * k + dst, v);
  }
}

__global__ void __launch_bounds__(kProjThreads) project_bwd_kernel(MobgsProjectBwd a) {
  __shared__ CamSmem sm;
  load_cams(sm, a.cams, nullptr, nullptr);
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = g < a.N;
  const ProjCfg cfg = make_cfg(a.cams);
  float p[3] = {0, 0, 0}, q[4] = {1, 0, 0, 0}, s[3] = {1, 1, 1};
  if (in_range) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { p[i] = a.means[3 * g + i]; s[i] = a.scales[3 * g + i]; }
    const float4 q4 = *reinterpret_cast<const float4*>(a.quats + 4 * g);
    q[0] = q4.x; q[1] = q4.y; q[2] = q4.z; q[3] = q4.w;
  }
  float vp[3] = {0, 0, 0}, vq[4] = {0, 0, 0, 0}, vs[3] = {0, 0, 0};
  for (int k = 0; k < a.cams.K; ++k) {
    const size_t i = (size_t)k * a.N + g;
    const bool active = in_range && a.radii[i] > 0;
    ProjGrad gr;
    if (active) {
      ProjState st;
      proj_state(p, q, s, sm.cam[k], cfg, st);
      float vmx = 0, vmy = 0, vd = 0, vca = 0, vcb = 0, vcc = 0;
      if (a.v_means2d) { const float* v = a.v_means2d + i * a.v_means2d_stride; vmx = v[0]; vmy = v[1]; }
      if (a.v_depths) vd = a.v_depths[i * a.v_depths_stride];
      if (a.v_conics) { const float* v = a.v_conics + i * a.v_conics_stride; vca = v[0]; vcb = v[1]; vcc = v[2]; }
      project_bwd(p, s, sm.cam[k], st, vmx, vmy, vd, vca, vcb, vcc, gr);
#pragma unroll
      for (int j = 0; j < 3; ++j) { vp[j] += gr.p[j]; vs[j] += gr.s[j]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) vq[j] += gr.q[j];
    }
    if (a.v_viewmats) reduce_viewmat_grad(sm.v_view, k, gr, active);
  }
  if (a.v_viewmats) flush_viewmat_grad(sm.v_view, a.v_viewmats, a.cams.K);
  if (in_range) {
#pragma unroll
    for (int j = 0; j < 3; ++j) { a.v_means[3 * g + j] = vp[j]; a.v_scales[3 * g + j] = vs[j]; }
    *reinterpret_cast<float4*>(a.v_quats + 4 * g) = make_float4(vq[0], vq[1], vq[2], vq[3]);
  }
}

// ------------------------------------------------------------------------------------------
// fused MoBGS synthesis + projection
// ------------------------------------------------------------------------------------------
struct DynRow {
  float rot[4], omega[4], scale[3], opac, fdc[6], ft[3], trbf, off[3];
  int n_ctrl;
};

__device__ __forceinline__ void load_dyn(const MobgsDynamicParams& d, int j, DynRow& r) {
  const float4 r4 = *reinterpret_cast<const float4*>(d.rotation + 4 * j);
  const float4 o4 = *reinterpret_cast<const float4*>(d.omega + 4 * j);
  r.rot[0] = r4.x; r.rot[1] = r4.y; r.rot[2] = r4.z; r.rot[3] = r4.w;
  r.omega[0] = o4.x; r.omega[1] = o4.y; r.omega[2] = o4.z; r.omega[3] = o4.w;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    r.scale[i] = expf(d.scaling[3 * j + i]);
    r.ft[i] = d.features_t[3 * j + i];
    r.off[i] = d.offset ? d.offset[3 * j + i] : 0.f;
  }
  const float2* f2 = reinterpret_cast<const float2*>(d.features_dc + 6 * j);
#pragma unroll
  for (int i = 0; i < 3; ++i) { const float2 f = f2[i]; r.fdc[2 * i] = f.x; r.fdc[2 * i + 1] = f.y; }
  r.opac = sigmoidf(d.opacity[j]);
  r.trbf = d.trbf_center[j];
  int n = (int)d.control_num[j];
  n = n < 2 ? 2 : (n > d.n_ctrl_max ? d.n_ctrl_max : n);
  r.n_ctrl = n;
}

// Fused counting pass (MobgsSynthFwd.bin_tile_counts): the lists that read record set k take this Gaussian while its
// projection is still in registers.  The loop over the lists is warp-uniform (k and the list table are); a lane whose
// Gaussian is outside the list's range or culled takes part with live = false (bin_count_record reserves per warp).
__device__ __forceinline__ void synth_bin(const MobgsSynthFwd& a, const BinTarget& bt, int k, int g, const ProjOut& o, float opac) {
  const GaussGeom geo = {o.mx, o.my, opac, o.ca, o.cb, o.cc};
  for (int l = 0; l < a.bin_n_lists; ++l) {
    if (a.bin_lists.rec_k[l] != k) continue;
    const bool live = o.radius > 0 && g >= a.bin_lists.g_begin[l] && g < a.bin_lists.g_end[l];
    bin_count_record(bt, l, g, live, geo, o.radius, o.depth);
  }
}

__global__ void __launch_bounds__(kProjThreads) synth_project_fwd_kernel(const __grid_constant__ MobgsSynthFwd a) {
  __shared__ CamSmem sm;
  load_cams(sm, a.cams, a.t_spline, a.t_poly);
  const int N = a.st.Ns + a.dy.Nd;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  const ProjCfg cfg = make_cfg(a.cams);
  const int K = a.cams.K;
  const bool binning = a.bin_tile_counts != nullptr;
  const int tiles_x = (a.cams.width + kTile - 1) / kTile, tiles_y = (a.cams.height + kTile - 1) / kTile;
  const BinTarget bt = {a.bin_tile_counts, a.bin_entries, a.bin_entry_capacity, a.bin_entry_cursor,
                        a.cams.width, a.cams.height, tiles_x, tiles_y, a.bin_tight};
  if (g < a.st.Ns) {
    float p[3], q[4], s[3], col[10];
#pragma unroll
    for (int i = 0; i < 3; ++i) { p[i] = a.st.xyz[3 * g + i]; s[i] = expf(a.st.scaling[3 * g + i]); }
    const float4 q4 = *reinterpret_cast<const float4*>(a.st.rotation + 4 * g);
    q[0] = q4.x; q[1] = q4.y; q[2] = q4.z; q[3] = q4.w;
    const float opac = sigmoidf(a.st.opacity[g]);
#pragma unroll
    for (int i = 0; i < 6; ++i) col[i] = a.st.features_dc[6 * g + i];
    col[6] = col[7] = col[8] = 0.f;
    for (int k = 0; k < K; ++k) {
      ProjState st;
      const ProjOut o = project_fwd(p, q, s, sm.cam[k], cfg, st);
      const size_t i = (size_t)k * N + g;
      col[9] = o.depth;
      store_record(a.records + i * kRecFloats, o, opac, col);
      a.radii[i] = o.radius;
      a.depths[i] = o.depth;
      if (a.means3d) { a.means3d[3 * i] = p[0]; a.means3d[3 * i + 1] = p[1]; a.means3d[3 * i + 2] = p[2]; }
      if (binning) synth_bin(a, bt, k, g, o, opac);
    }
  } else {
    const int j = g - a.st.Ns;
    DynRow r;
    load_dyn(a.dy, j, r);
    const float* ctrl = a.dy.control_xyz + (size_t)j * a.dy.n_ctrl_max * 3;
    float col[10];
#pragma unroll
    for (int i = 0; i < 6; ++i) col[i] = r.fdc[i];
    for (int k = 0; k < K; ++k) {
      const SplineTaps tp = hermite_taps(sm.t_spline[k], r.n_ctrl);
      float p[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float* c = ctrl + 3 * tp.idx[t];
        p[0] += tp.w[t] * c[0]; p[1] += tp.w[t] * c[1]; p[2] += tp.w[t] * c[2];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) p[i] = p[i] * 1e-2f + r.off[i];
      const float dt = sm.t_poly[k] - r.trbf;
      float q[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) q[i] = r.rot[i] + dt * r.omega[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) col[6 + i] = dt * r.ft[i];
      ProjState st;
      const ProjOut o = project_fwd(p, q, r.scale, sm.cam[k], cfg, st);
      const size_t i = (size_t)k * N + g;
      col[9] = o.depth;
      store_record(a.records + i * kRecFloats, o, r.opac, col);
      a.radii[i] = o.radius;
      a.depths[i] = o.depth;
      if (a.means3d) { a.means3d[3 * i] = p[0]; a.means3d[3 * i + 1] = p[1]; a.means3d[3 * i + 2] = p[2]; }
      if (binning) synth_bin(a, bt, k, g, o, r.opac);
    }
  }
}

#ifndef MOBGS_CTRL_RED
#define MOBGS_CTRL_RED 1
#endif
#ifndef MOBGS_SYNTH_BWD_MIN_CTAS
#define MOBGS_SYNTH_BWD_MIN_CTAS 3
#endif
__global__ void __launch_bounds__(kProjThreads, MOBGS_SYNTH_BWD_MIN_CTAS) synth_project_bwd_kernel(const __grid_constant__ MobgsSynthBwd a) {
  __shared__ CamSmem sm;
  load_cams(sm, a.cams, a.t_spline, a.t_poly);
  const int N = a.st.Ns + a.dy.Nd;
  // [g_lo, g_hi): the Gaussians this launch differentiates (0, 0 = all) — a data-parallel caller splits the
  // backward into ranges so that the all-reduce of one range's gradients overlaps the next range's kernel
  const int g = a.g_lo + blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = g < (a.g_hi > 0 ? a.g_hi : N);
  const bool is_static = g < a.st.Ns;
  const ProjCfg cfg = make_cfg(a.cams);
  const int K = a.cams.K;
  const int j = g - a.st.Ns;
  // pose-only mode (every parameter-gradient pointer NULL): only v_viewmats is produced — eval.py's test-time
  // pose optimisation freezes all Gaussians (eval.py:246-255) and differentiates render(w2c=...) w.r.t. the pose
  const bool pose_only = a.v_xyz == nullptr && a.v_control_xyz == nullptr;
  const bool v_ctrl_on = !pose_only;

  float p[3] = {0, 0, 0}, s[3] = {1, 1, 1}, q0[4] = {1, 0, 0, 0};
  float opac = 0.f;
  DynRow r;
  r.n_ctrl = 2; r.trbf = 0.f;
  const float* ctrl = nullptr;
  float* v_ctrl = nullptr;
  if (in_range) {
    if (is_static) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { p[i] = a.st.xyz[3 * g + i]; s[i] = expf(a.st.scaling[3 * g + i]); }
      const float4 q4 = *reinterpret_cast<const float4*>(a.st.rotation + 4 * g);
      q0[0] = q4.x; q0[1] = q4.y; q0[2] = q4.z; q0[3] = q4.w;
      opac = sigmoidf(a.st.opacity[g]);
    } else {
      load_dyn(a.dy, j, r);
      ctrl = a.dy.control_xyz + (size_t)j * a.dy.n_ctrl_max * 3;
      if (v_ctrl_on) v_ctrl = a.v_control_xyz + (size_t)j * a.dy.n_ctrl_max * 3;
#pragma unroll
      for (int i = 0; i < 3; ++i) s[i] = r.scale[i];
      opac = r.opac;
    }
  }
  float vp[3] = {0, 0, 0}, vq[4] = {0, 0, 0, 0}, vs[3] = {0, 0, 0}, vom[4] = {0, 0, 0, 0};
  float vop = 0.f, vfdc[6] = {0, 0, 0, 0, 0, 0}, vft[3] = {0, 0, 0};

  for (int k = 0; k < K; ++k) {
    const size_t i = (size_t)k * N + g;
    const bool active = in_range && a.radii[i] > 0;
    ProjGrad gr;
    if (active) {
      float q[4];
      float dt = 0.f;
      SplineTaps tp;
      if (is_static) {
#pragma unroll
        for (int c = 0; c < 4; ++c) q[c] = q0[c];
      } else {
        tp = hermite_taps(sm.t_spline[k], r.n_ctrl);
        p[0] = p[1] = p[2] = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float* c = ctrl + 3 * tp.idx[t];
          p[0] += tp.w[t] * c[0]; p[1] += tp.w[t] * c[1]; p[2] += tp.w[t] * c[2];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) p[c] = p[c] * 1e-2f + r.off[c];
        dt = sm.t_poly[k] - r.trbf;
#pragma unroll
        for (int c = 0; c < 4; ++c) q[c] = r.rot[c] + dt * r.omega[c];
      }
      ProjState st;
      proj_state(p, q, s, sm.cam[k], cfg, st);
      const float4* v4 = reinterpret_cast<const float4*>(a.v_records + i * kRecFloats);
      const float4 v0 = v4[0], v1 = v4[1], v2 = v4[2], v3 = v4[3];
      // layout: x y opac ca | cb cc c0 c1 | c2 c3 c4 c5 | c6 c7 c8 c9(depth)
      project_bwd(p, s, sm.cam[k], st, v0.x, v0.y, v3.w, v0.w, v1.x, v1.y, gr);
      vop += v0.z;
      vfdc[0] += v1.z; vfdc[1] += v1.w; vfdc[2] += v2.x; vfdc[3] += v2.y; vfdc[4] += v2.z; vfdc[5] += v2.w;
#pragma unroll
      for (int c = 0; c < 3; ++c) vs[c] += gr.s[c];
#pragma unroll
      for (int c = 0; c < 4; ++c) vq[c] += gr.q[c];
      if (is_static) {
#pragma unroll
        for (int c = 0; c < 3; ++c) vp[c] += gr.p[c];
      } else {
        vft[0] += dt * v3.x; vft[1] += dt * v3.y; vft[2] += dt * v3.z;
#pragma unroll
        for (int c = 0; c < 4; ++c) vom[c] += dt * gr.q[c];
#pragma unroll
        for (int c = 0; c < 3; ++c) vp[c] += gr.p[c];   // gradient of the `coherent` offset
        // this thread is the only writer of its control-point row: plain read-modify-write
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float w = tp.w[t] * 1e-2f;
          if (w != 0.f && v_ctrl_on) {
            float* c = v_ctrl + 3 * tp.idx[t];
#if MOBGS_CTRL_RED
            // fire-and-forget reductions: no load in the dependent chain (the row is zeroed by the caller)
            atomicAdd(c, w * gr.p[0]); atomicAdd(c + 1, w * gr.p[1]); atomicAdd(c + 2, w * gr.p[2]);
#else
            c[0] += w * gr.p[0]; c[1] += w * gr.p[1]; c[2] += w * gr.p[2];
#endif
          }
        }
      }
    }
    if (a.v_viewmats) reduce_viewmat_grad(sm.v_view, k, gr, active);
    if (a.zero_v_records) {
      // Hand the buffer back zeroed (the next backward's scatter target then needs no memset).  The 32 gradient records
      // a warp has just read for sub-frame k are 2 KB of contiguous memory: the warp zeroes them — all of them, read or
      // not — with four fully coalesced 512-byte stores.  (Each lane zeroing its own record behind its loads, 16-byte
      // stores 64 bytes apart, made this kernel 3.7x slower.)
      __syncwarp();                                      // every lane's loads of this chunk are done
      const int g0 = g - (int)(threadIdx.x & 31);        // first Gaussian of the warp
      const int g_end = a.g_hi > 0 ? a.g_hi : N;
      float4* chunk = reinterpret_cast<float4*>(a.v_records + ((size_t)k * N + g0) * kRecFloats);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int f4 = jj * 32 + (int)(threadIdx.x & 31);          // float4 index inside the chunk: record g0 + f4 / 4
        if (g0 + f4 / 4 < g_end) __stcs(chunk + f4, z);
      }
    }
  }
  if (a.v_viewmats) flush_viewmat_grad(sm.v_view, a.v_viewmats, K);
  if (!in_range || pose_only) return;
  const float vo_logit = vop * opac * (1.f - opac);
  if (is_static) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { a.v_xyz[3 * g + c] = vp[c]; a.v_scaling_s[3 * g + c] = vs[c] * s[c]; }
    *reinterpret_cast<float4*>(a.v_rotation_s + 4 * g) = make_float4(vq[0], vq[1], vq[2], vq[3]);
    a.v_opacity_s[g] = vo_logit;
#pragma unroll
    for (int c = 0; c < 6; ++c) a.v_features_dc_s[6 * g + c] = vfdc[c];
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      a.v_scaling_d[3 * j + c] = vs[c] * s[c];
      a.v_features_t[3 * j + c] = vft[c];
      if (a.v_offset) a.v_offset[3 * j + c] = vp[c];
    }
    *reinterpret_cast<float4*>(a.v_rotation_d + 4 * j) = make_float4(vq[0], vq[1], vq[2], vq[3]);
    *reinterpret_cast<float4*>(a.v_omega + 4 * j) = make_float4(vom[0], vom[1], vom[2], vom[3]);
    a.v_opacity_d[j] = vo_logit;
#pragma unroll
    for (int c = 0; c < 6; ++c) a.v_features_dc_d[6 * j + c] = vfdc[c];
  }
}

// ------------------------------------------------------------------------------------------
// SoA -> packed records (gsplat-compatible operator path)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_records_kernel(MobgsPack a) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)a.K * a.N;
  if (i >= total) return;
  const int g = (int)(i % a.N);
  float col[10];
#pragma unroll
  for (int c = 0; c < 10; ++c) col[c] = 0.f;
  const float* cp = a.colors_per_cam ? a.colors + i * a.D : a.colors + (size_t)g * a.D;
  for (int c = 0; c < a.D; ++c) col[c] = cp[c];
  if (a.depths) col[a.D] = a.depths[i];
  float4* r4 = reinterpret_cast<float4*>(a.records + i * kRecFloats);
  const float2 m = *reinterpret_cast<const float2*>(a.means2d + 2 * i);
  const float* cn = a.conics + 3 * i;
  r4[0] = make_float4(m.x, m.y, a.opacities[g], cn[0]);
  r4[1] = make_float4(cn[1], cn[2], col[0], col[1]);
  r4[2] = make_float4(col[2], col[3], col[4], col[5]);
  r4[3] = make_float4(col[6], col[7], col[8], col[9]);
}

}  // namespace mobgs

using namespace mobgs;

static int check_cams(const MobgsCameras& c) {
  MOBGS_REQUIRE(c.K >= 1 && c.K <= kMaxK, "K=%d out of range [1,%d]", c.K, kMaxK);
  MOBGS_REQUIRE(c.width > 0 && c.height > 0, "bad image size %dx%d", c.width, c.height);
  MOBGS_REQUIRE(c.viewmats && c.Ks, "viewmats / Ks must not be NULL");
  return MOBGS_OK;
}

extern "C" int mobgs_project_fwd(const MobgsProjectFwd* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  if (int e = check_cams(a->cams)) return e;
  MOBGS_REQUIRE(a->N >= 0, "N < 0");
  if (a->N == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->means && a->quats && a->scales, "means/quats/scales must not be NULL");
  const int grid = (a->N + kProjThreads - 1) / kProjThreads;
  project_fwd_kernel<<<grid, kProjThreads, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("project_fwd");
}

extern "C" int mobgs_project_bwd(const MobgsProjectBwd* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  if (int e = check_cams(a->cams)) return e;
  if (a->N == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->means && a->quats && a->scales && a->radii, "inputs must not be NULL");
  MOBGS_REQUIRE(a->v_means && a->v_quats && a->v_scales, "gradient outputs must not be NULL");
  const int grid = (a->N + kProjThreads - 1) / kProjThreads;
  project_bwd_kernel<<<grid, kProjThreads, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("project_bwd");
}

static int check_params(const MobgsStaticParams& s, const MobgsDynamicParams& d) {
  MOBGS_REQUIRE(s.Ns >= 0 && d.Nd >= 0, "negative Gaussian count");
  if (s.Ns > 0) MOBGS_REQUIRE(s.xyz && s.rotation && s.scaling && s.opacity && s.features_dc, "static params NULL");
  if (d.Nd > 0) {
    MOBGS_REQUIRE(d.control_xyz && d.control_num && d.rotation && d.omega && d.scaling && d.opacity &&
                      d.features_dc && d.features_t && d.trbf_center, "dynamic params NULL");
    MOBGS_REQUIRE(d.n_ctrl_max >= 2 && d.n_ctrl_max <= 64, "n_ctrl_max=%d out of range", d.n_ctrl_max);
  }
  return MOBGS_OK;
}

extern "C" int mobgs_synth_project_fwd(const MobgsSynthFwd* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  if (int e = check_cams(a->cams)) return e;
  if (int e = check_params(a->st, a->dy)) return e;
  const int N = a->st.Ns + a->dy.Nd;
  if (N == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->records && a->radii && a->depths, "outputs must not be NULL");
  MOBGS_REQUIRE(a->dy.Nd == 0 || (a->t_spline && a->t_poly), "t_spline / t_poly must not be NULL");
  if (a->bin_tile_counts) {
    MOBGS_REQUIRE(a->bin_n_lists >= 1 && a->bin_n_lists <= MOBGS_MAX_K, "bin_n_lists=%d out of range", a->bin_n_lists);
    MOBGS_REQUIRE(!a->bin_entries || (a->bin_entry_cursor && a->bin_entry_capacity >= 0), "bin_entries need a cursor and a capacity");
    const int tiles = ((a->cams.width + kTile - 1) / kTile) * ((a->cams.height + kTile - 1) / kTile);
    cudaMemsetAsync(a->bin_tile_counts, 0, sizeof(int) * (size_t)a->bin_n_lists * tiles, (cudaStream_t)stream);
    if (a->bin_entry_cursor) cudaMemsetAsync(a->bin_entry_cursor, 0, sizeof(int), (cudaStream_t)stream);
  }
  const int grid = (N + kProjThreads - 1) / kProjThreads;
  synth_project_fwd_kernel<<<grid, kProjThreads, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("synth_project_fwd");
}

extern "C" int mobgs_synth_project_bwd(const MobgsSynthBwd* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  if (int e = check_cams(a->cams)) return e;
  if (int e = check_params(a->st, a->dy)) return e;
  const int N = a->st.Ns + a->dy.Nd;
  if (N == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->radii && a->v_records, "radii / v_records must not be NULL");
  const bool none = !a->v_xyz && !a->v_rotation_s && !a->v_scaling_s && !a->v_opacity_s && !a->v_features_dc_s &&
                    !a->v_control_xyz && !a->v_rotation_d && !a->v_omega && !a->v_scaling_d && !a->v_opacity_d &&
                    !a->v_features_dc_d && !a->v_features_t && !a->v_offset;
  if (none) {   // pose-only: eval.py's test-time pose optimisation (all Gaussians frozen)
    MOBGS_REQUIRE(a->v_viewmats, "pose-only backward needs v_viewmats");
  } else {
    const int lo = a->g_lo, hi = a->g_hi > 0 ? a->g_hi : N;        // only the tensors the range touches are needed
    if (a->st.Ns > 0 && lo < a->st.Ns)
      MOBGS_REQUIRE(a->v_xyz && a->v_rotation_s && a->v_scaling_s && a->v_opacity_s && a->v_features_dc_s,
                    "static gradient outputs NULL");
    if (a->dy.Nd > 0 && hi > a->st.Ns)
      MOBGS_REQUIRE(a->v_control_xyz && a->v_rotation_d && a->v_omega && a->v_scaling_d && a->v_opacity_d &&
                        a->v_features_dc_d && a->v_features_t, "dynamic gradient outputs NULL");
  }
  MOBGS_REQUIRE(a->g_lo >= 0 && a->g_hi >= 0 && a->g_hi <= N && (a->g_hi == 0 || a->g_lo < a->g_hi),
                "bad Gaussian range [%d, %d)", a->g_lo, a->g_hi);
  const int count = (a->g_hi > 0 ? a->g_hi : N) - a->g_lo;
  const int grid = (count + kProjThreads - 1) / kProjThreads;
  synth_project_bwd_kernel<<<grid, kProjThreads, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("synth_project_bwd");
}

extern "C" int mobgs_pack_records(const MobgsPack* a, void* stream) {
  MOBGS_REQUIRE(a, "NULL args");
  MOBGS_REQUIRE(a->K >= 1 && a->N >= 0, "bad K/N");
  const int dtot = a->D + (a->depths ? 1 : 0);
  MOBGS_REQUIRE(a->D >= 1 && dtot <= MOBGS_MAX_COLORS, "D=%d (+depth) exceeds %d channels", a->D, MOBGS_MAX_COLORS);
  if (a->N == 0) return MOBGS_OK;
  MOBGS_REQUIRE(a->means2d && a->conics && a->opacities && a->colors && a->records, "NULL pointer");
  const size_t total = (size_t)a->K * a->N;
  const int grid = (int)((total + 255) / 256);
  pack_records_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("pack_records");
}
