// TEST HARNESS ONLY — compiles mobgs_b200/csrc/gs_math.cuh for the host so the per-Gaussian
// arithmetic (projection fwd/VJP, Hermite taps) can be checked against the PyTorch oracle's
// autograd on a machine without a GPU.  Nothing in mobgs_b200/ loads this library.
#include "../../mobgs_b200/csrc/gs_math.cuh"
#include "../../mobgs_b200/csrc/blend_units.cuh"

using namespace mobgs;

// geom per Gaussian: mx my opac ca cb cc (tile-local frame: tile origin (0,0)); lanes in {32,16,8}.
// out_mask[g] = unit_mask<lanes>; out_need[g] = bit u set iff some pixel of unit u has
// opac * exp(-sigma) >= 1/255 with sigma >= 0 (what the blend inner loop would accept), in double.
template <int L>
static void unit_mask_case(int n, const float* geom, unsigned* out_mask, unsigned* out_need) {
  using G = UnitGeom<L>;
  for (int g = 0; g < n; ++g) {
    const float* r = geom + 6 * g;
    out_mask[g] = unit_mask<L>(r[0], r[1], r[2], r[3], r[4], r[5], 0.f, 0.f);
    unsigned need = 0;
    for (int u = 0; u < G::kUnits; ++u)
      for (int q = 0; q < L; ++q) {
        const int lx = (u % G::kUX) * G::kUW + q % G::kUW, ly = (u / G::kUX) * G::kUH + q / G::kUW;
        const double dx = (double)r[0] - (lx + 0.5), dy = (double)r[1] - (ly + 0.5);
        const double sigma = 0.5 * ((double)r[3] * dx * dx + (double)r[5] * dy * dy) + (double)r[4] * dx * dy;
        if (sigma >= 0 && (double)r[2] * exp(-sigma) >= 1.0 / 255.0) need |= 1u << u;
      }
    out_need[g] = need;
  }
}
extern "C" {

// viewmat: 16 floats row-major; K: 9 floats. out per Gaussian: mx my depth ca cb cc radius(as float)
void hm_project_fwd(int n, const float* means, const float* quats, const float* scales,
                    const float* viewmat, const float* Kmat, int width, int height, float eps2d,
                    float near_plane, float far_plane, float radius_clip, float* out) {
  Cam cam;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) cam.r[3 * i + j] = viewmat[4 * i + j];
    cam.t[i] = viewmat[4 * i + 3];
  }
  cam.fx = Kmat[0]; cam.cx = Kmat[2]; cam.fy = Kmat[4]; cam.cy = Kmat[5];
  ProjCfg cfg{width, height, eps2d, near_plane, far_plane, radius_clip};
  for (int g = 0; g < n; ++g) {
    ProjState st;
    ProjOut o = project_fwd(means + 3 * g, quats + 4 * g, scales + 3 * g, cam, cfg, st);
    float* r = out + 7 * g;
    r[0] = o.mx; r[1] = o.my; r[2] = o.depth; r[3] = o.ca; r[4] = o.cb; r[5] = o.cc; r[6] = (float)o.radius;
  }
}

// v_in per Gaussian: v_mx v_my v_depth v_ca v_cb v_cc.  grads out: p(3) q(4) s(3) per Gaussian;
// v_view: 12 floats (r row-major 9, t 3), summed over Gaussians.
void hm_project_bwd(int n, const float* means, const float* quats, const float* scales,
                    const float* viewmat, const float* Kmat, int width, int height, float eps2d,
                    float near_plane, float far_plane, float radius_clip, const float* v_in,
                    float* v_out, float* v_view) {
  Cam cam;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) cam.r[3 * i + j] = viewmat[4 * i + j];
    cam.t[i] = viewmat[4 * i + 3];
  }
  cam.fx = Kmat[0]; cam.cx = Kmat[2]; cam.fy = Kmat[4]; cam.cy = Kmat[5];
  ProjCfg cfg{width, height, eps2d, near_plane, far_plane, radius_clip};
  for (int i = 0; i < 12; ++i) v_view[i] = 0.f;
  for (int g = 0; g < n; ++g) {
    ProjState st;
    ProjOut o = project_fwd(means + 3 * g, quats + 4 * g, scales + 3 * g, cam, cfg, st);
    float* r = v_out + 10 * g;
    for (int i = 0; i < 10; ++i) r[i] = 0.f;
    if (o.radius <= 0) continue;
    const float* v = v_in + 6 * g;
    ProjGrad gr;
    project_bwd(means + 3 * g, scales + 3 * g, cam, st, v[0], v[1], v[2], v[3], v[4], v[5], gr);
    for (int i = 0; i < 3; ++i) r[i] = gr.p[i];
    for (int i = 0; i < 4; ++i) r[3 + i] = gr.q[i];
    for (int i = 0; i < 3; ++i) r[7 + i] = gr.s[i];
    for (int i = 0; i < 9; ++i) v_view[i] += gr.r[i];
    for (int i = 0; i < 3; ++i) v_view[9 + i] += gr.t[i];
  }
}

// out: idx[4] (as float) then w[4]
void hm_hermite_taps(float t, int n, float* out) {
  SplineTaps s = hermite_taps(t, n);
  for (int i = 0; i < 4; ++i) { out[i] = (float)s.idx[i]; out[4 + i] = s.w[i]; }
}

void hm_unit_mask(int lanes, int n, const float* geom, unsigned* out_mask, unsigned* out_need) {
  if (lanes == 32) unit_mask_case<32>(n, geom, out_mask, out_need);
  else if (lanes == 16) unit_mask_case<16>(n, geom, out_mask, out_need);
  else unit_mask_case<8>(n, geom, out_mask, out_need);
}

}
