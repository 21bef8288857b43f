// Per-Gaussian arithmetic of the MoBGS render path, shared by all kernels.
//
// Everything here is `MOBGS_HD` (host + device) so that tests/host_math/ can run the
// *same* source on the CPU against the autograd of the PyTorch oracle without a GPU.
// That host build is a debugging harness only — nothing in mobgs_b200/ links it.
//
// Reference semantics restated:
//   * cubic Hermite control-point spline   gaussian_renderer/__init__.py:23-56
//   * activations / time-linear terms      scene/gaussian_model.py:98-106, 222-224, 241-246
//   * gsplat 1.4.0 fully_fused_projection  (external dependency, README.md:26) fwd + VJP
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MOBGS_HD __host__ __device__ __forceinline__
#else
#define MOBGS_HD inline
#endif

namespace mobgs {

// gsplat 1.4.0 constants that are not runtime arguments
constexpr float kFovMargin = 0.3f;        // persp_proj clamp margin (x tan(fov/2))
constexpr float kRadiusDiscFloor = 0.01f;  // max(0.01, b^2 - det) under the eigenvalue sqrt
constexpr float kAlphaMax = 0.999f;
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kTStop = 1e-4f;
constexpr int kRecFloats = 16;             // packed record: x y opac ca | cb cc c0 c1 | c2..c5 | c6..c9
constexpr int kMaxColors = 10;

struct Cam {
  float r[9];   // rotation, row-major (viewmat[:3,:3])
  float t[3];   // translation      (viewmat[:3,3])
  float fx, fy, cx, cy;
};

struct ProjCfg {
  int width, height;
  float eps2d, near_plane, far_plane, radius_clip;
};

// Everything the VJP needs again is recomputed from the inputs; nothing is cached.
struct ProjOut {
  float mx, my;       // mean2d (pixels)
  float depth;        // camera-space z
  float ca, cb, cc;   // conic = inverse of blurred 2D covariance (upper triangle)
  int radius;         // 0 => culled, other fields are 0
};

MOBGS_HD void quat_to_rot(const float qn[4], float R[9]) {
  const float w = qn[0], x = qn[1], y = qn[2], z = qn[3];
  R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - w * z); R[2] = 2.f * (x * z + w * y);
  R[3] = 2.f * (x * y + w * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - w * x);
  R[6] = 2.f * (x * z - w * y); R[7] = 2.f * (y * z + w * x); R[8] = 1.f - 2.f * (x * x + y * y);
}

// Intermediate state shared by forward and backward.
struct ProjState {
  float qn[4], inv_norm;
  float Rq[9];        // rotation of the Gaussian
  float A[9];         // cam.r * Rq
  float W[9];         // A * diag(s)     (Sigma_cam = W W^T)
  float pc[3];        // camera-space mean
  float rz, tx, ty;   // 1/z and frustum-clamped x,y used only inside J
  bool x_clamped, y_clamped;
  float j00, j02, j11, j12;
  float U[6];         // J W  (2x3)     (Sigma_2d = U U^T)
  float a, b, c, det; // blurred 2D covariance and its determinant
};

// Returns false when the Gaussian is culled before the conic exists (z range / det<=0).
MOBGS_HD bool proj_state(const float p[3], const float q[4], const float s[3], const Cam& cam,
                         const ProjCfg& cfg, ProjState& st) {
  const float* r = cam.r;
  st.pc[0] = r[0] * p[0] + r[1] * p[1] + r[2] * p[2] + cam.t[0];
  st.pc[1] = r[3] * p[0] + r[4] * p[1] + r[5] * p[2] + cam.t[1];
  st.pc[2] = r[6] * p[0] + r[7] * p[1] + r[8] * p[2] + cam.t[2];
  const float z = st.pc[2];
  if (!(z >= cfg.near_plane) || !(z <= cfg.far_plane)) return false;

  const float n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  st.inv_norm = 1.0f / sqrtf(n2);
  for (int i = 0; i < 4; ++i) st.qn[i] = q[i] * st.inv_norm;
  quat_to_rot(st.qn, st.Rq);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const float aij = r[3 * i] * st.Rq[j] + r[3 * i + 1] * st.Rq[3 + j] + r[3 * i + 2] * st.Rq[6 + j];
      st.A[3 * i + j] = aij;
      st.W[3 * i + j] = aij * s[j];
    }

  const float tan_fovx = 0.5f * cfg.width / cam.fx, tan_fovy = 0.5f * cfg.height / cam.fy;
  const float lim_xp = (cfg.width - cam.cx) / cam.fx + kFovMargin * tan_fovx;
  const float lim_xn = cam.cx / cam.fx + kFovMargin * tan_fovx;
  const float lim_yp = (cfg.height - cam.cy) / cam.fy + kFovMargin * tan_fovy;
  const float lim_yn = cam.cy / cam.fy + kFovMargin * tan_fovy;
  st.rz = 1.0f / z;
  const float xz = st.pc[0] * st.rz, yz = st.pc[1] * st.rz;
  st.x_clamped = !(xz <= lim_xp && xz >= -lim_xn);
  st.y_clamped = !(yz <= lim_yp && yz >= -lim_yn);
  st.tx = z * fminf(lim_xp, fmaxf(-lim_xn, xz));
  st.ty = z * fminf(lim_yp, fmaxf(-lim_yn, yz));
  const float rz2 = st.rz * st.rz;
  st.j00 = cam.fx * st.rz; st.j02 = -cam.fx * st.tx * rz2;
  st.j11 = cam.fy * st.rz; st.j12 = -cam.fy * st.ty * rz2;
  float a = 0.f, b = 0.f, c = 0.f;
  for (int j = 0; j < 3; ++j) {
    const float u0 = st.j00 * st.W[j] + st.j02 * st.W[6 + j];
    const float u1 = st.j11 * st.W[3 + j] + st.j12 * st.W[6 + j];
    st.U[j] = u0; st.U[3 + j] = u1;
    a += u0 * u0; b += u0 * u1; c += u1 * u1;
  }
  st.a = a + cfg.eps2d; st.b = b; st.c = c + cfg.eps2d;
  st.det = st.a * st.c - st.b * st.b;
  return st.det > 0.f;
}

MOBGS_HD ProjOut project_fwd(const float p[3], const float q[4], const float s[3], const Cam& cam,
                             const ProjCfg& cfg, ProjState& st) {
  ProjOut o; o.mx = o.my = o.depth = o.ca = o.cb = o.cc = 0.f; o.radius = 0;
  if (!proj_state(p, q, s, cam, cfg, st)) return o;
  const float mid = 0.5f * (st.a + st.c);
  const float v1 = mid + sqrtf(fmaxf(kRadiusDiscFloor, mid * mid - st.det));
  const float radius = ceilf(3.f * sqrtf(v1));
  if (radius <= cfg.radius_clip) return o;
  const float mx = cam.fx * st.pc[0] * st.rz + cam.cx, my = cam.fy * st.pc[1] * st.rz + cam.cy;
  if (mx + radius <= 0.f || mx - radius >= (float)cfg.width || my + radius <= 0.f ||
      my - radius >= (float)cfg.height)
    return o;
  const float inv_det = 1.0f / st.det;
  o.mx = mx; o.my = my; o.depth = st.pc[2];
  o.ca = st.c * inv_det; o.cb = -st.b * inv_det; o.cc = st.a * inv_det;
  o.radius = (int)radius;
  return o;
}

struct ProjGrad {
  float p[3], q[4], s[3];   // wrt world mean, raw quaternion, (activated) scale
  float r[9], t[3];         // wrt viewmat rotation / translation (per Gaussian contribution)
};

// VJP of project_fwd for a Gaussian that was *not* culled.  `st` must come from the same inputs.
MOBGS_HD void project_bwd(const float p[3], const float s[3], const Cam& cam, const ProjState& st,
                          float v_mx, float v_my, float v_depth, float v_ca, float v_cb, float v_cc,
                          ProjGrad& g) {
  // conic = inv(Sigma2): v_Sigma2 = -Q G Q, Q = conic, G = sym(v_conic)
  const float inv_det = 1.0f / st.det;
  const float qa = st.c * inv_det, qb = -st.b * inv_det, qc = st.a * inv_det;
  const float ga = v_ca, gb = 0.5f * v_cb, gc = v_cc;
  // T = Q G
  const float t00 = qa * ga + qb * gb, t01 = qa * gb + qb * gc;
  const float t10 = qb * ga + qc * gb, t11 = qb * gb + qc * gc;
  const float Va = -(t00 * qa + t01 * qb);
  const float Vb = -(t00 * qb + t01 * qc);
  const float Vc = -(t10 * qb + t11 * qc);
  // Sigma2 = U U^T -> v_U = 2 V U
  float vU[6];
  for (int j = 0; j < 3; ++j) {
    vU[j] = 2.f * (Va * st.U[j] + Vb * st.U[3 + j]);
    vU[3 + j] = 2.f * (Vb * st.U[j] + Vc * st.U[3 + j]);
  }
  // U = J W
  float v_j00 = 0.f, v_j02 = 0.f, v_j11 = 0.f, v_j12 = 0.f;
  float vW[9];
  for (int j = 0; j < 3; ++j) {
    v_j00 += vU[j] * st.W[j];      v_j02 += vU[j] * st.W[6 + j];
    v_j11 += vU[3 + j] * st.W[3 + j]; v_j12 += vU[3 + j] * st.W[6 + j];
    vW[j] = st.j00 * vU[j];
    vW[3 + j] = st.j11 * vU[3 + j];
    vW[6 + j] = st.j02 * vU[j] + st.j12 * vU[3 + j];
  }
  // W = A diag(s);  A = cam.r Rq
  float vA[9];
  for (int j = 0; j < 3; ++j) {
    g.s[j] = vW[j] * st.A[j] + vW[3 + j] * st.A[3 + j] + vW[6 + j] * st.A[6 + j];
    vA[j] = vW[j] * s[j]; vA[3 + j] = vW[3 + j] * s[j]; vA[6 + j] = vW[6 + j] * s[j];
  }
  float vRq[9];
  const float* r = cam.r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      // v_r = vA Rq^T ; v_Rq = r^T vA
      g.r[3 * i + j] = vA[3 * i] * st.Rq[3 * j] + vA[3 * i + 1] * st.Rq[3 * j + 1] + vA[3 * i + 2] * st.Rq[3 * j + 2];
      vRq[3 * i + j] = r[i] * vA[j] + r[3 + i] * vA[3 + j] + r[6 + i] * vA[6 + j];
    }
  // camera-space mean
  const float x = st.pc[0], y = st.pc[1];
  const float rz = st.rz, rz2 = rz * rz, rz3 = rz2 * rz;
  float vx = cam.fx * rz * v_mx;
  float vy = cam.fy * rz * v_my;
  float vz = -(cam.fx * x * v_mx + cam.fy * y * v_my) * rz2 + v_depth;
  vz += -cam.fx * rz2 * v_j00 - cam.fy * rz2 * v_j11;
  if (!st.x_clamped) { vx += -cam.fx * rz2 * v_j02; vz += 2.f * cam.fx * st.tx * rz3 * v_j02; }
  else               { vz += cam.fx * st.tx * rz3 * v_j02; }
  if (!st.y_clamped) { vy += -cam.fy * rz2 * v_j12; vz += 2.f * cam.fy * st.ty * rz3 * v_j12; }
  else               { vz += cam.fy * st.ty * rz3 * v_j12; }
  const float vpc[3] = {vx, vy, vz};
  for (int i = 0; i < 3; ++i) {
    g.t[i] = vpc[i];
    g.p[i] = r[i] * vpc[0] + r[3 + i] * vpc[1] + r[6 + i] * vpc[2];
    for (int j = 0; j < 3; ++j) g.r[3 * i + j] += vpc[i] * p[j];
  }
  // rotation matrix -> normalised quaternion -> raw quaternion
  const float w = st.qn[0], qx = st.qn[1], qy = st.qn[2], qz = st.qn[3];
  const float* v = vRq;
  float vq[4];
  vq[0] = 2.f * (-qz * v[1] + qy * v[2] + qz * v[3] - qx * v[5] - qy * v[6] + qx * v[7]);
  vq[1] = 2.f * (qy * v[1] + qz * v[2] + qy * v[3] - 2.f * qx * v[4] - w * v[5] + qz * v[6] + w * v[7] - 2.f * qx * v[8]);
  vq[2] = 2.f * (-2.f * qy * v[0] + qx * v[1] + w * v[2] + qx * v[3] + qz * v[5] - w * v[6] + qz * v[7] - 2.f * qy * v[8]);
  vq[3] = 2.f * (-2.f * qz * v[0] - w * v[1] + qx * v[2] + w * v[3] - 2.f * qz * v[4] + qy * v[5] + qx * v[6] + qy * v[7]);
  const float dot = vq[0] * w + vq[1] * qx + vq[2] * qy + vq[3] * qz;
  for (int i = 0; i < 4; ++i) g.q[i] = (vq[i] - st.qn[i] * dot) * st.inv_norm;
}

// ---------------------------------------------------------------------------------------------
// a1: cubic Hermite spline over n (<= P) control points; returns the 4 taps and their weights so
// that forward (sum w_i p_i) and backward (v_p_i += w_i g) share one definition.
// ---------------------------------------------------------------------------------------------
struct SplineTaps {
  int idx[4];     // control-point indices (may coincide)
  float w[4];     // d out / d p[idx[i]]
};

MOBGS_HD SplineTaps hermite_taps(float t, int n) {
  SplineTaps s;
  const float ts = t * (float)(n - 1);
  int i1 = (int)floorf(ts);
  i1 = i1 < 0 ? 0 : (i1 > n - 2 ? n - 2 : i1);
  int i0 = i1 - 1; i0 = i0 < 0 ? 0 : i0;           // <= n-1 always
  int i2 = i1 + 1; i2 = i2 > n - 1 ? n - 1 : i2;
  int i3 = i1 + 2; i3 = i3 > n - 1 ? n - 1 : i3;
  const float u = ts - (float)i1;
  const float om = 1.f - u;
  const float h00 = (1.f + 2.f * u) * om * om;
  const float h10 = u * om * om;
  const float h01 = u * u * (3.f - 2.f * u);
  const float h11 = u * u * (u - 1.f);
  float w0 = 0.f, w1 = h00, w2 = h01, w3 = 0.f;
  if (i0 == i1) { w2 += h10; w1 -= h10; } else { w2 += 0.5f * h10; w0 -= 0.5f * h10; }
  if (i3 == i2) { w2 += h11; w1 -= h11; } else { w3 += 0.5f * h11; w1 -= 0.5f * h11; }
  s.idx[0] = i0; s.idx[1] = i1; s.idx[2] = i2; s.idx[3] = i3;
  s.w[0] = w0; s.w[1] = w1; s.w[2] = w2; s.w[3] = w3;
  return s;
}

MOBGS_HD float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace mobgs
