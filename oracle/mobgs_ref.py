"""ORACLE — test infrastructure only (see oracle/gsplat_ref.py header; parity unpinned
for the gsplat half, pinned against the reference's own Python for this half by
tests/test_oracle.py + tests/golden/ fixtures generated with
tests/golden/make_golden.py from /root/reference).

CPU restatement of the MoBGS renderer layer that sits on top of the two gsplat
operators:

  * interpolate_cubic_hermite            gaussian_renderer/__init__.py:23-56
  * GaussianModel getters / activations  scene/gaussian_model.py:98-106, 209-257
  * Sandwich RGB decoder                 helper_model.py:7-28
  * render / get_flow / get_flow_static  gaussian_renderer/__init__.py:59-316, 318-492, 494-552

The model / camera arguments only need the attribute API listed in SURVEY.md §8b
(the reference's GaussianModel / Camera satisfy it, and so do the light stand-ins
in mobgs_b200/scene.py).  Runs on whatever device the parameters live on.
"""
from __future__ import annotations

import functools

import torch

from . import gsplat_ref as G


# ----------------------------------------------------------------------------
# a1: cubic Hermite spline over a per-Gaussian number of control points
# ----------------------------------------------------------------------------
def hermite_spline(control_xyz: torch.Tensor, t: torch.Tensor, n_ctrl: torch.Tensor) -> torch.Tensor:
    """control_xyz [Nd,P,3]; t scalar tensor (already clamped by the caller where the
    reference clamps); n_ctrl [Nd,1] int64 in [2..P].  Returns [Nd,3] (not yet x1e-2).

    Follows gaussian_renderer/__init__.py:23-56: segment i = clamp(floor(t(n-1)), 0, n-2),
    neighbours clamped into [0, n-1], one-sided tangents where a neighbour collapses."""
    n = n_ctrl.reshape(-1).to(torch.int64)
    ts = t.to(control_xyz.dtype) * (n - 1).to(control_xyz.dtype)          # [Nd]
    zero = torch.zeros_like(n)
    i1 = torch.minimum(torch.maximum(torch.floor(ts).to(torch.int64), zero), n - 2)
    i0 = torch.minimum(torch.maximum(i1 - 1, zero), n - 1)
    i2 = torch.minimum(torch.maximum(i1 + 1, zero), n - 1)
    i3 = torch.minimum(torch.maximum(i1 + 2, zero), n - 1)
    u = (ts - i1.to(ts.dtype))[:, None]

    def pick(idx):
        return torch.gather(control_xyz, 1, idx[:, None, None].expand(-1, 1, 3)).squeeze(1)

    p0, p1, p2, p3 = pick(i0), pick(i1), pick(i2), pick(i3)
    m0 = torch.where((i0 == i1)[:, None], p2 - p1, (p2 - p0) / 2)
    m1 = torch.where((i3 == i2)[:, None], p2 - p1, (p3 - p1) / 2)
    h00 = (1 + 2 * u) * (1 - u) ** 2
    h10 = u * (1 - u) ** 2
    h01 = u ** 2 * (3 - 2 * u)
    h11 = u ** 2 * (u - 1)
    return h00 * p1 + h10 * m0 + h01 * p2 + h11 * m1


# ----------------------------------------------------------------------------
# a8: Sandwich decoder  (helper_model.py:19-28), bias-free 1x1 convs
# ----------------------------------------------------------------------------
def sandwich(img9: torch.Tensor, rays6: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor) -> torch.Tensor:
    """img9 [B,9,H,W] = albedo|spec|timefeature; rays6 [B,6,H,W]; w1 [6,12]; w2 [3,6]."""
    albedo, spec, tfeat = img9[:, 0:3], img9[:, 3:6], img9[:, 6:9]
    x = torch.cat([spec, tfeat, rays6], dim=1)
    h = torch.relu(torch.einsum("oc,bchw->bohw", w1, x))
    return torch.sigmoid(albedo + torch.einsum("oc,bchw->bohw", w2, h))


def _decoder_weights(dyn_pc):
    dec = dyn_pc.rgbdecoder
    return dec.mlp1.weight.reshape(6, 12), dec.mlp2.weight.reshape(3, 6)


# ----------------------------------------------------------------------------
# a2: per-time attribute synthesis
# ----------------------------------------------------------------------------
def _time_scalar(cam, like: torch.Tensor) -> torch.Tensor:
    return torch.as_tensor(cam.time, dtype=like.dtype, device=like.device)


def dynamic_attributes(dyn_pc, t_cam: torch.Tensor, clamp_time: bool):
    """Returns (means[Nd,3], quats[Nd,4] normalised, scales, opac[Nd,1], colors[Nd,9])
    at camera time t_cam (a 0-d tensor).  render():93-125."""
    dt_poly = (t_cam - dyn_pc.get_trbfcenter).detach()                     # [Nd,1]
    quats = dyn_pc.rotation_activation(dyn_pc.get_rotation_dy(dyn_pc._rotation, dt_poly))
    t_spline = torch.clamp(t_cam, 0, 1) if clamp_time else t_cam
    means = hermite_spline(dyn_pc.get_control_xyz, t_spline, dyn_pc.current_control_num) * 1e-2
    scales = dyn_pc.scaling_activation(dyn_pc._scaling)
    colors = dyn_pc.get_features(dt_poly)
    return means, quats, scales, dyn_pc.get_opacity, colors


def static_attributes(stat_pc):
    return (stat_pc.get_xyz, stat_pc.get_rotation_stat, stat_pc.get_scaling,
            stat_pc.get_opacity, stat_pc.get_features_static)


def _raster(means, quats, scales, opac, colors, bg, viewmat, K, cam, mode, window=None):
    return G.rasterization(
        means=means, quats=quats, scales=scales, opacities=opac.squeeze(-1), colors=colors,
        backgrounds=bg, viewmats=viewmat[None], Ks=K[None],
        width=int(cam.image_width), height=int(cam.image_height), packed=False, render_mode=mode,
        window=window)


def _project(means, quats, scales, viewmat, K, cam):
    return G.fully_fused_projection(
        means=means, covars=None, quats=quats, scales=scales, viewmats=viewmat[None], Ks=K[None],
        width=int(cam.image_width), height=int(cam.image_height))[1]


def _pixel_grid(cam, like):
    W, H = int(cam.image_width), int(cam.image_height)
    return torch.as_tensor(cam.get_pixels(W, H, use_center=False)).type_as(like)


# ----------------------------------------------------------------------------
# a3: render()
# ----------------------------------------------------------------------------
def render_ref(viewpoint_camera, stat_pc, dyn_pc, pipe, bg_color, scaling_modifier=1.0,
               override_color=None, stage="fine", cam_type=None, is_static=False, over_t=None,
               over_vde=None, get_static=False, get_dynamic=False, stat_stat=True, ref_wc=None,
               iter_fact=1, flow=None, coherent=None, target_ts=None, target_w2cs=None,
               get_heatmap=False, w2c=None, delta_exposure=None, get_flow=False, cluster=None,
               window=None):
    """`window` = (x0, y0, w, h), tile aligned (oracle-only extension): every image in the result is
    the crop of the full-frame render to that window (all Gaussians are projected, only the window is
    rasterised) — how the full-size parity tests and bench.py's CPU arm sample a 1 M-Gaussian frame."""
    cam = viewpoint_camera
    viewmat = cam.world_view_transform.transpose(0, 1) if w2c is None else w2c
    K = cam.K
    bg9 = torch.cat([bg_color[:3]] * 3, dim=-1)
    like = dyn_pc._scaling
    t0 = _time_scalar(cam, like)
    warped = delta_exposure is not None
    t_cam = t0 + delta_exposure / cam.max_time if warped else t0

    d_means, d_quats, d_scales, d_opac, d_cols = dynamic_attributes(dyn_pc, t_cam, clamp_time=warped)
    if coherent is not None:
        d_means = d_means + coherent
    s_means, s_quats, s_scales, s_opac, s_cols = static_attributes(stat_pc)
    w1, w2 = _decoder_weights(dyn_pc)

    rays = cam.cam_ray
    if window is not None:
        rays = rays[..., window[1]:window[1] + window[3], window[0]:window[0] + window[2]]
    rast = functools.partial(_raster, window=window)

    def decode(img10):
        return sandwich(img10[..., :-1].permute(0, 3, 1, 2), rays, w1, w2).squeeze(0)

    out = {k: None for k in ("s_render", "s_depth", "d_render", "d_depth", "d_alpha", "d_means3d",
                             "s_alpha", "blending_factor", "world_coordinates", "splat_center",
                             "ori_flow", "ori_coord_map", "labels", "centroids")}
    if get_dynamic:
        d_img, _, _ = rast(d_means, d_quats, d_scales, d_opac, d_cols, bg9[None], viewmat, K, cam, "RGB+ED")
        out["d_depth"] = d_img[..., -1]
        out["d_render"] = decode(d_img)
        ones = torch.ones(d_cols.shape[0], 1, dtype=like.dtype, device=like.device)
        d_alpha, _, _ = rast(d_means, d_quats, d_scales, d_opac, ones, bg9[0:1][None], viewmat, K, cam, "RGB")
        out["d_alpha"] = d_alpha[..., 0]
        out["d_means3d"] = d_means

    means = torch.cat([s_means, d_means], 0)
    quats = torch.cat([s_quats, d_quats], 0)
    scales = torch.cat([s_scales, d_scales], 0)
    opac = torch.cat([s_opac, d_opac], 0)
    cols = torch.cat([s_cols, d_cols], 0)

    want_flow = warped and get_flow
    if want_flow:
        o_means, o_quats, _, _, _ = dynamic_attributes(dyn_pc, t0, clamp_time=False)
        ori_m2d = _project(torch.cat([s_means, o_means], 0), torch.cat([s_quats, o_quats], 0),
                           scales, viewmat, K, cam)

    img, _, info = rast(means, quats, scales, opac, cols, bg9[None], viewmat, K, cam, "RGB+ED")
    depth = img[..., -1]
    radii = info["radii"].squeeze(0)
    if info["means2d"].requires_grad:
        info["means2d"].retain_grad()
    rendered = decode(img)

    if get_static:
        s_img, _, _ = rast(s_means, s_quats, s_scales, s_opac, s_cols, bg9[None], viewmat, K, cam, "RGB+ED")
        # reference quirk kept (renderer :250): s_depth is the last *column* of the decoded RGB
        out["s_depth"] = rendered[..., -1]
        out["s_render"] = decode(s_img)
        ones = torch.ones(s_cols.shape[0], 1, dtype=like.dtype, device=like.device)
        s_alpha, _, _ = rast(s_means, s_quats, s_scales, s_opac, ones, bg9[0:1][None], viewmat, K, cam, "RGB")
        out["s_alpha"] = s_alpha[..., 0]

    if want_flow:
        flow_2d = (ori_m2d - info["means2d"].clone().detach()).squeeze(0)
        rendered_flow, _, _ = rast(means, quats, scales, opac, flow_2d, None, viewmat, K, cam, "RGB")
        out["ori_flow"] = rendered_flow
        grid = _pixel_grid(cam, rendered_flow)
        if window is not None:
            grid = grid[window[1]:window[1] + window[3], window[0]:window[0] + window[2]]
        out["ori_coord_map"] = grid + rendered_flow

    out.update({
        "render": rendered, "viewspace_points": info["means2d"], "visibility_filter": radii > 0,
        "radii": radii, "depth": depth, "means_3d_final": means * 1e2,
        "colors_precomp_final": cols, "means_3d": d_means,
    })
    return out


# ----------------------------------------------------------------------------
# a3: get_flow()   gaussian_renderer/__init__.py:318-492
# ----------------------------------------------------------------------------
def get_flow_ref(viewpoint_camera, stat_pc, dyn_pc, pipe, bg_color, delta_exposure=None):
    cam = viewpoint_camera
    viewmat = cam.world_view_transform.transpose(0, 1)
    K = cam.K
    bg9 = torch.cat([bg_color[:3]] * 3, dim=-1)
    like = dyn_pc._scaling
    t0 = _time_scalar(cam, like)
    t_exp = t0 + delta_exposure / cam.max_time

    m_means, m_quats, d_scales, d_opac, _ = dynamic_attributes(dyn_pc, t0, clamp_time=True)
    e_means, e_quats, _, _, e_cols = dynamic_attributes(dyn_pc, t_exp, clamp_time=True)
    s_means, s_quats, s_scales, s_opac, s_cols = static_attributes(stat_pc)

    ones = torch.ones(e_cols.shape[0], 1, dtype=like.dtype, device=like.device)
    latent_alpha, _, _ = _raster(e_means, e_quats, d_scales, d_opac, ones, bg9[0:1][None], viewmat, K, cam, "RGB")
    latent_alpha = latent_alpha[..., 0]

    mid_means = torch.cat([s_means, m_means], 0)
    mid_quats = torch.cat([s_quats, m_quats], 0)
    exp_means = torch.cat([s_means, e_means], 0)
    exp_quats = torch.cat([s_quats, e_quats], 0)
    scales = torch.cat([s_scales, d_scales], 0)
    opac = torch.cat([s_opac, d_opac], 0)
    exp_cols = torch.cat([s_cols, e_cols], 0)

    mid_m2d = _project(mid_means, mid_quats, scales, viewmat, K, cam)
    exp_m2d = _project(exp_means, exp_quats, scales, viewmat, K, cam)
    e2m = (mid_m2d - exp_m2d).squeeze(0)
    e2m_flow, _, _ = _raster(exp_means, exp_quats, scales, opac, e2m, None, viewmat, K, cam, "RGB")
    grid = _pixel_grid(cam, e2m_flow)
    exp2mid = grid + e2m_flow
    m2e_flow, _, _ = _raster(mid_means, mid_quats, scales, opac, -e2m, None, viewmat, K, cam, "RGB")
    mid2exp = grid + m2e_flow

    img, _, _ = _raster(exp_means, exp_quats, scales, opac, exp_cols, bg9[None], viewmat, K, cam, "RGB+ED")
    w1, w2 = _decoder_weights(dyn_pc)
    latent_img = sandwich(img[..., :-1].permute(0, 3, 1, 2), cam.cam_ray, w1, w2).squeeze(0)
    return exp2mid, mid2exp, latent_img, latent_alpha


# ----------------------------------------------------------------------------
# a3: get_flow_static()   gaussian_renderer/__init__.py:494-552
# ----------------------------------------------------------------------------
def get_flow_static_ref(source_camera, target_camera, splat_camera, stat_pc, dyn_pc, pipe, bg_color):
    s_means, s_quats, s_scales, s_opac, _ = static_attributes(stat_pc)
    K = source_camera.K
    src_m2d = _project(s_means, s_quats, s_scales, source_camera.world_view_transform.transpose(0, 1), K, source_camera)
    tgt_m2d = _project(s_means, s_quats, s_scales, target_camera.world_view_transform.transpose(0, 1), K, target_camera)
    flow_2d = (src_m2d - tgt_m2d).squeeze(0)
    rendered_flow, _, _ = _raster(s_means, s_quats, s_scales, s_opac, flow_2d, None,
                                  splat_camera.world_view_transform.transpose(0, 1), K, splat_camera, "RGB")
    return flow_2d, rendered_flow


# ----------------------------------------------------------------------------
# a9: the blur model — pixel mean of the K decoded sub-frames (train.py:540-541)
# ----------------------------------------------------------------------------
def blur_mean(images):
    return torch.mean(torch.stack(list(images), dim=0), dim=0) + 1e-10


def camera_rays_ref(rot, centre, ppx, ppy, sfx, sfy, width, height):
    """Camera.cam_ray for K cameras: scene/cameras.py:140-146 with get_pixels_torch :244-253,
    pixels_to_local_viewdirs_torch :255-266 and pixels_to_viewdirs_torch :268-284.
    rot [K,3,3] = Camera.R (camera-to-world), centre [K,3] -> [K,6,H,W].
    Pinned by tests/golden/camera_rays.npz (the reference's own methods, tests/golden/make_golden.py)."""
    xx, yy = torch.meshgrid(torch.arange(width, dtype=torch.float32), torch.arange(height, dtype=torch.float32), indexing="xy")
    pixels = torch.stack([xx, yy], dim=-1) + 0.5
    y = (pixels[..., 1] - ppy) / sfy
    x = (pixels[..., 0] - ppx) / sfx
    local = torch.stack([x, y, torch.ones_like(x)], dim=-1)
    local = (local / torch.norm(local, dim=-1, keepdim=True)).to(rot.dtype)
    out = []
    for k in range(rot.shape[0]):
        v = torch.matmul(rot[k], local.reshape(-1, 3)[..., None])[..., 0]
        v = (v / torch.norm(v, dim=-1, keepdim=True)).view(height, width, 3)
        origin, _ = torch.broadcast_tensors(centre[k], v)
        out.append(torch.cat((origin, v), dim=-1).permute(2, 0, 1))
    return torch.stack(out)
