"""Drop-in for the reference's `gaussian_renderer` package (gaussian_renderer/__init__.py):
same `render` / `get_flow` / `get_flow_static` signatures and return structure, so the
reference's train.py / eval.py / utils/scene_utils.py import it unchanged (put the directory
that contains this package's parent `mobgs_b200/` first on sys.path as `gaussian_renderer`, see
INTEGRATION.md), but one fused projection launch + at most three blend launches per call instead
of five independent gsplat pipelines.

Works with the reference's GaussianModel / Camera objects or the stand-ins of mobgs_b200.scene.
"""
from __future__ import annotations

import torch

from .. import fused, ops
from . import network_gui  # noqa: F401  (train.py does `from gaussian_renderer import render, network_gui`)

ED_ALPHA_FLOOR = 1e-10
TIGHT_TILES = True


def _static_params(pc):
    return (pc._xyz, pc._rotation, pc._scaling, pc._opacity, pc._features_dc)


def _dynamic_params(pc):
    return (pc.control_xyz, pc._rotation, pc._omega, pc._scaling, pc._opacity, pc._features_dc,
            pc._features_t, pc._trbf_center)


_BG_CACHE = {}


def _bg10(bg_color, dev):
    """bg[:3] tiled over the 9 feature channels + 0 for the depth channel (render():90-91), cached per
    background tensor (object identity + version) — it is the same tensor on every call of a training run."""
    hit = _BG_CACHE.get("bg")     # the cache holds the tensor itself, so its identity cannot be recycled
    if hit is not None and hit[0] is bg_color and hit[1] == bg_color._version and hit[2] == dev:
        return hit[3]
    b3 = bg_color[:3].detach().to(device=dev, dtype=torch.float32)
    out = torch.cat([b3, b3, b3, b3.new_zeros(1)])[None]
    _BG_CACHE["bg"] = (bg_color, bg_color._version, dev, out)
    return out


def _times(cam, delta_exposure, dev, clamp):
    """(t_spline, t_poly) as 0-d device tensors, built without a host sync."""
    t0 = float(cam.time)
    if delta_exposure is None:
        t = torch.tensor(t0, dtype=torch.float32, device=dev)
        return t, t
    if not torch.is_tensor(delta_exposure):      # host number: one 8-byte upload instead of a chain of tiny kernels
        tp = t0 + float(delta_exposure) / cam.max_time
        both = torch.tensor([min(max(tp, 0.0), 1.0) if clamp else tp, tp], dtype=torch.float32, device=dev)
        return both[0], both[1]
    t = t0 + delta_exposure.detach().to(device=dev, dtype=torch.float32) / cam.max_time
    return (torch.clamp(t, 0, 1) if clamp else t), t


def _decode(dyn_pc, img10, alpha, cam):
    """Sandwich decoder (dyn_pc.rgbdecoder's weights) + expected-depth division, one fp32 kernel.
    -> rgb [3,H,W], depth [1,H,W]"""
    dec = dyn_pc.rgbdecoder
    rgb, depth, _ = fused.decode(img10, alpha, cam.cam_ray, dec.mlp1.weight.reshape(6, 12),
                                 dec.mlp2.weight.reshape(3, 6))
    return rgb.squeeze(0), depth


def _alpha_render(alpha, bg_color):
    """rasterization(colors=1, backgrounds=bg[0:1]) channel 0 = (1 - T) + T * bg0."""
    return alpha + (1.0 - alpha) * bg_color[0].to(alpha)


def render(viewpoint_camera, stat_pc, dyn_pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0,
           override_color=None, stage="fine", cam_type=None, is_static=False, over_t=None, over_vde=None,
           get_static=False, get_dynamic=False, stat_stat=True, ref_wc=None, iter_fact=1, flow=None,
           coherent=None, target_ts=None, target_w2cs=None, get_heatmap=False, w2c=None,
           delta_exposure=None, get_flow=False, cluster=None):
    """Reference: gaussian_renderer/__init__.py:59-316 (same dict keys, :294-316)."""
    cam = viewpoint_camera
    dev = dyn_pc._scaling.device
    W, H = int(cam.image_width), int(cam.image_height)
    viewmat = cam.world_view_transform.transpose(0, 1) if w2c is None else w2c
    Kmat = cam.K
    Ns, Nd = stat_pc.get_xyz.shape[0], dyn_pc.get_xyz.shape[0]
    N = Ns + Nd
    warped = delta_exposure is not None
    t_spline, t_poly = _times(cam, delta_exposure, dev, clamp=True)
    bg10 = _bg10(bg_color, dev)

    records, radii, depths, means3d = fused.synth_project(
        _static_params(stat_pc), _dynamic_params(dyn_pc), dyn_pc.current_control_num,
        viewmat[None], Kmat[None], t_spline[None], t_poly[None], W, H, offset=coherent, want_means3d=True)
    # leaf carrying the projected means whose .grad receives d loss / d means2d of *this* render
    vsp = records[:, :, 0:2].detach().clone().requires_grad_(True)

    # one binning / blend / decode launch chain for every image this call returns: the combined
    # render plus (optionally) the dynamic-only and static-only renders are lists over the same
    # projected records (reference: five separate rasterization() pipelines, :143,:163,:201,:236,:255)
    specs = [(0, 0, N)]
    if get_dynamic:
        specs.append((0, Ns, N))
    if get_static:
        specs.append((0, 0, Ns))
    n_lists = len(specs)
    dec = dyn_pc.rgbdecoder
    rgb, depth_all, alpha, _ = fused.blend_decode(
        records, radii, depths, bg10.expand(n_lists, -1), cam.cam_ray, dec.mlp1.weight.reshape(6, 12),
        dec.mlp2.weight.reshape(3, 6), W, H, specs=specs, tight=TIGHT_TILES, vsp=vsp, vsp_k=0)
    rendered_image, depth = rgb[0], depth_all[0:1]
    radii0 = radii[0]

    d_image = d_depth = d_alpha = s_image = s_depth = s_alpha = None
    i = 1
    if get_dynamic:
        d_image, d_depth = rgb[i], depth_all[i:i + 1]
        d_alpha = _alpha_render(alpha[i:i + 1], bg_color)
        i += 1
    if get_static:
        s_depth = rendered_image[..., -1]   # reference quirk kept (:250)
        s_image = rgb[i]
        s_alpha = _alpha_render(alpha[i:i + 1], bg_color)

    rendered_flow = ori_coord_map = None
    if warped and get_flow:
        t0, _ = _times(cam, None, dev, clamp=False)
        ori_rec, _, _, _ = fused.synth_project(
            _static_params(stat_pc), _dynamic_params(dyn_pc), dyn_pc.current_control_num,
            viewmat[None], Kmat[None], t0[None], t0[None], W, H, offset=coherent)
        flow_2d = (ori_rec[..., 0:2] - records[..., 0:2].detach()).squeeze(0)
        rendered_flow, _ = ops.rasterize(records[..., 0:2], records[..., 3:6], flow_2d, records[0, :, 2],
                                         None, depths, radii, W, H, tight=TIGHT_TILES)
        ori_coord_map = _pixel_grid(cam, W, H, dev) + rendered_flow

    means3d0 = means3d[0]
    return {"render": rendered_image,
            "s_render": s_image,
            "s_depth": s_depth,
            "d_render": d_image,
            "d_depth": d_depth,
            "d_alpha": d_alpha,
            "d_means3d": means3d0[Ns:] if get_dynamic else None,
            "s_alpha": s_alpha,
            "viewspace_points": vsp,
            "visibility_filter": radii0 > 0,
            "radii": radii0,
            "depth": depth,
            "blending_factor": None,
            "world_coordinates": None,
            "splat_center": None,
            "means_3d_final": means3d0 * 1e2,
            "colors_precomp_final": records[0, :, 6:15],
            "ori_flow": rendered_flow,
            "ori_coord_map": ori_coord_map,
            "means_3d": means3d0[Ns:],
            "labels": None,
            "centroids": None}


def get_flow(viewpoint_camera, stat_pc, dyn_pc, pipe, bg_color: torch.Tensor, delta_exposure=None):
    """Reference: gaussian_renderer/__init__.py:318-492.  One projection launch (mid + exposure time)
    and two binning/blend launch chains feed all four rasterisations (the K = 1 case of
    get_flow_batched).  Returns (exp2mid_coord_map [1,H,W,2], mid2exp_coord_map [1,H,W,2],
    latent_img [3,H,W], latent_alpha [1,H,W])."""
    dev = dyn_pc._scaling.device
    de = torch.as_tensor(delta_exposure, dtype=torch.float32, device=dev).reshape(1)
    exp2mid, mid2exp, latent_img, latent_alpha = get_flow_batched(viewpoint_camera, stat_pc, dyn_pc, pipe,
                                                                  bg_color, de)
    return exp2mid, mid2exp, latent_img[0], latent_alpha


_GRID_CACHE = {}


def _pixel_grid(cam, W, H, dev):
    """cam.get_pixels(W, H, use_center=False) on the device.  It is the integer pixel grid [H,W,2] (x, y) for
    every camera (dycheck_geometry/camera.py:600-613: np.meshgrid of aranges), so it is built once per
    (W, H, device) on the device instead of a host meshgrid + a 16 MB upload per get_flow call."""
    key = (W, H, str(dev))
    g = _GRID_CACHE.get(key)
    if g is None:
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev),
                                torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
        g = torch.stack([xs, ys], dim=-1)
        if len(_GRID_CACHE) > 8:
            _GRID_CACHE.clear()
        _GRID_CACHE[key] = g
    return g


def get_flow_batched(viewpoint_camera, stat_pc, dyn_pc, pipe, bg_color: torch.Tensor, delta_exposures, rays=None):
    """K get_flow() calls (train.py:563-579 issues one per latent sub-frame) as ONE projection launch (K+1
    record sets: the mid time once + the K exposure times) and three launch chains that bin every distinct
    geometry once and walk it once per payload group:

      * K exposure-time lists: the latent image (:473, RGB+ED + decoder) and the exp2mid flow (:437) share
        geometry and order, so ONE walk composites the 10 image channels and the 2 flow channels (colour
        = mid - exp projected mean, taken from record set 0 on the fly; MobgsBlendFwd.flow_ref);
      * K dynamic-only lists of the same record sets for latent_alpha (:379) — alpha only, one channel;
      * the mid-time geometry ONCE: the K mid2exp flows (:456) differ only in their 2 colour channels, so all
        2K channels ride ceil(2K/10) walks of a single shared binning (fused.midflow_records).

    4K binned + walked lists before, 2K + 1 binned and 2K + ceil(2K/10) walked now (K of them over the dynamic
    Gaussians only).  Returns (exp2mid_coord [K,H,W,2], mid2exp_coord [K,H,W,2], latent_img [K,3,H,W],
    latent_alpha [K,H,W]) — `torch.cat` of what K reference calls return.
    `rays`: optional replacement for `viewpoint_camera.cam_ray` — a [1,6,H,W] tensor or a one-camera
    mobgs_b200.cameras.RayPose (rays generated inside the blend kernels, no ray image read)."""
    cam = viewpoint_camera
    dev = dyn_pc._scaling.device
    W, H = int(cam.image_width), int(cam.image_height)
    viewmat = cam.world_view_transform.transpose(0, 1)
    Ns, Nd = stat_pc.get_xyz.shape[0], dyn_pc.get_xyz.shape[0]
    N = Ns + Nd
    t0 = torch.as_tensor(float(cam.time), dtype=torch.float32, device=dev)
    de = torch.as_tensor(delta_exposures, dtype=torch.float32, device=dev).reshape(-1)
    K = de.numel()
    t_poly = torch.cat([t0[None], t0 + de / cam.max_time])
    rec, radii, depths, _ = fused.synth_project(
        _static_params(stat_pc), _dynamic_params(dyn_pc), dyn_pc.current_control_num,
        viewmat[None].expand(K + 1, -1, -1), cam.K[None].expand(K + 1, -1, -1), t_poly.clamp(0, 1), t_poly, W, H)
    grid = _pixel_grid(cam, W, H, dev)

    # one autograd node for all three render groups (fused._FlowRender): their backward passes accumulate into one
    # gradient-record buffer
    dec = dyn_pc.rgbdecoder
    rgb, flow_e2m, alpha_d, mid = fused.flow_render(
        rec, radii, depths, _bg10(bg_color, dev).expand(K, -1), cam.cam_ray if rays is None else rays,
        dec.mlp1.weight.reshape(6, 12), dec.mlp2.weight.reshape(3, 6), W, H, Ns, tight=TIGHT_TILES)
    exp2mid = grid + flow_e2m
    M = mid.shape[0]
    flows_m2e = mid.permute(1, 2, 0, 3).reshape(H, W, M * 10)[..., :2 * K].reshape(H, W, K, 2).permute(2, 0, 1, 3)
    mid2exp = grid + flows_m2e
    return exp2mid, mid2exp, rgb, _alpha_render(alpha_d, bg_color)


def get_flow_static(source_camera, target_camera, splat_camera, stat_pc, dyn_pc, pipe, bg_color: torch.Tensor):
    """Reference: gaussian_renderer/__init__.py:494-552."""
    means, quats = stat_pc.get_xyz, stat_pc._rotation
    scales, opac = stat_pc.get_scaling, stat_pc.get_opacity.squeeze(-1)
    W, H = int(source_camera.image_width), int(source_camera.image_height)
    Kmat = source_camera.K
    views = torch.stack([source_camera.world_view_transform.transpose(0, 1),
                         target_camera.world_view_transform.transpose(0, 1)])
    _, m2d, _, _ = ops.project(means, quats, scales, views, torch.stack([Kmat, Kmat]), W, H)
    flow_2d = m2d[0] - m2d[1]
    Ws, Hs = int(splat_camera.image_width), int(splat_camera.image_height)
    radii, sm2d, sdep, scon = ops.project(means, quats, scales,
                                          splat_camera.world_view_transform.transpose(0, 1)[None], Kmat[None], Ws, Hs)
    rendered_flow, _ = ops.rasterize(sm2d, scon, flow_2d, opac, None, sdep, radii, Ws, Hs, tight=TIGHT_TILES)
    return flow_2d, rendered_flow
