"""K-batched blurry-view render: the B200-first form of the loop the reference runs per training
view (train.py:441, :502-516, :540-541): K latent sub-frame `render()` calls followed by
`mean(stack(images)) + 1e-10`.  Here all K sub-frames share one launch of every kernel:

    synth+project (K cameras, K times)  ->  bin + per-tile sort (K*T segments)
    ->  blend (K*T CTAs) with the expected-depth + decoder epilogue fused in  ->  sub-frame mean

and the backward mirrors it, ending in one pass that writes every parameter gradient once.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

import os

from . import fused
from .gaussian_renderer import _bg10, _dynamic_params, _static_params
from .ops import BinPlan

# ablation only (tools/ablation.sh): MOBGS_ABL_AABB_TILES=1 lists every tile of gsplat's 3-sigma
# square instead of pruning the provably empty ones
_DEFAULT_TIGHT = os.environ.get("MOBGS_ABL_AABB_TILES", "0") != "1"


def render_subframes(stat_pc, dyn_pc, viewmats: torch.Tensor, Ks: torch.Tensor, t_spline: torch.Tensor,
                     t_poly: torch.Tensor, rays: torch.Tensor, bg_color: torch.Tensor, width: int,
                     height: int, center_k: Optional[int] = None, tight: Optional[bool] = None,
                     offset: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """viewmats [K,4,4] (may require grad: they come from the BLCE pose network), Ks [K,3,3] or
    [3,3], t_spline / t_poly [K] device tensors (see gaussian_renderer._times), rays [K,6,H,W]
    (Camera.cam_ray of each warped camera) or [1,6,H,W] — or a mobgs_b200.cameras.RayPose (12 pose floats per camera;
    rays are then generated inside the blend kernels and no ray image exists).

    Returns: "render" [3,H,W] (the blurred prediction, mean over K + 1e-10), "subframes"
    [K,3,H,W], "depth" [K,H,W], "alpha" [K,H,W], "radii" [K,N], "viewspace_points" [1,N,2] (leaf whose
    .grad receives d loss / d means2d of sub-frame `center_k`, default K//2)."""
    K = viewmats.shape[0]
    dev = viewmats.device
    tight = _DEFAULT_TIGHT if tight is None else tight
    if Ks.dim() == 2:
        Ks = Ks[None].expand(K, -1, -1)
    ck = K // 2 if center_k is None else center_k
    N = stat_pc.get_xyz.shape[0] + dyn_pc.get_xyz.shape[0]
    plan = BinPlan([(k, 0, N) for k in range(K)], width, height, tight)      # the K lists are counted by the projection launch
    records, radii, depths, _ = fused.synth_project(
        _static_params(stat_pc), _dynamic_params(dyn_pc), dyn_pc.current_control_num,
        viewmats, Ks, t_spline, t_poly, width, height, offset=offset, binning=plan)
    vsp = records[ck:ck + 1, :, 0:2].detach().clone().requires_grad_(True)
    bg10 = _bg10(bg_color, dev).expand(K, -1)
    dec = dyn_pc.rgbdecoder
    rgb, depth, alpha, mean = fused.blend_decode(
        records, radii, depths, bg10, rays, dec.mlp1.weight.reshape(6, 12), dec.mlp2.weight.reshape(3, 6),
        width, height, tight=tight, vsp=vsp, vsp_k=ck, want_mean=True, plan=plan)
    return {"render": mean, "subframes": rgb, "depth": depth, "alpha": alpha, "radii": radii,
            "viewspace_points": vsp, "visibility_filter": radii[ck] > 0}


def render_blurry_view(viewpoint_cam, warped_cams, exposure_time, stat_pc, dyn_pc, pipe, bg_color,
                       use_delta_exposure: bool = True, tight: bool = True,
                       rays: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Fused equivalent of the reference's per-view render loop (train.py:441 + :497-541):

        render_pkg = render(viewpoint_cam, ..., get_static=True, get_dynamic=True)        # centre
        for k, cam in enumerate(warped_cams):                                             # K = num_warp
            images[k] = image_ori if k == K//2 else render(cam, ..., delta_exposure=exposure_time[k])["render"]
        pred_image = mean(stack(images)) + 1e-10

    as ONE projection launch (K record sets) and ONE binning / blend+decode launch chain over K + 2
    lists: the K sub-frames, plus the dynamic-only and static-only renders of the centre sub-frame
    (the only ones train.py reads s_render / d_alpha / s_alpha / d_depth from).  The 4 x (K-1)
    static/dynamic renders of the warped sub-frames that the reference computes and drops
     are simply not rendered.

    `viewpoint_cam` / `warped_cams[k]` are reference `Camera` objects (or mobgs_b200.scene
    stand-ins); `exposure_time` is blceKernel.get_warped_cams' second output ([K] tensor);
    use_delta_exposure mirrors `iteration > blceopt.start_warp_dynamic` (train.py:503-506).
    `rays`: optional pre-stacked [K,6,H,W] camera rays of `warped_cams` (centre camera at K//2), or their
    mobgs_b200.cameras.RayPose (get_warped_cams_batched(...).ray_pose): rays generated inside the blend kernels.

    Returns the keys train.py consumes: "render" (blurred prediction), "subframes" [K,3,H,W],
    "depths" [K,H,W], and the centre render's "render_center", "depth", "s_render", "s_depth",
    "s_alpha", "d_render", "d_depth", "d_alpha", "viewspace_points", "visibility_filter", "radii"."""
    K = len(warped_cams)
    half = K // 2
    dev = dyn_pc._scaling.device
    W, H = int(viewpoint_cam.image_width), int(viewpoint_cam.image_height)
    Ns, Nd = stat_pc.get_xyz.shape[0], dyn_pc.get_xyz.shape[0]
    N = Ns + Nd
    cams = [viewpoint_cam if k == half else warped_cams[k] for k in range(K)]
    viewmats = torch.stack([c.world_view_transform.transpose(0, 1) for c in cams])
    Ks = torch.stack([viewpoint_cam.K] * K)
    t0 = torch.as_tensor(float(viewpoint_cam.time), dtype=torch.float32, device=dev)
    et = torch.as_tensor(exposure_time, dtype=torch.float32, device=dev).detach()
    if not use_delta_exposure:
        et = torch.zeros_like(et)
    t_poly = t0 + et / viewpoint_cam.max_time
    is_center = torch.arange(K, device=dev) == half
    t_poly = torch.where(is_center, t0, t_poly)
    t_spline = torch.where(is_center, t0, t_poly.clamp(0, 1))      # the centre render does not clamp (render():114)
    if rays is None:       # `rays` [K,6,H,W]: the K cameras' cam_ray already stacked (mobgs_b200.cameras.camera_rays builds them
        rays = torch.cat([c.cam_ray for c in cams]).to(dev)     # in one launch), saving the concatenation and its backward

    specs = [(k, 0, N) for k in range(K)] + [(half, Ns, N), (half, 0, Ns)]
    plan = BinPlan(specs, W, H, tight)          # all K + 2 lists are counted by the projection launch
    records, radii, depths, _ = fused.synth_project(
        _static_params(stat_pc), _dynamic_params(dyn_pc), dyn_pc.current_control_num,
        viewmats, Ks, t_spline, t_poly, W, H, binning=plan)
    vsp = records[half:half + 1, :, 0:2].detach().clone().requires_grad_(True)
    bg10 = _bg10(bg_color, dev).expand(K + 2, -1)
    dec = dyn_pc.rgbdecoder
    rgb, depth, alpha, mean = fused.blend_decode(
        records, radii, depths, bg10, rays, dec.mlp1.weight.reshape(6, 12), dec.mlp2.weight.reshape(3, 6),
        W, H, specs=specs, tight=tight, vsp=vsp, vsp_k=half, want_mean=True, mean_K=K, plan=plan)
    bg0 = bg_color[0].to(alpha)
    center = rgb[half]
    return {"render": mean, "subframes": rgb[:K], "depths": depth[:K], "render_center": center,
            "depth": depth[half:half + 1], "d_render": rgb[K], "d_depth": depth[K:K + 1],
            "d_alpha": alpha[K:K + 1] + (1.0 - alpha[K:K + 1]) * bg0, "s_render": rgb[K + 1],
            "s_depth": center[..., -1], "s_alpha": alpha[K + 1:K + 2] + (1.0 - alpha[K + 1:K + 2]) * bg0,
            "viewspace_points": vsp, "visibility_filter": radii[half] > 0, "radii": radii[half]}
