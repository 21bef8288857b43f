"""Camera rays for K cameras in one launch.

`camera_rays` computes what scene/cameras.py:132-146 stores as `Camera.cam_ray` — for the K warped
sub-frame cameras of a blurry view at once (the reference rebuilds an H x W meshgrid, the local view
directions and a batched 3x3 matmul inside every `Camera(...)` constructor, K times per view per step,
scene/blce.py:139-159).  Differentiable w.r.t. the camera-to-world rotation and the camera centre
(the pose gradient that trains the BLCE network).  CUDA fp32 only, no fallback.
"""
from __future__ import annotations

import torch

from . import _lib


class _CameraRays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rot, centre, ppx, ppy, sfx, sfy, W, H):
        if not (rot.is_cuda and centre.is_cuda and rot.dtype == torch.float32 and centre.dtype == torch.float32):
            raise RuntimeError("camera_rays needs CUDA fp32 tensors (there is no CPU fallback)")
        rot_c, cen_c = rot.contiguous(), centre.contiguous()
        K = rot_c.shape[0]
        rays = torch.empty(K, 6, H, W, device=rot.device)
        a = _lib.CameraRays()
        a.K, a.H, a.W = K, H, W
        a.ppx, a.ppy, a.sfx, a.sfy = ppx, ppy, sfx, sfy
        a.rot, a.centre, a.rays = rot_c.data_ptr(), cen_c.data_ptr(), rays.data_ptr()
        _lib.call("mobgs_camera_rays_fwd", a, _lib.current_stream())
        ctx.save_for_backward(rot_c, cen_c)
        ctx.geom = (ppx, ppy, sfx, sfy, W, H)
        return rays

    @staticmethod
    def backward(ctx, g):
        rot_c, cen_c = ctx.saved_tensors
        ppx, ppy, sfx, sfy, W, H = ctx.geom
        g = g.contiguous()
        K = rot_c.shape[0]
        v_rot, v_cen = torch.empty_like(rot_c), torch.empty_like(cen_c)
        a = _lib.CameraRays()
        a.K, a.H, a.W = K, H, W
        a.ppx, a.ppy, a.sfx, a.sfy = ppx, ppy, sfx, sfy
        a.rot, a.centre = rot_c.data_ptr(), cen_c.data_ptr()
        a.v_rays, a.v_rot, a.v_centre = g.data_ptr(), v_rot.data_ptr(), v_cen.data_ptr()
        _lib.call("mobgs_camera_rays_bwd", a, _lib.current_stream())
        return v_rot, v_cen, None, None, None, None, None, None


def camera_rays(rot: torch.Tensor, centre: torch.Tensor, ppx: float, ppy: float, sfx: float, sfy: float,
                width: int, height: int) -> torch.Tensor:
    """rot [K,3,3] camera-to-world rotations (Camera.R), centre [K,3] camera centres ->
    cam_ray [K,6,H,W] = [centre | normalised world-space view direction of every pixel centre]."""
    return _CameraRays.apply(rot, centre, float(ppx), float(ppy), float(sfx), float(sfy), int(width), int(height))


def camera_rays_from_w2c(viewmats: torch.Tensor, fx: float, fy: float, cx: float, cy: float,
                         width: int, height: int) -> torch.Tensor:
    """Same from K world-to-camera matrices [K,4,4] (a pinhole camera: ppx = cx, sfx = fx, ...)."""
    c2w = torch.inverse(viewmats)
    return camera_rays(c2w[:, :3, :3], c2w[:, :3, 3], cx, cy, fx, fy, width, height)
