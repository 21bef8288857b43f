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
                         width: int, height: int, rigid: bool = False) -> torch.Tensor:
    """Same from K world-to-camera matrices [K,4,4] (a pinhole camera: ppx = cx, sfx = fx, ...); rigid: see
    ray_pose_from_w2c."""
    if rigid:
        rot, centre = rigid_c2w(viewmats)
        return camera_rays(rot, centre, cx, cy, fx, fy, width, height)
    c2w = torch.inverse(viewmats)
    return camera_rays(c2w[:, :3, :3], c2w[:, :3, 3], cx, cy, fx, fy, width, height)


class RayPose:
    """The K cameras of `camera_rays` WITHOUT the ray images: `pose` [K,12] = (camera-to-world rotation, row-major 9 |
    camera centre 3) per camera — an ordinary differentiable tensor — plus the pixel-grid constants.  Accepted wherever
    the fused path takes `rays` (fused.blend_decode, subframes.render_subframes / render_blurry_view): the blend
    kernels then evaluate each pixel's ray in registers (MobgsBlendFwd.dec_pose) and reduce the gradient of the 12
    pose floats in the backward prologue (MobgsBlendBwd.v_pose_partial), so the [K,6,H,W] tensor that
    scene/cameras.py:132-146 materialises per warped camera (348 MB per blurry view at 1080p, K = 7; written once,
    read by forward and backward, and its gradient image written and reduced again) never exists."""

    def __init__(self, pose: torch.Tensor, ppx: float, ppy: float, sfx: float, sfy: float):
        if pose.dim() != 2 or pose.shape[1] != 12:
            raise ValueError("pose must be [K,12] (rotation row-major | centre)")
        self.pose = pose
        self.intr = (float(ppx), float(ppy), float(sfx), float(sfy))

    @property
    def shape(self):                     # leading axis = number of cameras, like a rays tensor
        return self.pose.shape

    def __getitem__(self, idx):          # slicing cameras (sub-frame sharding)
        p = self.pose[idx]
        return RayPose(p if p.dim() == 2 else p[None], *self.intr)

    def rays(self, width: int, height: int) -> torch.Tensor:
        """materialise the [K,6,H,W] ray images (tests / the operator path)"""
        return camera_rays(self.pose[:, :9].reshape(-1, 3, 3), self.pose[:, 9:], *self.intr, width, height)


def ray_pose(rot: torch.Tensor, centre: torch.Tensor, ppx: float, ppy: float, sfx: float, sfy: float) -> RayPose:
    """rot [K,3,3] camera-to-world rotations, centre [K,3] -> RayPose (see there)."""
    return RayPose(torch.cat([rot.reshape(rot.shape[0], 9), centre.reshape(centre.shape[0], 3)], dim=1), ppx, ppy, sfx, sfy)


def rigid_c2w(viewmats: torch.Tensor):
    """(rotation [K,3,3], centre [K,3]) of the camera-to-world transforms of K RIGID world-to-camera matrices [K,4,4]:
    R_c2w = R^T, c = -R^T t — closed form, differentiable, and without the device->host `info` check (a stream
    synchronisation) that `torch.inverse` performs."""
    rot = viewmats[:, :3, :3].transpose(1, 2)
    centre = -(rot @ viewmats[:, :3, 3:4]).squeeze(-1)
    return rot, centre


def ray_pose_from_w2c(viewmats: torch.Tensor, fx: float, fy: float, cx: float, cy: float, rigid: bool = False) -> RayPose:
    """Same from K world-to-camera matrices [K,4,4] (a pinhole camera: ppx = cx, sfx = fx, ...).  rigid=True: the
    matrices are rigid transforms, inverted in closed form (rigid_c2w) instead of by torch.inverse."""
    if rigid:
        rot, centre = rigid_c2w(viewmats)
    else:
        c2w = torch.inverse(viewmats)
        rot, centre = c2w[:, :3, :3], c2w[:, :3, 3]
    return ray_pose(rot, centre, cx, cy, fx, fy)
