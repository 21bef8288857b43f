"""The K warped sub-frame cameras of a blurry view in one batch.

Reference: `blceKernel.get_warped_cams` (scene/blce.py:139-159) runs the BLCE pose network, inverts the K
camera-to-world matrices and then constructs K full `Camera` objects (scene/cameras.py:18-151), each of which
re-derives `world_view_transform`, calls `torch.inverse` for its camera centre and rebuilds the H x W pixel grid,
the local view directions and a batched 3x3 matmul for `cam_ray` (:132-146) — K x (~15 launches + one inverse) per
view per step, all of it on the critical path between the centre render and the K sub-frame renders.

`get_warped_cams_batched` keeps the network call (the reference's own module object, out of scope, tiny) and
replaces everything after it by: one batched inverse for the K world-to-camera matrices, one for the K camera
centres, and ONE `mobgs_camera_rays_fwd` launch for all K ray images (differentiable w.r.t. rotation and
centre, so the pose gradient still reaches the BLCE network).  The result is a list of light camera objects with
the attributes train.py and the renderer read from a warped camera (`R`, `T`, `world_view_transform`,
`camera_center`, `cam_ray`, `K`, `time`, `max_time`, `image_width`, `image_height`, `uid`, `metadata`,
`get_pixels`) plus the stacked tensors (`.viewmats [K,4,4]`, `.ray_pose` = 12 pose floats per camera, `.rays [K,6,H,W]` built only on
demand) that `mobgs_b200.subframes.render_blurry_view(..., rays=cams.ray_pose)` consumes: the blend kernels generate
every pixel's ray in registers, so on the fused path no ray image is ever written or read.
"""
from __future__ import annotations

import sys
from typing import Callable, Optional

import numpy as np
import torch


class WarpedCamera:
    """Attribute view of one warped sub-frame camera (what scene/cameras.py:Camera exposes to train.py:497-530 and
    to gaussian_renderer.render)."""

    def __init__(self, base, R, T, wvt, centre, batch, k):
        self._base = base
        self.R, self.T = R, T                          # camera-to-world rotation, world-to-camera translation (torch)
        self.world_view_transform = wvt                # [4,4], transposed world-to-camera (cameras.py:126)
        self.camera_center = centre
        self._batch, self._k = batch, k

    @property
    def cam_ray(self):                                 # [1,6,H,W]; materialised (for all K cameras, one launch) on first use
        return self._batch.rays[self._k:self._k + 1]

    def __getattr__(self, name):                       # everything else (K, time, max_time, image sizes, uid, metadata,
        return getattr(self._base, name)               # image, depth, mask, get_pixels, ...) is the dataset camera's


class WarpedCameras(list):
    """list of WarpedCamera + the stacked tensors of the batch: `viewmats` [K,4,4] world-to-camera, `ray_pose`
    (mobgs_b200.cameras.RayPose: 12 pose floats per camera — what render_blurry_view(..., rays=cams.ray_pose) consumes;
    no ray image is built) and `rays` [K,6,H,W], built lazily by one camera_rays launch only if somebody reads it or a
    camera's `cam_ray` (the drop-in render() does)."""
    viewmats: torch.Tensor
    _rays = None
    _rays_fn = None

    @property
    def rays(self):
        if self._rays is None:
            self._rays = self._rays_fn()
        return self._rays


def get_warped_cams_batched(kernel, cam, fwd_cam=None, bwd_cam=None, rays_fn: Optional[Callable] = None):
    """Drop-in for `blcekernel.get_warped_cams(cam, fwd_cam, bwd_cam)` (train.py:472): -> (cams, exposure_time).
    `kernel` is the reference's blceKernel instance.  rays_fn(rot[K,3,3], centre[K,3], ppx, ppy, sfx, sfy, W, H) ->
    [K,6,H,W]; default = mobgs_b200.cameras.camera_rays (the CUDA kernel)."""
    if rays_fn is None:
        from .cameras import camera_rays as rays_fn
    ref_mod = sys.modules[type(kernel).__module__]                       # the reference's scene.blce
    Rt = kernel.get_Rt_c2w(cam)
    blur_feature = ref_mod.compute_frequency_blur_feature(cam.image.cuda())
    warped_c2w, exposure_time = kernel.model(Rt, blur_feature, cam.uid)  # [K,4,4], [K]            (blce.py:151)
    warped_w2c = torch.inverse(warped_c2w)                               # one batched inverse      (:152)
    R = warped_c2w[:, :3, :3]                                            # (:153)
    T = warped_w2c[:, :3, 3]                                             # (:154)
    Kc = R.shape[0]
    bottom = torch.tensor([0, 0, 0, 1], dtype=R.dtype, device=R.device).expand(Kc, 1, 4)
    # Camera(R, T): world_view_transform = getWorld2View2_torch(R, T)^T = [[R^T | T], [0 0 0 1]]^T   (cameras.py:126-130)
    viewmats = torch.cat([torch.cat([R.transpose(1, 2), T[..., None]], dim=-1), bottom], dim=1)
    wvt = viewmats.transpose(1, 2)
    centres = torch.inverse(wvt)[:, 3, :3]                               # cameras.py:130, batched
    m = cam.metadata
    W, H = int(m.image_size_x), int(m.image_size_y)
    intr = (float(m.principal_point_x), float(m.principal_point_y), float(m.scale_factor_x), float(m.scale_factor_y))
    out = WarpedCameras()
    out.extend(WarpedCamera(cam, R[k], T[k], wvt[k], centres[k], out, k) for k in range(Kc))
    out.viewmats = viewmats
    out._rays_fn = lambda: rays_fn(R.contiguous(), centres.contiguous(), *intr, W, H)
    from .cameras import ray_pose
    out.ray_pose = ray_pose(R, centres, *intr)
    return out, exposure_time
