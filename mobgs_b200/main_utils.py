"""Host-side mirror of the one function of the reference's main_utils.py that sits on the per-view training path:
`get_normals` (main_utils.py:95-141, called at train.py:590 on `pred_depth + 1e-6`).

The reference builds the pixel grid and the view directions with numpy on the host every call, ships the [H,W,3]
array to the device and runs ~12 torch launches; here it is one launch of `mobgs_depth_normals`
(csrc/normals.cu) that evaluates the view direction of each pixel in registers.  Same name, argument meaning and
return shape, so an unmodified train.py picks it up with

    import main_utils, mobgs_b200.main_utils
    main_utils.get_normals = mobgs_b200.main_utils.get_normals     (and train.get_normals, imported by name)

Forward only: the reference appends the result to `pred_normals` / `normal_tensor` (train.py:591, :615) and never
uses it in a loss, so no gradient is propagated to the depth.
"""
from __future__ import annotations

import torch

from . import _lib as L
from .ops import _f32c, _p, _stream


def depth_normals(z: torch.Tensor, ppx: float, ppy: float, sfx: float, sfy: float, skew: float = 0.0,
                  use_center: bool = True) -> torch.Tensor:
    """z [B,H,W] (CUDA) -> camera-space normals [B,3,H,W]; intrinsics as dycheck_geometry.camera.Camera names them."""
    z = _f32c(z.detach())
    if z.dim() != 3:
        raise ValueError(f"z must be [B,H,W], got {tuple(z.shape)}")
    B, H, W = z.shape
    out = torch.empty(B, 3, H, W, device=z.device)
    a = L.Normals(B, W, H, _p(z), float(ppx), float(ppy), float(sfx), float(sfy), float(skew),
                  0.5 if use_center else 0.0, _p(out))
    L.call("mobgs_depth_normals", a, _stream())
    return out


def get_normals(z: torch.Tensor, camera_metadata) -> torch.Tensor:
    """main_utils.get_normals(z [1,H,W], camera_metadata) -> [1,3,H,W].  `camera_metadata` is the view's
    dycheck_geometry Camera (viewpoint_cam.metadata); its image size must match z."""
    H, W = int(z.shape[-2]), int(z.shape[-1])
    if (int(camera_metadata.image_size_x), int(camera_metadata.image_size_y)) != (W, H):
        raise ValueError("camera_metadata.image_size does not match the depth map")
    return depth_normals(z.reshape(-1, H, W), camera_metadata.principal_point_x, camera_metadata.principal_point_y,
                         camera_metadata.scale_factor_x, camera_metadata.scale_factor_y, camera_metadata.skew,
                         bool(camera_metadata.use_center))
