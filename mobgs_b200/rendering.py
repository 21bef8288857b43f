"""Drop-in for the two `gsplat.rendering` operators the MoBGS renderer imports
(reference gaussian_renderer/__init__.py:15): same names, argument meaning and return
structure as gsplat 1.4.0 for the kwargs the reference passes, backed by
libmobgs_b200.so.  `compat/gsplat/rendering.py` re-exports these so the unmodified reference
renderer file runs on top of them.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from . import ops

TILE_SIZE = 16
ED_ALPHA_FLOOR = 1e-10

# default for the exact-results list pruning described in include/mobgs_b200.h (mobgs_tile_count)
TIGHT_TILES = True


def fully_fused_projection(
    means: torch.Tensor, covars: Optional[torch.Tensor], quats: torch.Tensor, scales: torch.Tensor,
    viewmats: torch.Tensor, Ks: torch.Tensor, width: int, height: int, eps2d: float = 0.3,
    near_plane: float = 0.01, far_plane: float = 1e10, radius_clip: float = 0.0,
    packed: bool = False, sparse_grad: bool = False, calc_compensations: bool = False,
    camera_model: str = "pinhole",
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, None]:
    """-> (radii i32[C,N], means2d[C,N,2], depths[C,N], conics[C,N,3], compensations=None)."""
    if covars is not None or packed or sparse_grad or calc_compensations or camera_model != "pinhole":
        raise NotImplementedError(
            "mobgs_b200.fully_fused_projection covers the reference's call pattern only: "
            "covars=None, packed=False, pinhole, no compensations")
    radii, means2d, depths, conics = ops.project(
        means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip)
    return radii, means2d, depths, conics, None


def rasterization(
    means: torch.Tensor, quats: torch.Tensor, scales: torch.Tensor, opacities: torch.Tensor,
    colors: torch.Tensor, viewmats: torch.Tensor, Ks: torch.Tensor, width: int, height: int,
    near_plane: float = 0.01, far_plane: float = 1e10, radius_clip: float = 0.0, eps2d: float = 0.3,
    sh_degree: Optional[int] = None, packed: bool = True, tile_size: int = TILE_SIZE,
    backgrounds: Optional[torch.Tensor] = None, render_mode: str = "RGB", sparse_grad: bool = False,
    absgrad: bool = False, rasterize_mode: str = "classic", channel_chunk: int = 32,
    distributed: bool = False, camera_model: str = "pinhole", covars: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, torch.Tensor, Dict]:
    """-> (render_colors [C,H,W,D(+1)], render_alphas [C,H,W,1], meta)."""
    if (sh_degree is not None or packed or sparse_grad or absgrad or distributed or covars is not None
            or rasterize_mode != "classic" or camera_model != "pinhole" or tile_size != TILE_SIZE):
        raise NotImplementedError(
            "mobgs_b200.rasterization covers the reference's call pattern only: sh_degree=None, "
            "packed=False, classic, pinhole, tile_size=16")
    if render_mode not in ("RGB", "RGB+ED", "RGB+D", "D", "ED"):
        raise ValueError(render_mode)
    C = viewmats.shape[0]
    radii, means2d, depths, conics = ops.project(
        means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip)
    append_depth = render_mode in ("RGB+D", "RGB+ED")
    if render_mode in ("D", "ED"):
        colors = depths[..., None]
        if backgrounds is not None:
            backgrounds = torch.zeros(C, 1, device=means.device)
    render_colors, render_alphas = ops.rasterize(
        means2d, conics, colors, opacities, backgrounds, depths, radii, width, height,
        append_depth=append_depth, tight=TIGHT_TILES)
    if render_mode in ("ED", "RGB+ED"):
        render_colors = torch.cat(
            [render_colors[..., :-1],
             render_colors[..., -1:] / render_alphas.clamp(min=ED_ALPHA_FLOOR)], dim=-1)
    meta = {
        "camera_ids": None, "gaussian_ids": None,
        "radii": radii, "means2d": means2d, "depths": depths, "conics": conics,
        "opacities": opacities[None].expand(C, -1),
        "tile_width": math.ceil(width / tile_size), "tile_height": math.ceil(height / tile_size),
        "width": width, "height": height, "tile_size": tile_size, "n_cameras": C,
    }
    return render_colors, render_alphas, meta
