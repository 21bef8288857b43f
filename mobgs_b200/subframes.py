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

from . import fused
from .gaussian_renderer import _bg10, _dynamic_params, _static_params


def render_subframes(stat_pc, dyn_pc, viewmats: torch.Tensor, Ks: torch.Tensor, t_spline: torch.Tensor,
                     t_poly: torch.Tensor, rays: torch.Tensor, bg_color: torch.Tensor, width: int,
                     height: int, center_k: Optional[int] = None, tight: bool = True,
                     offset: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """viewmats [K,4,4] (may require grad: they come from the BLCE pose network), Ks [K,3,3] or
    [3,3], t_spline / t_poly [K] device tensors (see gaussian_renderer._times), rays [K,6,H,W]
    (Camera.cam_ray of each warped camera) or [1,6,H,W].

    Returns: "render" [3,H,W] (the blurred prediction, mean over K + 1e-10), "subframes"
    [K,3,H,W], "depth" [K,H,W], "alpha" [K,H,W], "radii" [K,N], "viewspace_points" [1,N,2] (leaf whose
    .grad receives d loss / d means2d of sub-frame `center_k`, default K//2)."""
    K = viewmats.shape[0]
    dev = viewmats.device
    if Ks.dim() == 2:
        Ks = Ks[None].expand(K, -1, -1)
    ck = K // 2 if center_k is None else center_k
    records, radii, depths, _ = fused.synth_project(
        _static_params(stat_pc), _dynamic_params(dyn_pc), dyn_pc.current_control_num,
        viewmats, Ks, t_spline, t_poly, width, height, offset=offset)
    vsp = records[ck:ck + 1, :, 0:2].detach().clone().requires_grad_(True)
    bg10 = _bg10(bg_color, dev).expand(K, -1)
    dec = dyn_pc.rgbdecoder
    rgb, depth, alpha, mean = fused.blend_decode(
        records, radii, depths, bg10, rays, dec.mlp1.weight.reshape(6, 12), dec.mlp2.weight.reshape(3, 6),
        width, height, tight=tight, vsp=vsp, vsp_k=ck, want_mean=True)
    return {"render": mean, "subframes": rgb, "depth": depth, "alpha": alpha, "radii": radii,
            "viewspace_points": vsp, "visibility_filter": radii[ck] > 0}
