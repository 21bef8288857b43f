"""ORACLE — test infrastructure only (see oracle/gsplat_ref.py header: the gsplat half is PARITY
UNPINNED; the MoBGS half is pinned by tests/golden/).  Used by bench.py's `cpu_baseline` leg and
`--impl reference` arm, and by tests/.  The product path never imports it.

The reference's train step for one blurry view (train.py:441, :502-516, :540-541: K latent sub-frame
`render()` calls -> pixel mean -> L1 -> backward) on the CPU oracle, split into its two cost classes so
that a BOUNDED SAMPLE of a large workload can be timed and scaled honestly:

  * per-Gaussian work  (attribute synthesis, spline, projection; forward + backward) — done for ALL
    Gaussians of the workload and all K sub-frames, exactly as the full step would;
  * per-pixel work     (tile-list membership, alpha compositing, expected depth, Sandwich decoder, blur
    mean, L1; forward + backward) — done for ONE tile-aligned window of the frame, with exactly the
    per-pixel candidate sequences of the full-frame render (oracle.gsplat_ref.rasterize_to_pixels(window=)).

full-frame step time ~= t_gaussian + t_pixel * (W H) / (w h).
"""
from __future__ import annotations

import time
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import gsplat_ref as G
from . import mobgs_ref as M


def blurry_view_step(stat_pc, dyn_pc, cams: Sequence, deltas: Sequence[float], bg_color: torch.Tensor,
                     target: torch.Tensor, window: Optional[Tuple[int, int, int, int]] = None,
                     backward: bool = True) -> Dict[str, float]:
    """One train step of one blurry view on the oracle; `target` is [3,H,W] (cropped here when a
    window is given).  Follows render_ref's combined-render path (renderer :93-125, :181-227) per
    sub-frame and train.py:540-541 for the mean.  Returns the loss and wall-clock seconds per class."""
    K = len(cams)
    W, H = int(cams[0].image_width), int(cams[0].image_height)
    like = dyn_pc._scaling
    bg10 = torch.cat([bg_color[:3]] * 3 + [bg_color.new_zeros(1)], dim=-1)
    w1, w2 = M._decoder_weights(dyn_pc)
    t_g = t_p = 0.0

    # ---- per-Gaussian forward: synthesis + projection of every sub-frame ----
    t0 = time.perf_counter()
    s_means, s_quats, s_scales, s_opac, s_cols = M.static_attributes(stat_pc)
    geo = []
    for k, cam in enumerate(cams):
        tc = M._time_scalar(cam, like)
        warped = deltas[k] is not None
        t_cam = tc + deltas[k] / cam.max_time if warped else tc
        d_means, d_quats, d_scales, d_opac, d_cols = M.dynamic_attributes(dyn_pc, t_cam, clamp_time=warped)
        means = torch.cat([s_means, d_means], 0)
        quats = torch.cat([s_quats, d_quats], 0)
        scales = torch.cat([s_scales, d_scales], 0)
        opac = torch.cat([s_opac, d_opac], 0).squeeze(-1)
        cols = torch.cat([s_cols, d_cols], 0)
        viewmat = cam.world_view_transform.transpose(0, 1)
        radii, m2d, depths, conics, _ = G.fully_fused_projection(means, None, quats, scales, viewmat[None],
                                                                 cam.K[None], W, H)
        cols10 = torch.cat([cols, depths[0][:, None]], dim=-1)          # render_mode="RGB+ED"
        geo.append([m2d[0], conics[0], cols10, opac, radii[0], depths[0]])
    t_g += time.perf_counter() - t0

    # cut the graph between the two classes so that their backward passes can be timed separately
    leaves = []
    for gk in geo:
        for i in range(4):
            leaf = gk[i].detach().requires_grad_(gk[i].requires_grad)
            leaves.append((gk[i], leaf))
            gk[i] = leaf

    # ---- per-pixel forward ----
    t0 = time.perf_counter()
    if window is None:
        window_ = (0, 0, W, H)
    else:
        window_ = tuple(int(v) for v in window)
    x0, y0, w, h = window_
    imgs = []
    for k, cam in enumerate(cams):
        m2d, conics, cols10, opac, radii, depths = geo[k]
        rc, ra = G.rasterize_to_pixels(m2d, conics, cols10, opac, radii, depths, W, H, bg10, window=window)
        rays = cam.cam_ray[..., y0:y0 + h, x0:x0 + w]
        imgs.append(M.sandwich(rc[None, ..., :-1].permute(0, 3, 1, 2), rays, w1, w2).squeeze(0))
    pred = M.blur_mean(imgs)
    loss = (pred - target[:, y0:y0 + h, x0:x0 + w]).abs().mean()
    t_p += time.perf_counter() - t0

    if backward and loss.requires_grad:
        t0 = time.perf_counter()
        loss.backward()
        t_p += time.perf_counter() - t0
        t0 = time.perf_counter()
        outs = [o for o, l in leaves if l.grad is not None]
        grads = [l.grad for o, l in leaves if l.grad is not None]
        if outs:
            torch.autograd.backward(outs, grads)
        t_g += time.perf_counter() - t0
    return {"loss": float(loss.detach()), "t_gaussian": t_g, "t_pixel": t_p, "window_pixels": w * h,
            "frame_pixels": W * H, "K": K}


def extrapolate_full_step(t_gaussian: float, t_pixel: float, window_pixels: int, frame_pixels: int) -> float:
    """seconds of the full-frame step implied by a windowed sample"""
    return t_gaussian + t_pixel * (frame_pixels / float(window_pixels))
