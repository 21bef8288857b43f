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


def project_subframes(stat_pc, dyn_pc, cams: Sequence, deltas: Sequence[Optional[float]]):
    """Per-Gaussian half of K sub-frame renders (renderer :93-125, :181-201): for every camera the combined
    static + dynamic attribute set projected by the oracle.  -> list of
    [means2d [N,2], conics [N,3], colors10 [N,10] (9 features + camera depth, "RGB+ED"), opacities [N],
     radii i32 [N], depths [N]] with the autograd graph back to the parameters attached."""
    W, H = int(cams[0].image_width), int(cams[0].image_height)
    like = dyn_pc._scaling
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
    return geo


def render_window(geo, cams: Sequence, dyn_pc, bg_color: torch.Tensor, window=None):
    """Per-pixel half: rasterise (one tile-aligned window of) every sub-frame from `geo`, expected depth,
    Sandwich decoder (renderer :201-227).  -> (rgb [K,3,h,w], depth [K,h,w], alpha [K,h,w])."""
    W, H = int(cams[0].image_width), int(cams[0].image_height)
    x0, y0, w, h = (0, 0, W, H) if window is None else tuple(int(v) for v in window)
    bg10 = torch.cat([bg_color[:3]] * 3 + [bg_color.new_zeros(1)], dim=-1)
    w1, w2 = M._decoder_weights(dyn_pc)
    rgb, dep, alp = [], [], []
    for k, cam in enumerate(cams):
        m2d, conics, cols10, opac, radii, depths = geo[k]
        rc, ra = G.rasterize_to_pixels(m2d, conics, cols10, opac, radii, depths, W, H, bg10, window=window)
        rays = cam.cam_ray[..., y0:y0 + h, x0:x0 + w]
        rgb.append(M.sandwich(rc[None, ..., :-1].permute(0, 3, 1, 2), rays, w1, w2).squeeze(0))
        dep.append(rc[..., -1] / ra[..., 0].clamp(min=G.ED_ALPHA_FLOOR))
        alp.append(ra[..., 0])
    return torch.stack(rgb), torch.stack(dep), torch.stack(alp)


def blurry_view_step(stat_pc, dyn_pc, cams: Sequence, deltas: Sequence[float], bg_color: torch.Tensor,
                     target: torch.Tensor, window: Optional[Tuple[int, int, int, int]] = None,
                     backward: bool = True) -> Dict[str, float]:
    """One train step of one blurry view on the oracle; `target` is [3,H,W] (cropped here when a
    window is given).  Follows render_ref's combined-render path (renderer :93-125, :181-227) per
    sub-frame and train.py:540-541 for the mean.  Returns the loss and wall-clock seconds per class."""
    K = len(cams)
    W, H = int(cams[0].image_width), int(cams[0].image_height)
    t_g = t_p = 0.0

    # ---- per-Gaussian forward: synthesis + projection of every sub-frame ----
    t0 = time.perf_counter()
    geo = project_subframes(stat_pc, dyn_pc, cams, deltas)
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
    x0, y0, w, h = (0, 0, W, H) if window is None else tuple(int(v) for v in window)
    rgb, _, _ = render_window(geo, cams, dyn_pc, bg_color, window)
    pred = M.blur_mean(list(rgb))
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


@torch.no_grad()
def threshold_events(geo_k, width: int, height: int, window, rel_alpha: float = 2e-5, rel_T: float = 2e-4):
    """Where can two correct fp32 implementations of gsplat's rasteriser legitimately disagree?  Only where
    a discrete decision sits on its threshold: alpha >= 1/255, sigma >= 0, alpha clamped at 0.999,
    T (1 - alpha) <= 1e-4 (per pixel), and ceil(3 sqrt(lambda)) (per Gaussian, changes tile membership).
    For one sub-frame's projected set `geo_k` and one window this returns
      flip_pixels [h,w] bool  — pixels with a candidate within the relative margins of a per-pixel threshold,
      affected [N] bool       — Gaussians that are candidates (tile-list members with alpha >= ~1/255) at a flip
                                pixel: a flip changes the transmittance of everything behind it at that pixel.
    Margins: alpha differs between implementations by the exponent's rounding (|p| <= 8, ex2.approx) ~ 1e-6
    relative -> 2e-5; T is a product of up to a few hundred (1 - alpha) factors -> 2e-4."""
    m2d, conics, _cols, opac, radii, depths = [t.detach() for t in geo_k]
    x0, y0, w, h = (int(v) for v in window)
    ts = G.TILE_SIZE
    tile_w, tile_h = -(-width // ts), -(-height // ts)
    order = torch.argsort(depths, stable=True)
    order = order[radii[order] > 0]
    tx0, ty0, tx1, ty1 = G.tile_bounds(m2d[order], radii[order], tile_w, tile_h, ts)
    keep = (tx0 < (x0 + w - 1) // ts + 1) & (tx1 > x0 // ts) & (ty0 < (y0 + h - 1) // ts + 1) & (ty1 > y0 // ts)
    ids = order[keep]
    tx0, ty0, tx1, ty1 = tx0[keep], ty0[keep], tx1[keep], ty1[keep]
    pid = torch.arange(w * h)
    pyi, pxi = pid // w + y0, pid % w + x0
    px, py = pxi.to(m2d.dtype) + 0.5, pyi.to(m2d.dtype) + 0.5
    tx, ty = (pxi // ts)[:, None], (pyi // ts)[:, None]
    in_tile = (tx >= tx0[None]) & (tx < tx1[None]) & (ty >= ty0[None]) & (ty < ty1[None])
    dx, dy = m2d[ids, 0][None] - px[:, None], m2d[ids, 1][None] - py[:, None]
    cn = conics[ids]
    sigma = 0.5 * (cn[None, :, 0] * dx * dx + cn[None, :, 2] * dy * dy) + cn[None, :, 1] * dx * dy
    raw = opac[ids][None] * torch.exp(-sigma)
    alpha = raw.clamp(max=G.ALPHA_MAX)
    ok = in_tile & (sigma >= 0) & (alpha >= G.ALPHA_MIN)
    a = torch.where(ok, alpha, torch.zeros_like(alpha))
    T_incl = torch.cumprod(1.0 - a, dim=1)
    stop = ok & (T_incl <= G.T_STOP)
    live = torch.cumsum(stop.to(torch.int32), dim=1) == 0                   # candidates reached before the stop
    near = in_tile & live & (
        ((raw - G.ALPHA_MIN).abs() <= rel_alpha * G.ALPHA_MIN)
        | ((raw - G.ALPHA_MAX).abs() <= rel_alpha * G.ALPHA_MAX)
        | (sigma.abs() <= 1e-6)
        | (ok & ((T_incl - G.T_STOP).abs() <= rel_T * G.T_STOP)))
    # first stop candidate itself is "reached": include its T test
    first_stop = stop & (torch.cumsum(stop.to(torch.int32), dim=1) == 1)
    near |= first_stop & ((T_incl - G.T_STOP).abs() <= rel_T * G.T_STOP)
    flip = near.any(dim=1)
    touched = (in_tile & (raw >= G.ALPHA_MIN * (1 - rel_alpha)) & (sigma >= -1e-6))[flip].any(dim=0)
    affected = torch.zeros(m2d.shape[0], dtype=torch.bool)
    affected[ids[touched]] = True
    return flip.reshape(h, w), affected


@torch.no_grad()
def radius_on_threshold(means, quats, scales, viewmat, Kmat, width, height, margin: float = 2e-4):
    """Gaussians whose 3-sigma radius 3 sqrt(lambda_max) lies within `margin` (relative) of an integer: the
    ceil() that sizes the tile AABB can legitimately differ there.  -> [N] bool"""
    R, t = viewmat[:3, :3], viewmat[:3, 3]
    mean_c = means @ R.T + t
    covar = G.quat_scale_to_covar(quats, scales)
    covar_c = torch.einsum("ij,njk,lk->nil", R, covar, R)
    fx, fy, cx, cy = Kmat[0, 0], Kmat[1, 1], Kmat[0, 2], Kmat[1, 2]
    x, y, z = mean_c.unbind(-1)
    z = torch.where(z > 0, z, torch.ones_like(z))
    lx0, lx1 = cx / fx + G.FOV_MARGIN * 0.5 * width / fx, (width - cx) / fx + G.FOV_MARGIN * 0.5 * width / fx
    ly0, ly1 = cy / fy + G.FOV_MARGIN * 0.5 * height / fy, (height - cy) / fy + G.FOV_MARGIN * 0.5 * height / fy
    tx = z * torch.minimum(lx1, torch.maximum(-lx0, x / z))
    ty = z * torch.minimum(ly1, torch.maximum(-ly0, y / z))
    zero = torch.zeros_like(z)
    J = torch.stack([fx / z, zero, -fx * tx / (z * z), zero, fy / z, -fy * ty / (z * z)], dim=-1).reshape(-1, 2, 3)
    c2 = J @ covar_c @ J.transpose(-1, -2)
    a, b, c = c2[:, 0, 0] + G.EPS2D, c2[:, 0, 1], c2[:, 1, 1] + G.EPS2D
    mid = 0.5 * (a + c)
    r = 3.0 * torch.sqrt(mid + torch.sqrt((mid * mid - (a * c - b * b)).clamp(min=G.RADIUS_DISC_FLOOR)))
    return (r - torch.round(r)).abs() <= margin * r.clamp(min=1.0)
