"""ORACLE — test infrastructure only.  PARITY UNPINNED (see below).

Dense pure-PyTorch CPU restatement of the two gsplat v1.4.0 operators the MoBGS
renderer calls (reference call sites: gaussian_renderer/__init__.py:143,163,190,
201,236,255,274,379,411,422,437,456,473,513,524,538):

    gsplat.rendering.fully_fused_projection(means, covars=None, quats, scales,
                                            viewmats, Ks, width, height)
    gsplat.rendering.rasterization(means, quats, scales, opacities, colors,
                                   viewmats, Ks, width, height, packed=False,
                                   render_mode="RGB"|"RGB+ED", backgrounds=...)

gsplat is an un-vendored pip dependency of the reference (README.md:26,
`gsplat==1.4.0`) and is not installable in this image, and the reference holds no
tests / golden vectors for this boundary.  The arithmetic
below restates gsplat 1.4.0's published CUDA kernels
(fully_fused_projection_fwd_kernel, isect_tiles, rasterize_to_pixels_fwd_kernel)
from their documented behaviour; every constant is a named module-level value.
Until gsplat itself can be run next to it, parity is *unpinned*; the
self-consistency tests in tests/test_oracle.py (closed-form single Gaussian,
occlusion order, permutation invariance, fp64 gradcheck) partially substitute.

Everything is differentiable by torch.autograd (the VJPs gsplat hand-writes are
the exact derivatives of these forward formulas, including the "no gradient when
alpha is clamped at 0.999" branch, which autograd's min() reproduces).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product path never does.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

# ---- gsplat 1.4.0 constants (SURVEY.md §8c lists them for re-verification) ----
EPS2D = 0.3            # rasterization(eps2d=0.3): added to the 2D covariance diagonal
NEAR_PLANE = 0.01      # rasterization(near_plane=0.01)
FAR_PLANE = 1e10       # rasterization(far_plane=1e10)
RADIUS_CLIP = 0.0      # rasterization(radius_clip=0.0)
TILE_SIZE = 16         # rasterization(tile_size=16)
ALPHA_MAX = 0.999      # rasterize_to_pixels: alpha = min(0.999, opac * exp(-sigma))
ALPHA_MIN = 1.0 / 255.0  # skip if alpha < 1/255
T_STOP = 1e-4          # stop *before* blending when T * (1 - alpha) <= 1e-4
FOV_MARGIN = 0.3       # persp_proj: clamp x/z to the image frustum + 0.3 * tan(fov/2)
RADIUS_DISC_FLOOR = 0.01  # radius = ceil(3 sqrt(b + sqrt(max(0.01, b^2 - det))))
ED_ALPHA_FLOOR = 1e-10    # "RGB+ED": depth channel / clamp(alpha, 1e-10)


def quat_to_rotmat(quats: torch.Tensor) -> torch.Tensor:
    """wxyz quaternion -> rotation matrix, normalising inside (gsplat quat_to_rotmat)."""
    q = quats / quats.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack(
        [
            1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y),
        ],
        dim=-1,
    )
    return R.reshape(quats.shape[:-1] + (3, 3))


def quat_scale_to_covar(quats: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    """Sigma = (R S)(R S)^T  (gsplat quat_scale_to_covar_preci)."""
    R = quat_to_rotmat(quats)
    M = R * scales[..., None, :]
    return M @ M.transpose(-1, -2)


def fully_fused_projection(
    means: torch.Tensor,            # [N,3]
    covars: Optional[torch.Tensor],  # None at every reference call site
    quats: torch.Tensor,            # [N,4] wxyz, need not be normalised
    scales: torch.Tensor,           # [N,3]
    viewmats: torch.Tensor,         # [C,4,4] world->camera
    Ks: torch.Tensor,               # [C,3,3]
    width: int,
    height: int,
    eps2d: float = EPS2D,
    near_plane: float = NEAR_PLANE,
    far_plane: float = FAR_PLANE,
    radius_clip: float = RADIUS_CLIP,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, None]:
    """Restates gsplat 1.4.0 fully_fused_projection_fwd_kernel (pinhole, packed=False).

    Returns (radii i32 [C,N], means2d [C,N,2], depths [C,N], conics [C,N,3], None).
    Culled Gaussians have radii == 0 and zeros elsewhere (gsplat leaves them
    uninitialised)."""
    assert covars is None, "reference always passes covars=None"
    C = viewmats.shape[0]
    dt = means.dtype
    R = viewmats[:, :3, :3]                       # [C,3,3]
    t = viewmats[:, :3, 3]                        # [C,3]
    mean_c = torch.einsum("cij,nj->cni", R, means) + t[:, None, :]   # [C,N,3]
    covar = quat_scale_to_covar(quats, scales)                      # [N,3,3]
    covar_c = torch.einsum("cij,njk,clk->cnil", R, covar, R)         # [C,N,3,3]

    fx = Ks[:, 0, 0][:, None]
    fy = Ks[:, 1, 1][:, None]
    cx = Ks[:, 0, 2][:, None]
    cy = Ks[:, 1, 2][:, None]
    x, y, z = mean_c.unbind(-1)
    valid = (z >= near_plane) & (z <= far_plane)
    # keep the arithmetic finite for culled Gaussians (their outputs are masked)
    zs = torch.where(valid, z, torch.ones_like(z))

    tan_fovx = 0.5 * width / fx
    tan_fovy = 0.5 * height / fy
    lim_x_pos = (width - cx) / fx + FOV_MARGIN * tan_fovx
    lim_x_neg = cx / fx + FOV_MARGIN * tan_fovx
    lim_y_pos = (height - cy) / fy + FOV_MARGIN * tan_fovy
    lim_y_neg = cy / fy + FOV_MARGIN * tan_fovy
    rz = 1.0 / zs
    rz2 = rz * rz
    tx = zs * torch.minimum(lim_x_pos, torch.maximum(-lim_x_neg, x * rz))
    ty = zs * torch.minimum(lim_y_pos, torch.maximum(-lim_y_neg, y * rz))
    zero = torch.zeros_like(zs)
    J = torch.stack(
        [fx * rz, zero, -fx * tx * rz2, zero, fy * rz, -fy * ty * rz2], dim=-1
    ).reshape(C, -1, 2, 3)
    cov2d = J @ covar_c @ J.transpose(-1, -2)                        # [C,N,2,2]
    mean2d = torch.stack([fx * x * rz + cx, fy * y * rz + cy], dim=-1)

    a = cov2d[..., 0, 0] + eps2d
    b = 0.5 * (cov2d[..., 0, 1] + cov2d[..., 1, 0])
    c = cov2d[..., 1, 1] + eps2d
    det = a * c - b * b
    valid = valid & (det > 0)
    dets = torch.where(valid, det, torch.ones_like(det))
    conics = torch.stack([c / dets, -b / dets, a / dets], dim=-1)

    with torch.no_grad():
        mid = 0.5 * (a + c)
        v1 = mid + torch.sqrt(torch.clamp(mid * mid - det, min=RADIUS_DISC_FLOOR))
        radius = torch.ceil(3.0 * torch.sqrt(torch.clamp(v1, min=0)))
        valid = valid & (radius > radius_clip)
        mx, my = mean2d.detach().unbind(-1)
        inside = ~(
            (mx + radius <= 0) | (mx - radius >= width) | (my + radius <= 0) | (my - radius >= height)
        )
        valid = valid & inside
        radii = torch.where(valid, radius, torch.zeros_like(radius)).to(torch.int32)

    m = valid
    means2d = torch.where(m[..., None], mean2d, torch.zeros_like(mean2d))
    depths = torch.where(m, z, torch.zeros_like(z))
    conics = torch.where(m[..., None], conics, torch.zeros_like(conics))
    return radii, means2d.to(dt), depths.to(dt), conics.to(dt), None


def tile_bounds(means2d: torch.Tensor, radii: torch.Tensor, tile_w: int, tile_h: int,
                tile_size: int = TILE_SIZE):
    """isect_tiles: inclusive tile_min / exclusive tile_max of the square AABB
    mean +- radius, clamped to the tile grid (float->uint32 casts saturate at 0)."""
    r = radii.to(means2d.dtype) / tile_size
    tx = means2d[..., 0] / tile_size
    ty = means2d[..., 1] / tile_size
    x0 = torch.clamp(torch.floor(tx - r), 0, tile_w).long()
    y0 = torch.clamp(torch.floor(ty - r), 0, tile_h).long()
    x1 = torch.clamp(torch.ceil(tx + r), 0, tile_w).long()
    y1 = torch.clamp(torch.ceil(ty + r), 0, tile_h).long()
    vis = radii > 0
    x1 = torch.where(vis, x1, x0)
    y1 = torch.where(vis, y1, y0)
    return x0, y0, x1, y1


def rasterize_to_pixels(
    means2d: torch.Tensor,      # [N,2]
    conics: torch.Tensor,       # [N,3]
    colors: torch.Tensor,       # [N,D]
    opacities: torch.Tensor,    # [N]
    radii: torch.Tensor,        # [N] int
    depths: torch.Tensor,       # [N]
    width: int,
    height: int,
    backgrounds: Optional[torch.Tensor] = None,  # [D] or None
    tile_size: int = TILE_SIZE,
    pixel_chunk: int = 8192,
    return_stats: bool = False,
    window: Optional[Tuple[int, int, int, int]] = None,
):
    """isect_tiles + stable (tile, depth) sort + rasterize_to_pixels_fwd_kernel, dense.

    For every pixel the candidate sequence is: Gaussians whose tile AABB covers the
    pixel's 16x16 tile, in ascending depth (ties: ascending index — the stable LSB
    radix sort keeps emission order).

    window = (x0, y0, w, h), tile aligned: only the pixels of that window are evaluated (output
    [h,w,D]) and only the Gaussians whose tile AABB touches the window's tiles enter the dense
    pixel x Gaussian arithmetic — per pixel exactly the candidate sequence of the full-frame
    render, so a crop of a 1 M-Gaussian / 1080p frame costs what a small frame costs."""
    N = means2d.shape[0]
    D = colors.shape[-1]
    dev, dt = means2d.device, means2d.dtype
    tile_w = math.ceil(width / tile_size)
    tile_h = math.ceil(height / tile_size)

    order = torch.argsort(depths.detach(), stable=True)
    vis = radii[order] > 0
    order = order[vis]
    m2 = means2d[order]
    cn = conics[order]
    col = colors[order]
    op = opacities[order]
    x0, y0, x1, y1 = tile_bounds(m2.detach(), radii[order], tile_w, tile_h, tile_size)
    if window is None:
        wx0, wy0, ww, wh = 0, 0, width, height
    else:
        wx0, wy0, ww, wh = (int(v) for v in window)
        assert wx0 % tile_size == 0 and wy0 % tile_size == 0 and wx0 >= 0 and wy0 >= 0
        assert wx0 + ww <= width and wy0 + wh <= height
        t0x, t0y = wx0 // tile_size, wy0 // tile_size
        t1x, t1y = (wx0 + ww - 1) // tile_size + 1, (wy0 + wh - 1) // tile_size + 1
        keep = (x0 < t1x) & (x1 > t0x) & (y0 < t1y) & (y1 > t0y)
        m2, cn, col, op = m2[keep], cn[keep], col[keep], op[keep]
        x0, y0, x1, y1 = x0[keep], y0[keep], x1[keep], y1[keep]

    P = ww * wh
    out_c = []
    out_a = []
    n_pairs = 0
    n_blend = 0
    for s in range(0, P, pixel_chunk):
        e = min(P, s + pixel_chunk)
        pid = torch.arange(s, e, device=dev)
        pyi = pid // ww + wy0
        pxi = pid % ww + wx0
        px = pxi.to(dt) + 0.5
        py = pyi.to(dt) + 0.5
        tx = (pxi // tile_size)[:, None]
        ty = (pyi // tile_size)[:, None]
        in_tile = (tx >= x0[None]) & (tx < x1[None]) & (ty >= y0[None]) & (ty < y1[None])
        dx = m2[None, :, 0] - px[:, None]
        dy = m2[None, :, 1] - py[:, None]
        sigma = 0.5 * (cn[None, :, 0] * dx * dx + cn[None, :, 2] * dy * dy) + cn[None, :, 1] * dx * dy
        alpha = torch.clamp(op[None] * torch.exp(-sigma), max=ALPHA_MAX)
        ok = in_tile & (sigma >= 0) & (alpha >= ALPHA_MIN)
        a = torch.where(ok, alpha, torch.zeros_like(alpha))
        one_m = 1.0 - a
        T_incl = torch.cumprod(one_m, dim=1)                      # T after each candidate
        with torch.no_grad():
            stop = ok & (T_incl <= T_STOP)
            stopped = torch.cumsum(stop.to(torch.int32), dim=1) > 0   # this one and all after
        a = torch.where(stopped, torch.zeros_like(a), a)
        one_m = 1.0 - a
        T_incl = torch.cumprod(one_m, dim=1)
        T_excl = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], dim=1)
        w = a * T_excl
        c = w @ col
        T_fin = T_incl[:, -1] if T_incl.shape[1] > 0 else torch.ones(e - s, device=dev, dtype=dt)
        if backgrounds is not None:
            c = c + T_fin[:, None] * backgrounds[None, :]
        out_c.append(c)
        out_a.append(1.0 - T_fin)
        if return_stats:
            n_pairs += int(in_tile.sum())
            n_blend += int((a > 0).sum())
    rc = torch.cat(out_c, 0).reshape(wh, ww, D)
    ra = torch.cat(out_a, 0).reshape(wh, ww, 1)
    if return_stats:
        return rc, ra, {"pixel_pairs": n_pairs, "blended_pairs": n_blend}
    return rc, ra


def rasterization(
    means: torch.Tensor,
    quats: torch.Tensor,
    scales: torch.Tensor,
    opacities: torch.Tensor,     # [N]
    colors: torch.Tensor,        # [N,D]
    viewmats: torch.Tensor,      # [C,4,4]
    Ks: torch.Tensor,            # [C,3,3]
    width: int,
    height: int,
    near_plane: float = NEAR_PLANE,
    far_plane: float = FAR_PLANE,
    radius_clip: float = RADIUS_CLIP,
    eps2d: float = EPS2D,
    sh_degree=None,
    packed: bool = False,
    tile_size: int = TILE_SIZE,
    backgrounds: Optional[torch.Tensor] = None,   # [C,D]
    render_mode: str = "RGB",
    window: Optional[Tuple[int, int, int, int]] = None,
    **unused,
):
    """Restates gsplat.rendering.rasterization for the kwargs the reference passes.

    Returns (render_colors [C,H,W,D(+1)], render_alphas [C,H,W,1], meta).  `window` (oracle-only
    extension, see rasterize_to_pixels): project everything, rasterise one tile-aligned crop."""
    assert sh_degree is None and not packed
    assert render_mode in ("RGB", "RGB+ED", "RGB+D", "D", "ED")
    C = viewmats.shape[0]
    radii, means2d, depths, conics, _ = fully_fused_projection(
        means, None, quats, scales, viewmats, Ks, width, height,
        eps2d=eps2d, near_plane=near_plane, far_plane=far_plane, radius_clip=radius_clip)
    if colors.dim() == 2:
        colors = colors[None].expand(C, -1, -1)
    if render_mode in ("RGB+D", "RGB+ED"):
        colors = torch.cat([colors, depths[..., None]], dim=-1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, torch.zeros_like(backgrounds[:, :1])], dim=-1)
    elif render_mode in ("D", "ED"):
        colors = depths[..., None]
        if backgrounds is not None:
            backgrounds = torch.zeros_like(backgrounds[:, :1])
    rcs, ras = [], []
    for c in range(C):
        rc, ra = rasterize_to_pixels(
            means2d[c], conics[c], colors[c], opacities, radii[c], depths[c], width, height,
            backgrounds[c] if backgrounds is not None else None, tile_size, window=window)
        rcs.append(rc)
        ras.append(ra)
    render_colors = torch.stack(rcs, 0)
    render_alphas = torch.stack(ras, 0)
    if render_mode in ("ED", "RGB+ED"):
        render_colors = torch.cat(
            [render_colors[..., :-1],
             render_colors[..., -1:] / render_alphas.clamp(min=ED_ALPHA_FLOOR)], dim=-1)
    meta = {
        "radii": radii, "means2d": means2d, "depths": depths, "conics": conics,
        "opacities": opacities[None].expand(C, -1),
        "tile_width": math.ceil(width / tile_size), "tile_height": math.ceil(height / tile_size),
        "width": width, "height": height, "tile_size": tile_size, "n_cameras": C,
    }
    return render_colors, render_alphas, meta
