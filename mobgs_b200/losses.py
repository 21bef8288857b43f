"""Fused photometric loss with the reference's function names.

    photo_loss = l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))          train.py:621-628

`l1_loss` / `ssim` have the signatures of utils/loss_utils.py:233 / :351 for the arguments train.py
uses (mask=None, window_size=11, size_average=True); `photo_loss` is the fused form of train.py:621-628
— one forward and one backward kernel for both terms (INTEGRATION.md shows the two-line edit).
Gradients flow to the first argument only (the ground-truth image never requires grad in the
reference).  CUDA fp32 only; there is no fallback path.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib


def _window11() -> np.ndarray:
    """utils/loss_utils.py:338-341 gaussian(11, 1.5): float32 taps, normalised in float32."""
    g = torch.tensor([math.exp(-(x - 11 // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)], dtype=torch.float32)
    return (g / g.sum()).numpy()


_WINDOW = _window11()


def _check(img, gt):
    if not (img.is_cuda and gt.is_cuda and img.dtype == torch.float32 and gt.dtype == torch.float32):
        raise RuntimeError("mobgs_b200.losses needs CUDA fp32 tensors (there is no CPU fallback)")
    if img.shape != gt.shape or img.dim() < 2:
        raise RuntimeError(f"shape mismatch {tuple(img.shape)} vs {tuple(gt.shape)}")


class _PhotoLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, gt, lambda_dssim, want):
        _check(img, gt)
        img_c, gt_c = img.contiguous(), gt.contiguous()
        H, W = img_c.shape[-2:]
        planes = img_c.numel() // (H * W)
        need_grad = img.requires_grad
        sums = torch.empty(2, dtype=torch.float64, device=img.device)
        a = _lib.PhotoLossFwd()
        a.planes, a.H, a.W = planes, H, W
        a.img, a.gt, a.sums = img_c.data_ptr(), gt_c.data_ptr(), sums.data_ptr()
        for i in range(11):
            a.window[i] = float(_WINDOW[i])
        maps = None
        if need_grad:
            maps = torch.empty(3, planes, H, W, device=img.device)
            a.d_mu1, a.d_x2, a.d_xy = maps[0].data_ptr(), maps[1].data_ptr(), maps[2].data_ptr()
        _lib.call("mobgs_photo_loss_fwd", a, _lib.current_stream())
        n = img_c.numel()
        l1 = (sums[0] / n).float()
        ss = (sums[1] / n).float()
        ctx.save_for_backward(img_c, gt_c, maps)
        ctx.lam, ctx.want, ctx.shape = float(lambda_dssim), want, img.shape
        if want == "l1":
            return l1
        if want == "ssim":
            return ss
        return l1 + lambda_dssim * (1.0 - ss)

    @staticmethod
    def backward(ctx, g):
        img_c, gt_c, maps = ctx.saved_tensors
        H, W = img_c.shape[-2:]
        planes = img_c.numel() // (H * W)
        n = img_c.numel()
        g = g.contiguous().float()
        v = torch.empty_like(img_c)
        a = _lib.PhotoLossBwd()
        a.planes, a.H, a.W = planes, H, W
        a.img, a.gt = img_c.data_ptr(), gt_c.data_ptr()
        for i in range(11):
            a.window[i] = float(_WINDOW[i])
        a.d_mu1, a.d_x2, a.d_xy = maps[0].data_ptr(), maps[1].data_ptr(), maps[2].data_ptr()
        a.v_loss = g.data_ptr()
        a.scale_l1 = {"l1": 1.0 / n, "ssim": 0.0}.get(ctx.want, 1.0 / n)
        a.scale_ssim = {"l1": 0.0, "ssim": 1.0 / n}.get(ctx.want, -ctx.lam / n)
        a.v_img = v.data_ptr()
        _lib.call("mobgs_photo_loss_bwd", a, _lib.current_stream())
        return v.view(ctx.shape), None, None, None


def photo_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float) -> torch.Tensor:
    """train.py:621-628: `l1_loss(image, gt) + lambda_dssim * (1.0 - ssim(image, gt))` in two launches."""
    return _PhotoLoss.apply(image, gt, lambda_dssim, "photo")


def l1_loss(network_output: torch.Tensor, gt: torch.Tensor, mask=None) -> torch.Tensor:
    """utils/loss_utils.py:233-239 with mask=None (the photometric / depth use, train.py:621, :651): one
    streaming launch (mobgs_reg_loss_fwd with an empty alpha term) that also leaves sign(x - y) for the
    backward, which is then a scalar multiply."""
    if mask is not None:
        raise RuntimeError("mobgs_b200.losses.l1_loss implements the mask=None form only")
    _check(network_output, gt)
    empty = network_output.new_empty(0)
    return _RegLoss.apply(network_output, gt, empty, 1.0, 0.0)[0]


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True) -> torch.Tensor:
    """utils/loss_utils.py:351-382 with the defaults train.py:626 uses."""
    if window_size != 11 or not size_average:
        raise RuntimeError("mobgs_b200.losses.ssim implements window_size=11, size_average=True only")
    return _PhotoLoss.apply(img1, img2, 0.0, "ssim")


class _FlowWarpLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ori, latent, exp2mid, mid2exp, latent_alpha, d_alpha):
        ts = [t.contiguous() for t in (ori, latent, exp2mid, mid2exp, latent_alpha, d_alpha)]
        for t in ts:
            if not (t.is_cuda and t.dtype == torch.float32):
                raise RuntimeError("flow_warp_loss needs CUDA fp32 tensors (there is no CPU fallback)")
        ori_c, lat_c, e2m_c, m2e_c, la_c, da_c = ts
        B, K, C, H, W = lat_c.shape
        if C != 3 or ori_c.shape != (B, 3, H, W) or e2m_c.shape != (B, K, H, W, 2) or m2e_c.shape != (B, K, H, W, 2) \
                or la_c.numel() != B * K * H * W or da_c.numel() != B * H * W:
            raise RuntimeError("flow_warp_loss: shape mismatch")
        sums = torch.empty(4, dtype=torch.float64, device=ori.device)
        a = _lib.FlowWarp()
        a.B, a.K, a.H, a.W = B, K, H, W
        a.ori, a.latent, a.exp2mid, a.mid2exp = ori_c.data_ptr(), lat_c.data_ptr(), e2m_c.data_ptr(), m2e_c.data_ptr()
        a.latent_alpha, a.d_alpha, a.sums = la_c.data_ptr(), da_c.data_ptr(), sums.data_ptr()
        _lib.call("mobgs_flow_warp_loss_fwd", a, _lib.current_stream())
        ctx.save_for_backward(ori_c, lat_c, e2m_c, m2e_c, la_c, da_c, sums)
        ctx.shapes = (latent.shape, exp2mid.shape, mid2exp.shape, latent_alpha.shape, d_alpha.shape)
        ctx.ori_shape = ori.shape
        return (sums[0] / (sums[1] + 1e-8) + sums[2] / (sums[3] + 1e-8)).float()

    @staticmethod
    def backward(ctx, g):
        ori_c, lat_c, e2m_c, m2e_c, la_c, da_c, sums = ctx.saved_tensors
        B, K, _, H, W = lat_c.shape
        g = g.contiguous().float()
        v_lat, v_e2m, v_m2e = torch.empty_like(lat_c), torch.empty_like(e2m_c), torch.empty_like(m2e_c)
        v_la, v_da = torch.empty_like(la_c), torch.empty_like(da_c)
        v_ori = torch.empty_like(ori_c) if ctx.needs_input_grad[0] else None
        a = _lib.FlowWarp()
        a.B, a.K, a.H, a.W = B, K, H, W
        a.ori, a.latent, a.exp2mid, a.mid2exp = ori_c.data_ptr(), lat_c.data_ptr(), e2m_c.data_ptr(), m2e_c.data_ptr()
        a.latent_alpha, a.d_alpha, a.sums = la_c.data_ptr(), da_c.data_ptr(), sums.data_ptr()
        a.v_loss = g.data_ptr()
        a.v_latent, a.v_exp2mid, a.v_mid2exp = v_lat.data_ptr(), v_e2m.data_ptr(), v_m2e.data_ptr()
        a.v_latent_alpha, a.v_d_alpha = v_la.data_ptr(), v_da.data_ptr()
        if v_ori is not None:
            a.v_ori = v_ori.data_ptr()
        _lib.call("mobgs_flow_warp_loss_bwd", a, _lib.current_stream())
        s = ctx.shapes
        return (v_ori.view(ctx.ori_shape) if v_ori is not None else None), v_lat.view(s[0]), v_e2m.view(s[1]), v_m2e.view(s[2]), v_la.view(s[3]), v_da.view(s[4])


def flow_warp_loss(ori_image, latent_img, exp2mid_coord, mid2exp_coord, latent_alpha, d_alpha) -> torch.Tensor:
    """The flow-warp loss of train.py:656-676 without its `lambda_flow_loss` factor, as one forward + one
    backward kernel.  ori_image [B,3,H,W] (`ori_image_tensor`), latent_img [B,K,3,H,W]
    (`latent_img_final_tensor`), exp2mid_coord / mid2exp_coord [B,K,H,W,2] in pixels (NOT normalised — the
    reference's in-place normalisation of train.py:659-662 / :667-670 happens inside), latent_alpha [B,K,1,H,W],
    d_alpha [B,1,H,W].  Gradients flow to every argument, `ori_image` included: in the reference it is the live
    centre render (train.py:469, :607), reached through the exp2mid grid_sample and as the mid2exp L1 target."""
    return _FlowWarpLoss.apply(ori_image, latent_img, exp2mid_coord, mid2exp_coord, latent_alpha, d_alpha)


class _RegLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, gt_depth, d_alpha, w_depth, w_mask):
        d_c, g_c, a_c = depth.contiguous(), gt_depth.contiguous(), d_alpha.contiguous()
        for t in (d_c, g_c, a_c):
            if not (t.is_cuda and t.dtype == torch.float32):
                raise RuntimeError("reg_loss needs CUDA fp32 tensors (there is no CPU fallback)")
        if d_c.shape != g_c.shape:
            raise RuntimeError("reg_loss: depth / gt_depth shape mismatch")
        need = depth.requires_grad or d_alpha.requires_grad
        sums = torch.empty(3, dtype=torch.float64, device=depth.device)
        gd = torch.empty_like(d_c) if need else None
        ga = torch.empty_like(a_c) if need else None
        a = _lib.RegLoss()
        a.n_depth, a.n_alpha = d_c.numel(), a_c.numel()
        a.depth, a.gt_depth, a.alpha, a.sums = d_c.data_ptr(), g_c.data_ptr(), a_c.data_ptr(), sums.data_ptr()
        if need:
            a.g_depth, a.g_alpha = gd.data_ptr(), ga.data_ptr()
        _lib.call("mobgs_reg_loss_fwd", a, _lib.current_stream())
        ctx.save_for_backward(gd, ga)
        ctx.scales = (w_depth / max(d_c.numel(), 1), w_mask)
        ctx.shapes = (depth.shape, d_alpha.shape)
        ctx.mark_non_differentiable(sums)
        total = (w_depth * sums[0] / max(d_c.numel(), 1) + w_mask * (sums[1] + sums[2])).float()
        return total, sums

    @staticmethod
    def backward(ctx, g, _g_sums):
        gd, ga = ctx.saved_tensors
        sd, sa = ctx.scales
        return (g * sd) * gd.view(ctx.shapes[0]), None, (g * sa) * ga.view(ctx.shapes[1]), None, None


def reg_loss(depth, gt_depth, d_alpha, w_depth: float = 0.2, w_mask: float = 1e-7):
    """train.py:651-655: `0.2 * l1_loss(depth, gt_depth) + 1e-7 * entropy_loss(d_alpha) + 1e-7 * sparsity_loss(d_alpha)`
    in one launch.  Returns (reg, sums) with sums = [sum |depth - gt|, entropy, sparsity] (float64, for logging:
    depth_loss = sums[0] / depth.numel())."""
    return _RegLoss.apply(depth, gt_depth, d_alpha, float(w_depth), float(w_mask))
