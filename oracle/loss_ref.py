"""ORACLE — test infrastructure only (tests/, smoke(), bench cpu_baseline may import it; the product never does).

CPU restatement of the photometric loss of the training step:

    l1_loss(network_output, gt)            utils/loss_utils.py:233-239 (mask=None branch)
    ssim(img1, img2)                       utils/loss_utils.py:338-382 (gaussian / create_window / _ssim)
    photo_loss = Ll1 + lambda_dssim * (1 - ssim)      train.py:621-628

PINNED: tests/golden/photo_loss.npz is produced by tests/golden/make_golden.py, which imports the
reference's own utils/loss_utils.py (CPU, fp32) and stores inputs, loss values and autograd gradients;
tests/test_oracle.py::test_photo_loss_oracle_matches_reference_golden holds this file to them.
"""
import math

import torch
import torch.nn.functional as F


def gaussian_window(window_size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return g / g.sum()


def l1_loss(x, y):
    return (x - y).abs().mean()


def ssim(img1, img2, window_size: int = 11):
    C = img1.size(-3)
    w1 = gaussian_window(window_size).to(img1.device, img1.dtype).unsqueeze(1)   # (loss_utils.py:355-357)
    window = (w1 @ w1.t()).unsqueeze(0).unsqueeze(0).expand(C, 1, window_size, window_size).contiguous()
    pad = window_size // 2
    conv = lambda t: F.conv2d(t, window, padding=pad, groups=C)
    mu1, mu2 = conv(img1), conv(img2)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = conv(img1 * img1) - mu1_sq
    s2 = conv(img2 * img2) - mu2_sq
    s12 = conv(img1 * img2) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()


def photo_loss(image, gt, lambda_dssim: float):
    return l1_loss(image, gt) + lambda_dssim * (1.0 - ssim(image, gt))


def l1_loss_masked(network_output, gt, mask):
    """utils/loss_utils.py:233-237 (mask branch)."""
    channel = gt.shape[1]
    mask = mask.expand(-1, channel, -1, -1)
    return torch.abs((network_output - gt) * mask).sum() / (mask.sum() + 1e-8)


def flow_warp_loss(ori_image, latent_img, exp2mid_coord, mid2exp_coord, latent_alpha, d_alpha):
    """train.py:656-676 without the lambda_flow_loss factor.  ori_image [B,3,H,W], latent_img [B,K,3,H,W],
    exp2mid_coord / mid2exp_coord [B,K,H,W,2] in pixels, latent_alpha [B,K,1,H,W], d_alpha [B,1,H,W].
    (The reference normalises the coordinate tensors in place; out of place here — same autograd graph.)
    Pinned by tests/golden/flow_warp_loss.npz (those train.py lines executed with the reference's own l1_loss)."""
    B, K, _, H, W = latent_img.shape

    def norm(c):
        c = torch.stack([c[..., 0] / (W - 1), c[..., 1] / (H - 1)], dim=-1)
        return (2.0 * c - 1.0).flatten(0, 1)

    ori_exp = ori_image.unsqueeze(1).expand(-1, K, -1, -1, -1).flatten(0, 1)
    warped_exp2mid = F.grid_sample(ori_exp, norm(exp2mid_coord), mode='bilinear', padding_mode='border').reshape(-1, K, 3, H, W)
    warped_mid2exp = F.grid_sample(latent_img.flatten(0, 1), norm(mid2exp_coord), mode='bilinear', padding_mode='border').reshape(-1, K, 3, H, W)
    return (l1_loss_masked(warped_exp2mid.flatten(0, 1), latent_img.flatten(0, 1), latent_alpha.flatten(0, 1))
            + l1_loss_masked(warped_mid2exp.flatten(0, 1), ori_exp, d_alpha.unsqueeze(1).expand(-1, K, -1, -1, -1).flatten(0, 1)))


def entropy_loss(alpha):
    """utils/loss_utils.py:264-276."""
    epsilon = 1e-6
    return -torch.sum(alpha * torch.log(alpha + epsilon) + (1 - alpha) * torch.log(1 - alpha + epsilon))


def sparsity_loss(alpha):
    """utils/loss_utils.py:285-295."""
    return torch.sum(alpha ** 2)


def reg_loss(depth, gt_depth, d_alpha, w_depth=0.2, w_mask=1e-7):
    """train.py:651-655.  Pinned by tests/golden/reg_loss.npz (the reference's own three functions)."""
    return w_depth * l1_loss(depth, gt_depth) + w_mask * entropy_loss(d_alpha) + w_mask * sparsity_loss(d_alpha)
