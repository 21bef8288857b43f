"""f2 parity: mobgs_b200.losses (mobgs_photo_loss_fwd / _bwd) against the golden vectors written by the
reference's own utils/loss_utils.py (tests/golden/photo_loss.npz) and against oracle/loss_ref.py on
seeded inputs.  Tolerance of this row (fp32, written here): 1e-5 relative on the loss values,
1e-4 of the gradient's max magnitude on gradients (north_star: 1e-4 abs / 1e-3 rel)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "photo_loss.npz")


def _grad_close(g, ref):
    scale = float(np.abs(ref).max())
    assert np.abs(g - ref).max() <= 1e-4 * scale + 1e-12, (float(np.abs(g - ref).max()), scale)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_photo_loss_matches_reference_golden(tag):
    from mobgs_b200 import losses
    z = np.load(GOLD)
    img = torch.from_numpy(z[f"{tag}_img"]).cuda().requires_grad_(True)
    gt = torch.from_numpy(z[f"{tag}_gt"]).cuda()
    lam = float(z[f"{tag}_lambda"])
    assert abs(float(losses.l1_loss(img, gt)) - float(z[f"{tag}_l1"])) <= 1e-5 * abs(float(z[f"{tag}_l1"]))
    assert abs(float(losses.ssim(img, gt)) - float(z[f"{tag}_ssim"])) <= 1e-5
    loss = losses.photo_loss(img, gt, lam)
    assert abs(float(loss) - float(z[f"{tag}_loss"])) <= 1e-5 * abs(float(z[f"{tag}_loss"]))
    loss.backward()
    _grad_close(img.grad.cpu().numpy(), z[f"{tag}_grad"])


@pytest.mark.parametrize("shape", [(1, 3, 288, 512), (2, 3, 33, 17), (1, 1, 16, 16), (3, 3, 5, 40)])
def test_photo_loss_matches_oracle(shape):
    from mobgs_b200 import losses
    from oracle import loss_ref as L
    g = torch.Generator().manual_seed(sum(shape))
    gt = torch.rand(shape, generator=g)
    img = (gt + 0.2 * torch.randn(shape, generator=g)).clamp(0, 1)
    img[..., :2, :3] = gt[..., :2, :3]                 # exact zeros of x - y: sign(0) = 0 like torch.abs
    a = img.clone().requires_grad_(True)
    ref = L.photo_loss(a, gt, 0.2)
    up = 0.37                                          # a non-trivial upstream gradient
    (ref * up).backward()
    b = img.clone().cuda().requires_grad_(True)
    out = losses.photo_loss(b, gt.cuda(), 0.2)
    (out * up).backward()
    assert abs(float(out) - float(ref)) <= 1e-5 * abs(float(ref))
    _grad_close(b.grad.cpu().numpy(), a.grad.numpy())
    # the separate terms and their gradients
    for ours, theirs in ((losses.l1_loss, L.l1_loss), (losses.ssim, L.ssim)):
        a2 = img.clone().requires_grad_(True)
        b2 = img.clone().cuda().requires_grad_(True)
        r, o = theirs(a2, gt), ours(b2, gt.cuda())
        r.backward(); o.backward()
        assert abs(float(o) - float(r)) <= 1e-5 * max(abs(float(r)), 1e-3)
        _grad_close(b2.grad.cpu().numpy(), a2.grad.numpy())


def test_photo_loss_forward_only_and_noncontiguous():
    from mobgs_b200 import losses
    from oracle import loss_ref as L
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 40, 56, 3, generator=g)
    y = torch.rand(2, 40, 56, 3, generator=g)
    xi, yi = x.permute(0, 3, 1, 2), y.permute(0, 3, 1, 2)            # channels-last views
    with torch.no_grad():
        out = losses.photo_loss(xi.cuda(), yi.cuda(), 0.2)
    assert abs(float(out) - float(L.photo_loss(xi, yi, 0.2))) <= 1e-5
    with pytest.raises(RuntimeError):
        losses.photo_loss(xi, yi, 0.2)                                  # CPU tensors: no fallback


def test_reg_loss_matches_reference_golden_and_oracle():
    """train.py:651-655 (depth L1 + entropy + sparsity of d_alpha) in one launch."""
    from mobgs_b200 import losses
    from oracle import loss_ref as L
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "reg_loss.npz"))
    d = torch.from_numpy(z["depth"]).cuda().requires_grad_(True)
    a = torch.from_numpy(z["d_alpha"]).cuda().requires_grad_(True)
    reg, sums = losses.reg_loss(d, torch.from_numpy(z["gt_depth"]).cuda(), a)
    assert abs(float(reg) - float(z["reg"])) <= 1e-5 * abs(float(z["reg"]))
    assert abs(float(sums[0]) / d.numel() - float(z["depth_loss"])) <= 1e-5 * float(z["depth_loss"])
    assert abs(float(sums[1]) - float(z["entropy"])) <= 1e-5 * abs(float(z["entropy"]))
    assert abs(float(sums[2]) - float(z["sparsity"])) <= 1e-5 * float(z["sparsity"])
    (reg * 3.0).backward()
    _grad_close(d.grad.cpu().numpy() / 3.0, z["g_depth"])
    _grad_close(a.grad.cpu().numpy() / 3.0, z["g_d_alpha"])
    # a larger seeded case against the oracle
    g = torch.Generator().manual_seed(9)
    dd, gt, al = 1 + 3 * torch.rand(2, 1, 288, 512, generator=g), 1 + 3 * torch.rand(2, 1, 288, 512, generator=g), torch.rand(2, 1, 288, 512, generator=g)
    do, ao = dd.clone().requires_grad_(True), al.clone().requires_grad_(True)
    ref = L.reg_loss(do, gt, ao)
    ref.backward()
    dc, ac = dd.cuda().requires_grad_(True), al.cuda().requires_grad_(True)
    out, _ = losses.reg_loss(dc, gt.cuda(), ac)
    out.backward()
    assert abs(float(out) - float(ref)) <= 1e-5 * abs(float(ref))
    _grad_close(dc.grad.cpu().numpy(), do.grad.numpy())
    _grad_close(ac.grad.cpu().numpy(), ao.grad.numpy())
