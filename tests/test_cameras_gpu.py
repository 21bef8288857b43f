"""f3 parity: mobgs_b200.cameras.camera_rays (mobgs_camera_rays_fwd / _bwd) against the golden written by
the reference's own scene/cameras.py methods and against oracle.mobgs_ref.camera_rays_ref.
Tolerance (fp32, written here): 1e-6 abs on the rays (unit vectors), 1e-4 of the max magnitude on
the 12 pose gradients (sums over H*W pixels)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "camera_rays.npz")


def test_camera_rays_match_reference_golden():
    from mobgs_b200.cameras import camera_rays
    z = np.load(GOLD)
    ppx, ppy, sfx, sfy, W, H = z["geom"]
    rot = torch.from_numpy(z["rot"]).cuda().requires_grad_(True)
    cen = torch.from_numpy(z["centre"]).cuda().requires_grad_(True)
    rays = camera_rays(rot, cen, ppx, ppy, sfx, sfy, int(W), int(H))
    assert np.abs(rays.detach().cpu().numpy() - z["rays"]).max() < 1e-6
    (rays * torch.from_numpy(z["weight"]).cuda()).sum().backward()
    assert np.abs(rot.grad.cpu().numpy() - z["g_rot"]).max() <= 1e-4 * np.abs(z["g_rot"]).max()
    assert np.abs(cen.grad.cpu().numpy() - z["g_centre"]).max() <= 1e-4 * np.abs(z["g_centre"]).max()


@pytest.mark.parametrize("K,W,H", [(1, 512, 288), (9, 160, 90), (2, 1, 1), (4, 17, 300)])
def test_camera_rays_match_oracle(K, W, H):
    from mobgs_b200.cameras import camera_rays, camera_rays_from_w2c
    from mobgs_b200.scene import subframe_w2c
    from oracle import mobgs_ref as M
    g = torch.Generator().manual_seed(K * 1000 + W)
    w2c = torch.stack([subframe_w2c(k, max(K, 2)) for k in range(K)])
    w2c[:, :3, 3] += 0.1 * torch.randn(K, 3, generator=g)
    c2w = torch.inverse(w2c)
    fx, fy, cx, cy = 0.9 * W, 0.8 * W, W / 2 + 0.3, H / 2 - 0.2
    rot_o, cen_o = c2w[:, :3, :3].clone().requires_grad_(True), c2w[:, :3, 3].clone().requires_grad_(True)
    ref = M.camera_rays_ref(rot_o, cen_o, cx, cy, fx, fy, W, H)
    wgt = torch.randn(ref.shape, generator=g)
    (ref * wgt).sum().backward()
    rot_c, cen_c = c2w[:, :3, :3].cuda().requires_grad_(True), c2w[:, :3, 3].cuda().requires_grad_(True)
    out = camera_rays(rot_c, cen_c, cx, cy, fx, fy, W, H)
    (out * wgt.cuda()).sum().backward()
    assert (out.detach().cpu() - ref.detach()).abs().max() < 1e-6
    for a, b in ((rot_c.grad.cpu(), rot_o.grad), (cen_c.grad.cpu(), cen_o.grad)):
        assert (a - b).abs().max() <= 1e-4 * b.abs().max() + 1e-7
    # the pinhole convenience entry point agrees with the PinholeCamera stand-in of the tests
    out2 = camera_rays_from_w2c(w2c.cuda(), fx, fy, cx, cy, W, H)
    assert (out2.cpu() - ref.detach()).abs().max() < 2e-6
