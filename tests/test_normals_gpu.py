"""main_utils.get_normals (train.py:590) parity: mobgs_depth_normals against the golden vectors written by the
reference's own function (tests/golden/normals.npz, make_normals_golden.py) and against oracle/normals_ref.py on
seeded inputs up to 1080p.  Tolerance (fp32, unit vectors): 2e-5 absolute per component."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "normals.npz")
TOL = 2e-5


@pytest.mark.parametrize("tag", ["centre", "skewed", "thin"])
def test_depth_normals_match_reference_golden(tag):
    from mobgs_b200.main_utils import depth_normals
    z = np.load(GOLD)
    ppx, ppy, sfx, sfy, skew, off = (float(v) for v in z[tag + "_intr"])
    got = depth_normals(torch.from_numpy(z[tag + "_z"]).cuda(), ppx, ppy, sfx, sfy, skew, use_center=off > 0)
    want = z[tag + "_normals"]
    assert tuple(got.shape) == want.shape
    assert np.abs(got.cpu().numpy() - want).max() <= TOL


@pytest.mark.parametrize("B,W,H", [(1, 512, 288), (2, 97, 61), (1, 1920, 1080), (1, 2, 5), (0, 8, 8)])
def test_get_normals_matches_oracle(B, W, H):
    """the reference-named entry point with a camera_metadata duck (the attributes main_utils.py:96-100 reads)"""
    from mobgs_b200.main_utils import depth_normals, get_normals
    from oracle import normals_ref
    g = torch.Generator().manual_seed(B + W + H)
    z = 1.0 + 3.0 * torch.rand(B, H, W, generator=g)
    intr = dict(ppx=W * 0.47, ppy=H * 0.52, sfx=0.9 * W, sfy=0.93 * W, skew=0.02)
    want = normals_ref.get_normals(z.numpy(), intr["ppx"], intr["ppy"], intr["sfx"], intr["sfy"], intr["skew"], 0.5)
    got = depth_normals(z.cuda(), **intr)
    assert tuple(got.shape) == (B, 3, H, W)
    if B:
        assert np.abs(got.cpu().numpy() - want).max() <= TOL
        unit = got[:, :, 1:-1, 1:-1].norm(dim=1)
        if unit.numel():
            assert (unit - 1).abs().max() < 1e-5
    if B == 1:
        meta = types.SimpleNamespace(principal_point_x=np.float32(intr["ppx"]), principal_point_y=np.float32(intr["ppy"]),
                                     scale_factor_x=np.float32(intr["sfx"]), scale_factor_y=np.float32(intr["sfy"]),
                                     skew=np.float32(intr["skew"]), use_center=True, image_size_x=W, image_size_y=H)
        again = get_normals(z.cuda().requires_grad_(True), meta)
        assert torch.equal(again, got) and not again.requires_grad
