"""f1 (second half) parity: mobgs_b200.losses.flow_warp_loss (mobgs_flow_warp_loss_fwd / _bwd) against the
golden written from train.py:656-676 with the reference's own l1_loss, and against oracle.loss_ref on seeded
inputs — gradients of every input, `ori_image_tensor` included (live in the reference, train.py:469, :607).  Tolerance (fp32, written here): 1e-5 relative on the loss, 1e-4 of the max magnitude on gradients
(north_star: 1e-4 abs / 1e-3 rel); at most 1e-3 of the coordinate-gradient elements may differ more — a
sample position that lands within rounding of a pixel boundary picks the other bilinear cell."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "flow_warp_loss.npz")
KEYS = ("ori", "latent", "exp2mid", "mid2exp", "latent_alpha", "d_alpha")


def _check_grads(got, ref):
    for k in KEYS:
        g, r = got[k], ref[k]
        scale = float(np.abs(r).max())
        bad = np.abs(g - r) > 1e-4 * scale + 1e-12
        assert bad.mean() <= (1e-3 if "2" in k else 0.0), (k, float(np.abs(g - r).max()), scale, float(bad.mean()))


def _run_ours(t, up=1.0):
    from mobgs_b200.losses import flow_warp_loss
    c = {k: v.clone().cuda().requires_grad_(True) for k, v in t.items()}
    loss = flow_warp_loss(c["ori"], c["latent"], c["exp2mid"], c["mid2exp"], c["latent_alpha"], c["d_alpha"])
    (loss * up).backward()
    return float(loss), {k: c[k].grad.cpu().numpy() for k in KEYS}


def test_flow_warp_loss_matches_reference_golden():
    z = np.load(GOLD)
    t = {k: torch.from_numpy(z[k]) for k in KEYS}
    loss, grads = _run_ours(t)
    assert abs(loss - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    _check_grads(grads, {k: z["g_" + k] for k in KEYS})


@pytest.mark.parametrize("B,K,H,W", [(1, 9, 72, 128), (2, 1, 5, 7), (3, 2, 33, 2)])
def test_flow_warp_loss_matches_oracle(B, K, H, W):
    from oracle import loss_ref as L
    g = torch.Generator().manual_seed(B * 100 + K)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    base = torch.stack([xs, ys], -1)[None, None].expand(B, K, -1, -1, -1)
    t = {"ori": torch.rand(B, 3, H, W, generator=g), "latent": torch.rand(B, K, 3, H, W, generator=g),
         "exp2mid": (base + 2.0 * torch.randn(B, K, H, W, 2, generator=g)).contiguous(),
         "mid2exp": (base + 2.0 * torch.randn(B, K, H, W, 2, generator=g)).contiguous(),
         "latent_alpha": torch.rand(B, K, 1, H, W, generator=g), "d_alpha": torch.rand(B, 1, H, W, generator=g)}
    t["d_alpha"][:, :, :1] = 0.0                       # masked-out pixels
    o = {k: v.clone().requires_grad_(True) for k, v in t.items()}
    ref = L.flow_warp_loss(o["ori"], o["latent"], o["exp2mid"], o["mid2exp"], o["latent_alpha"], o["d_alpha"])
    (ref * 0.7).backward()
    loss, grads = _run_ours(t, up=0.7)
    assert abs(loss - float(ref)) <= 1e-5 * abs(float(ref))
    _check_grads(grads, {k: o[k].grad.numpy() for k in KEYS})
