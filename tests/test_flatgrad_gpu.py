"""The Gaussian-parameter gradients of the fused path come out as views of ONE flat buffer that autograd
adopts as `.grad` (fused._SynthProject.backward), so dist.FlatGradients can all-reduce them in place."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_parameter_gradients_share_one_storage_and_reduce_in_place():
    from mobgs_b200.dist import FlatGradients
    from mobgs_b200.scene import make_camera, subframe_w2c, synthetic_scene
    from mobgs_b200.subframes import render_subframes
    K, W, H = 3, 64, 48
    stat, dyn, intr = synthetic_scene(301, 203, W, H, seed=3, device="cuda")     # odd sizes: padded segments
    cams = [make_camera(intr, subframe_w2c(k, K)) for k in range(K)]
    view = torch.stack([subframe_w2c(k, K) for k in range(K)]).cuda()
    t = torch.full((K,), 0.5, device="cuda")
    rays = torch.cat([c.cam_ray for c in cams]).cuda()
    out = render_subframes(stat, dyn, view, cams[0].K.cuda(), t, t, rays, torch.zeros(3, device="cuda"), W, H)
    out["render"].mean().backward()
    names = [(stat, "_xyz"), (stat, "_rotation"), (stat, "_scaling"), (stat, "_opacity"), (stat, "_features_dc"),
             (dyn, "control_xyz"), (dyn, "_rotation"), (dyn, "_omega"), (dyn, "_scaling"), (dyn, "_opacity"),
             (dyn, "_features_dc"), (dyn, "_features_t")]
    grads = [getattr(pc, n).grad for pc, n in names]
    assert all(g is not None for g in grads)
    assert len({g.untyped_storage().data_ptr() for g in grads}) == 1, "autograd did not adopt the flat-buffer views"
    assert all(g.data_ptr() % 16 == 0 for g in grads)
    params = [getattr(pc, n) for pc, n in names]
    before = [g.clone() for g in grads]
    fg = FlatGradients(params, inplace_shared=True)
    fg.reduce()                                   # single process: no collective, nothing may change
    assert fg.last_collective_elems >= sum(p.numel() for p in params)
    for p, b in zip(params, before):
        assert torch.equal(p.grad, b)
