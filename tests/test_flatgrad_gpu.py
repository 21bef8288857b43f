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


def test_chunked_backward_with_gradient_sink_equals_plain_backward():
    """mobgs_b200.dist.overlap_gradient_allreduce: the projection backward split into Gaussian ranges (each range's
    gradient block handed to the collective on a side stream, then unpacked by mobgs_copy_segments) must deliver
    the gradients of the single-launch backward (1e-5 of each tensor's max: two runs differ at rounding level because
    the blend backward's atomics are unordered).
    On one rank the collective is the identity, which is what lets this run on a single GPU."""
    import torch
    from mobgs_b200 import dist as D
    from mobgs_b200 import fused
    from mobgs_b200.scene import subframe_w2c, synthetic_scene
    from mobgs_b200.subframes import render_subframes
    from mobgs_b200.cameras import camera_rays_from_w2c
    K, W, H = 3, 96, 64
    sc, dc, intr = synthetic_scene(1500, 900, W, H, seed=8, device="cuda")
    Kmat = torch.tensor([[intr.fx, 0, intr.cx], [0, intr.fy, intr.cy], [0, 0, 1.0]], device="cuda")
    tp = torch.tensor([0.45, 0.5, 0.55], device="cuda")
    tgt = torch.rand(3, H, W, device="cuda")
    params = [p for pc in (sc, dc) for p in pc.parameters() if p.requires_grad]

    def run():
        for p in params:
            p.grad = None
        view = torch.stack([subframe_w2c(k, K) for k in range(K)]).cuda().requires_grad_(True)
        rays = camera_rays_from_w2c(view, intr.fx, intr.fy, intr.cx, intr.cy, W, H)
        out = render_subframes(sc, dc, view, Kmat, tp, tp, rays, torch.zeros(3, device="cuda"), W, H)
        (out["render"] - tgt).abs().mean().backward()
        torch.cuda.synchronize()
        return [None if p.grad is None else p.grad.clone() for p in params], view.grad.clone()

    plain, v_plain = run()
    for n_chunks in (2, 4, 7):
        sink = D.overlap_gradient_allreduce(True, n_chunks=n_chunks)
        try:
            chunked, v_chunked = run()
            fg = D.FlatGradients([p for p in params if p.grad is not None], inplace_shared=True)
            fg.reduce()                      # recognises the already-reduced buffer
            assert sink.reduced_storage is None
        finally:
            D.overlap_gradient_allreduce(False)
        assert fused.GRAD_SINK is None
        for a, b, p in zip(plain, chunked, params):
            assert (a is None) == (b is None)
            if a is None:
                continue
            # two runs are not bit-reproducible (the blend backward accumulates its gradient records with atomics,
            # the order of which varies); what the split changes must stay at that rounding level
            assert (a - b).abs().max() <= 1e-5 * a.abs().max(), (tuple(p.shape), float((a - b).abs().max()))
        assert (v_plain - v_chunked).abs().max() <= 1e-5 * v_plain.abs().max()


def test_recycled_gradient_records_give_identical_gradients():
    """The gradient-record buffer is kept across steps and zeroed behind the projection backward's reads
    (MobgsSynthBwd.zero_v_records): repeated steps — with a different view in between, so that the set of visible
    Gaussians changes — must give bit-identical gradients with recycling on and off, and a caller-supplied gradient
    tensor must never be modified."""
    import torch
    from mobgs_b200 import fused
    from mobgs_b200.cameras import ray_pose_from_w2c
    from mobgs_b200.scene import make_camera, subframe_w2c, synthetic_scene
    from mobgs_b200.subframes import render_subframes
    K, W, H = 3, 96, 64
    sc, dc, intr = synthetic_scene(300, 200, W, H, seed=9, device="cuda")
    params = [p for pc in (sc, dc) for p in pc.parameters() if p.requires_grad]
    Kmat = make_camera(intr, subframe_w2c(0, K)).K.cuda()
    t = torch.full((K,), 0.5).cuda()
    bg = torch.zeros(3, device="cuda")
    views = [torch.stack([subframe_w2c(k, K) for k in range(K)]).cuda(),
             torch.stack([subframe_w2c(k, K) @ subframe_w2c(2, 3) for k in range(K)]).cuda()]
    tgt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(0)).cuda()

    def run(view):
        for p in params:
            p.grad = None
        rp = ray_pose_from_w2c(view, intr.fx, intr.fy, intr.cx, intr.cy, rigid=True)
        out = render_subframes(sc, dc, view, Kmat, t, t, rp, bg, W, H)
        ((out["render"] - tgt).abs().mean() + out["depth"].mean() * 0.1).backward()
        return [None if p.grad is None else p.grad.clone() for p in params]

    old = fused.RECYCLE_GRAD_RECORDS
    try:
        fused.RECYCLE_GRAD_RECORDS = False
        ref = [run(views[i % 2]) for i in range(4)]
        fused.RECYCLE_GRAD_RECORDS = True
        fused._GRAD_REC_POOL.clear(); fused._GRAD_REC_OUT.clear()
        got = [run(views[i % 2]) for i in range(4)]
        assert len(fused._GRAD_REC_POOL) == 1, "the buffer must be back in the pool after a step"
        pooled = next(iter(fused._GRAD_REC_POOL.values()))
        assert float(pooled.abs().max()) == 0.0, "the recycled buffer must be all zero between steps"
    finally:
        fused.RECYCLE_GRAD_RECORDS = old
    for step, (a, b) in enumerate(zip(got, ref)):
        for x, y, p in zip(a, b, params):
            assert (x is None) == (y is None)
            if x is not None and x.numel():
                # atomics make the scatter order nondeterministic: compare to rounding, not bit for bit
                assert float((x - y).abs().max()) <= 1e-5 * (float(y.abs().max()) + 1e-12), (step, tuple(p.shape))
    # a gradient tensor supplied by the caller is read-only
    rec, radii, depths, _ = fused.synth_project(
        (sc._xyz, sc._rotation, sc._scaling, sc._opacity, sc._features_dc),
        (dc.control_xyz, dc._rotation, dc._omega, dc._scaling, dc._opacity, dc._features_dc, dc._features_t, dc._trbf_center),
        dc.current_control_num, views[0], Kmat[None].expand(K, -1, -1), t, t, W, H)
    g = torch.rand_like(rec)
    g0 = g.clone()
    rec.backward(g)
    assert torch.equal(g, g0)
