"""Host-side logic of this round that needs no GPU: the RayPose container (in-kernel camera rays), the closed-form rigid
inverse against torch.inverse and against the reference's camera-ray formula (oracle), the BinPlan bookkeeping of the
fused counting pass, and the gradient-record pool."""
import math

import pytest
import torch


def test_rigid_c2w_equals_torch_inverse_and_is_differentiable():
    from mobgs_b200.cameras import rigid_c2w
    from mobgs_b200.scene import subframe_w2c
    K = 5
    view = torch.stack([subframe_w2c(k, K) for k in range(K)]).double().requires_grad_(True)
    rot, cen = rigid_c2w(view)
    inv = torch.inverse(view)
    assert torch.allclose(rot, inv[:, :3, :3], atol=1e-12) and torch.allclose(cen, inv[:, :3, 3], atol=1e-12)
    g = torch.autograd.grad((rot.sum() + (cen * torch.arange(3.0).double()).sum()), view)[0]
    assert g.shape == view.shape and float(g.abs().max()) > 0


def test_ray_pose_container_and_reference_ray_formula():
    """RayPose holds [K,12] = (R row-major | centre); the ray it stands for is the reference's Camera.cam_ray
    (oracle.mobgs_ref.camera_rays_ref restates scene/cameras.py:132-146): checked here with the same arithmetic the
    kernels use (csrc/ray_math.cuh), in fp64 on the CPU."""
    from mobgs_b200.cameras import RayPose, ray_pose, ray_pose_from_w2c
    from mobgs_b200.scene import subframe_w2c
    from oracle.mobgs_ref import camera_rays_ref
    K, W, H = 3, 20, 12
    fx = fy = 0.9 * W
    cx, cy = W / 2, H / 2
    view = torch.stack([subframe_w2c(k, K) for k in range(K)])
    rp = ray_pose_from_w2c(view, fx, fy, cx, cy, rigid=True)
    assert isinstance(rp, RayPose) and tuple(rp.shape) == (K, 12) and rp.intr == (cx, cy, fx, fy)
    assert tuple(rp[1:3].shape) == (2, 12) and tuple(rp[1].shape) == (1, 12)
    with pytest.raises(ValueError):
        RayPose(torch.zeros(3, 11), 0, 0, 1, 1)
    c2w = torch.inverse(view)
    rp2 = ray_pose(c2w[:, :3, :3], c2w[:, :3, 3], cx, cy, fx, fy)
    assert torch.allclose(rp.pose, rp2.pose, atol=1e-6)
    want = camera_rays_ref(c2w[:, :3, :3], c2w[:, :3, 3], cx, cy, fx, fy, W, H)      # [K,6,H,W]
    pose = rp.pose.double()
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    lx, ly = (xs + 0.5 - cx) / fx, (ys + 0.5 - cy) / fy
    rn = torch.rsqrt(lx * lx + ly * ly + 1)
    l = torch.stack([lx * rn, ly * rn, rn])                                           # [3,H,W]
    for k in range(K):
        R = pose[k, :9].reshape(3, 3)
        w = torch.einsum("ij,jhw->ihw", R, l)
        d = w * torch.rsqrt((w * w).sum(0, keepdim=True))
        assert torch.allclose(d.float(), want[k, 3:], atol=2e-6)
        assert torch.allclose(pose[k, 9:].float()[:, None, None].expand(3, H, W), want[k, :3], atol=1e-6)


def test_bin_plan_uses_the_capacity_guess_of_the_same_launch_shape():
    from mobgs_b200 import ops
    dev = torch.device("cpu")
    ops._CAP_CACHE.clear()
    N, W, H = 1000, 64, 48
    specs = [(0, 0, N), (1, 0, N), (0, 300, N)]
    plan = ops.BinPlan(specs, W, H, tight=True)
    assert plan.prepare(N, dev) is None                          # no guess yet: the stand-alone counting kernels run
    tiles = math.ceil(W / 16) * math.ceil(H / 16)
    key = (3, N, W, H, tuple(tuple(s) for s in specs), True, dev.index)
    assert plan.key(N, dev) == key
    ops._CAP_CACHE[key] = 5000
    lists, counts, entries, cursor = plan.prepare(N, dev)
    assert counts.numel() == 3 * tiles and tuple(entries.shape) == (5000, 4) and cursor.numel() == 1
    assert [lists.rec_k[i] for i in range(3)] == [0, 1, 0] and lists.g_begin[2] == 300
    ops._CAP_CACHE[key] = int(ops.FUSED_COUNT_MAX_TILES * 3 * N) + 4097 + 1       # long lists: not fused
    assert ops.BinPlan(specs, W, H, tight=True).prepare(N, dev) is None
    old = ops.FUSED_COUNT
    try:
        ops.FUSED_COUNT = False
        ops._CAP_CACHE[key] = 5000
        assert ops.BinPlan(specs, W, H, tight=True).prepare(N, dev) is None
    finally:
        ops.FUSED_COUNT = old
        ops._CAP_CACHE.clear()


def test_gradient_record_pool_recycles_only_its_own_live_buffers():
    from mobgs_b200 import fused
    dev = torch.device("cpu")
    fused._GRAD_REC_POOL.clear(); fused._GRAD_REC_OUT.clear()
    a = fused._take_grad_records(2, 10, dev)
    assert tuple(a.shape) == (2, 10, 16) and float(a.abs().max()) == 0
    assert fused._recyclable(a)
    assert not fused._recyclable(torch.zeros(2, 10, 16))             # somebody else's tensor
    assert not fused._recyclable(a[:1])                              # wrong shape
    fused._recycle(a)
    assert fused._take_grad_records(2, 10, dev) is a                 # handed out again
    b = fused._take_grad_records(2, 10, dev)                         # pool empty: a fresh zero buffer
    assert b is not a
    fused._GRAD_REC_POOL.clear(); fused._GRAD_REC_OUT.clear()


def test_graph_capture_bookkeeping_and_no_cpu_path():
    """Host side of the CUDA-graphed step: build_tile_lists refuses to be captured without a capacity guess (the capture
    must never fall back to the synchronising exact-size path), and GraphedStep / the new single-launch helpers refuse to
    run without a CUDA device instead of routing around the library."""
    from mobgs_b200 import graphs, main_utils, ops
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        graphs.GraphedStep(lambda x: x, [torch.zeros(1)])
    with pytest.raises(RuntimeError):
        main_utils.depth_normals(torch.ones(1, 4, 4), 2.0, 2.0, 3.0, 3.0)
    assert ops.CAPTURE is None

    class FakeCapture:
        checks, keep = [], []
    ops.CAPTURE = FakeCapture()
    try:
        rec = torch.zeros(1, 4, 16)
        with pytest.raises(RuntimeError, match="capture"):
            ops.build_tile_lists(rec, torch.zeros(1, 4, dtype=torch.int32), torch.zeros(1, 4), 37, 29, consume=lambda l: None)
    finally:
        ops.CAPTURE = None
