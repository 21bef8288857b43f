"""Rays generated inside the blend kernels (MobgsBlendFwd/Bwd.dec_pose, mobgs_b200.cameras.RayPose) against the path
that materialises Camera.cam_ray as a [K,6,H,W] image (mobgs_camera_rays_fwd/bwd, itself pinned to the reference's
scene/cameras.py by tests/test_cameras_gpu.py): same renders, same parameter / decoder / view-matrix gradients, and
the same gradients of the camera-to-world rotation and camera centre (the pose gradient that trains the BLCE
network) — with image sizes that are not tile multiples and with lists that share a camera."""
import pytest
import torch

from mobgs_b200.scene import make_camera, subframe_w2c, synthetic_scene

pytestmark = pytest.mark.gpu


def _leaves(K, dev="cuda"):
    view = torch.stack([subframe_w2c(k, K) for k in range(K)]).to(dev)
    c2w = torch.inverse(view)
    rot = c2w[:, :3, :3].contiguous().clone().requires_grad_(True)
    cen = c2w[:, :3, 3].contiguous().clone().requires_grad_(True)
    return view, rot, cen


def _grads(params):
    out = [None if p.grad is None else p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    return out


def _same(a, b, what, rtol=2e-4):
    assert (a is None) == (b is None), what
    if a is None or b.numel() == 0:
        return
    scale = float(b.abs().max()) + 1e-30
    err = float((a - b).abs().max())
    assert err <= rtol * scale + 1e-9, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("ns,nd,W,H,K", [(400, 300, 97, 61, 3), (0, 200, 64, 48, 1), (500, 0, 130, 70, 5)])
def test_render_subframes_ray_pose_equals_ray_images(ns, nd, W, H, K):
    from mobgs_b200.cameras import camera_rays, ray_pose
    from mobgs_b200.subframes import render_subframes
    sc, dc, intr = synthetic_scene(ns, nd, W, H, seed=5 + K, device="cuda")
    params = [p for pc in (sc, dc) for p in pc.parameters() if p.requires_grad]
    bg = torch.tensor([0.2, 0.4, 0.1], device="cuda")
    Kmat = make_camera(intr, subframe_w2c(0, K)).K.cuda()
    t_poly = (0.5 + torch.linspace(-1, 1, K) * 0.4 / 23 if K > 1 else torch.tensor([0.52])).cuda()
    tgt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(1)).cuda()
    wd = (torch.rand(K, H, W, generator=torch.Generator().manual_seed(2)) * 0.1).cuda()
    res = []
    for mode in ("image", "pose"):
        view, rot, cen = _leaves(K)
        view = view.clone().requires_grad_(True)
        geom = (intr.cx, intr.cy, intr.fx, intr.fy)
        rays = camera_rays(rot, cen, *geom, W, H) if mode == "image" else ray_pose(rot, cen, *geom)
        out = render_subframes(sc, dc, view, Kmat, t_poly.clamp(0, 1), t_poly, rays, bg, W, H)
        ((out["render"] - tgt).abs().mean() + (out["depth"] * wd).mean() + 0.3 * out["subframes"].square().mean()).backward()
        res.append((out, rot.grad.clone(), cen.grad.clone(), view.grad.clone(), _grads(params), out["viewspace_points"].grad))
    (oi, g_rot_i, g_cen_i, g_view_i, gp_i, vsp_i), (op, g_rot_p, g_cen_p, g_view_p, gp_p, vsp_p) = res
    for key in ("render", "subframes", "depth", "alpha"):
        _same(op[key], oi[key], key, rtol=1e-6)
    _same(g_rot_p, g_rot_i, "d loss / d rotation")
    _same(g_cen_p, g_cen_i, "d loss / d centre")
    assert float(g_rot_i.abs().max()) > 0 and float(g_cen_i.abs().max()) > 0
    _same(g_view_p, g_view_i, "d loss / d viewmats")
    _same(vsp_p, vsp_i, "viewspace grad")
    for a, b, p in zip(gp_p, gp_i, params):
        _same(a, b, f"parameter gradient {tuple(p.shape)}")


def test_render_blurry_view_ray_pose_shared_cameras():
    """K + 2 lists over K record sets: the centre camera's pose gradient also collects the dynamic-only and static-only
    renders' contributions (dec_rays_per_k = 2)."""
    from mobgs_b200.cameras import camera_rays, ray_pose
    from mobgs_b200.subframes import render_blurry_view
    K, W, H = 5, 96, 64
    sc, dc, intr = synthetic_scene(350, 250, W, H, seed=31, device="cuda")
    params = [p for pc in (sc, dc) for p in pc.parameters() if p.requires_grad]
    bg = torch.tensor([0.15, 0.3, 0.45], device="cuda")
    expo = (torch.linspace(-1, 1, K) * 0.4).cuda()
    half = K // 2
    cams = [make_camera(intr, subframe_w2c(k, K, device="cuda"), time=0.55) for k in range(K)]
    g = torch.Generator().manual_seed(5)
    tgt = torch.rand(3, H, W, generator=g).cuda()
    keys = ("depth", "d_alpha", "s_alpha", "s_render", "d_depth", "d_render")
    res = []
    for mode in ("image", "pose"):
        _, rot, cen = _leaves(K)
        geom = (intr.cx, intr.cy, intr.fx, intr.fy)
        rays = camera_rays(rot, cen, *geom, W, H) if mode == "image" else ray_pose(rot, cen, *geom)
        out = render_blurry_view(cams[half], cams, expo, sc, dc, None, bg, rays=rays)
        ((out["render"] - tgt).abs().mean() + sum(out[k].mean() * 0.2 for k in keys)).backward()
        res.append((out, rot.grad.clone(), cen.grad.clone(), _grads(params)))
    (oi, g_rot_i, g_cen_i, gp_i), (op, g_rot_p, g_cen_p, gp_p) = res
    for key in ("render", "render_center") + keys:
        _same(op[key], oi[key], key, rtol=1e-6)
    _same(g_rot_p, g_rot_i, "d loss / d rotation")
    _same(g_cen_p, g_cen_i, "d loss / d centre")
    for a, b, p in zip(gp_p, gp_i, params):
        _same(a, b, f"parameter gradient {tuple(p.shape)}")


def test_ray_pose_without_pose_gradient_and_slicing():
    """frozen pose (no v_pose_partial) and RayPose slicing (sub-frame sharding)."""
    from mobgs_b200.cameras import ray_pose_from_w2c
    from mobgs_b200.subframes import render_subframes
    K, W, H = 4, 80, 48
    sc, dc, intr = synthetic_scene(200, 100, W, H, seed=3, device="cuda")
    view = torch.stack([subframe_w2c(k, K) for k in range(K)]).cuda()
    rp = ray_pose_from_w2c(view, intr.fx, intr.fy, intr.cx, intr.cy)
    Kmat = make_camera(intr, subframe_w2c(0, K)).K.cuda()
    t = torch.full((K,), 0.5).cuda()
    bg = torch.zeros(3, device="cuda")
    full = render_subframes(sc, dc, view, Kmat, t, t, rp, bg, W, H)
    full["render"].mean().backward()
    part = render_subframes(sc, dc, view[1:3], Kmat, t[1:3], t[1:3], rp[1:3], bg, W, H)
    _same(part["subframes"], full["subframes"][1:3], "sliced sub-frames", rtol=1e-6)
    ref = render_subframes(sc, dc, view, Kmat, t, t, rp.rays(W, H), bg, W, H)
    _same(full["subframes"], ref["subframes"], "pose vs materialised rays", rtol=1e-6)
    rigid = ray_pose_from_w2c(view, intr.fx, intr.fy, intr.cx, intr.cy, rigid=True)
    _same(rigid.pose, rp.pose, "closed-form rigid inverse vs torch.inverse", rtol=1e-5)


def test_counting_pass_fused_into_projection_gives_identical_results():
    """ops.BinPlan: the projection launch counts and records the tile intersections of the lists the blend will walk
    (MobgsSynthFwd.bin_tile_counts); the lists — and therefore images and gradients — must be exactly those of the
    stand-alone counting kernels.  Covers K plain lists (render_subframes) and K + 2 lists with index ranges over shared
    record sets (render_blurry_view), and checks that the fused path was really taken."""
    from mobgs_b200 import ops
    from mobgs_b200.cameras import ray_pose_from_w2c
    from mobgs_b200.subframes import render_blurry_view, render_subframes
    K, W, H = 5, 112, 80
    sc, dc, intr = synthetic_scene(500, 300, W, H, seed=8, device="cuda")
    params = [p for pc in (sc, dc) for p in pc.parameters() if p.requires_grad]
    view = torch.stack([subframe_w2c(k, K) for k in range(K)]).cuda()
    rp = ray_pose_from_w2c(view, intr.fx, intr.fy, intr.cx, intr.cy, rigid=True)
    Kmat = make_camera(intr, subframe_w2c(0, K)).K.cuda()
    t = (0.5 + torch.linspace(-1, 1, K) * 0.4 / 23).cuda()
    bg = torch.tensor([0.1, 0.2, 0.3], device="cuda")
    cams = [make_camera(intr, subframe_w2c(k, K, device="cuda"), time=0.55) for k in range(K)]
    expo = (torch.linspace(-1, 1, K) * 0.4).cuda()
    tgt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(1)).cuda()

    def run(kind):
        for p in params:
            p.grad = None
        if kind == "subframes":
            out = render_subframes(sc, dc, view, Kmat, t.clamp(0, 1), t, rp, bg, W, H)
            keys = ("render", "depth", "alpha")
        else:
            out = render_blurry_view(cams[K // 2], cams, expo, sc, dc, None, bg, rays=rp)
            keys = ("render", "depth", "d_alpha", "s_alpha", "s_render", "d_render")
        (out["render"] - tgt).abs().mean().backward()
        return [out[k].detach().clone() for k in keys], [None if p.grad is None else p.grad.clone() for p in params]

    taken = []
    orig = ops.BinPlan.prepare

    def spy(self, N, dev):
        r = orig(self, N, dev)
        taken.append(r is not None)
        return r

    old = ops.FUSED_COUNT
    try:
        for kind in ("subframes", "blurry"):
            ops._CAP_CACHE.clear()
            ops.FUSED_COUNT = False
            run(kind)                                   # first call of the shape: sets the capacity guess
            ref_out, ref_g = run(kind)                  # stand-alone counting kernel (recorded entries)
            ops.FUSED_COUNT = True
            ops.BinPlan.prepare = spy
            got_out, got_g = run(kind)
            ops.BinPlan.prepare = orig
            assert taken and taken[-1], "the fused counting pass was not taken"
            for a, b in zip(got_out, ref_out):
                assert torch.equal(a, b), kind
            for a, b, p in zip(got_g, ref_g, params):
                assert (a is None) == (b is None)
                if a is not None and a.numel():
                    assert float((a - b).abs().max()) <= 1e-5 * (float(b.abs().max()) + 1e-12), (kind, tuple(p.shape))
    finally:
        ops.FUSED_COUNT = old
        ops.BinPlan.prepare = orig
