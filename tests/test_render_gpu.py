"""GPU parity of the drop-in renderer (mobgs_b200.gaussian_renderer: render / get_flow /
get_flow_static) against the CPU oracle restatement of the reference renderer
(oracle/mobgs_ref.py, itself pinned to the reference's own Python by
tests/test_reference_parity.py + tests/golden/).  BASELINE config 1 shape: 1k Gaussians
(700 static / 300 dynamic), 128x128."""
import pytest
import torch

from oracle import mobgs_ref as M
from mobgs_b200.scene import make_camera, subframe_w2c, synthetic_scene

from test_ops_gpu import _close

pytestmark = pytest.mark.gpu

PARAMS = ("_xyz", "_rotation", "_scaling", "_opacity", "_features_dc")
DPARAMS = ("control_xyz", "_rotation", "_omega", "_scaling", "_opacity", "_features_dc", "_features_t")


def _pair(ns=700, nd=300, W=128, H=128, seed=1234, k=0, K=1, time=0.5):
    so, do, intr = synthetic_scene(ns, nd, W, H, seed=seed)
    sc, dc, _ = synthetic_scene(ns, nd, W, H, seed=seed, device="cuda")
    dc.rgbdecoder.load_state_dict(do.rgbdecoder.state_dict())
    cam_o = make_camera(intr, subframe_w2c(k, K), time=time)
    cam_c = make_camera(intr, subframe_w2c(k, K, device="cuda"), time=time)
    return (so, do, cam_o), (sc, dc, cam_c)


def _weighted(out, keys, seed=0):
    g = torch.Generator().manual_seed(seed)
    tot = 0
    ws = {}
    for k in keys:
        ws[k] = torch.rand(out[k].shape, generator=g)
    return ws


def _loss(out, ws):
    return sum((out[k] * ws[k].to(out[k].device)).sum() for k in ws)


def _check_param_grads(sc, dc, so, do):
    for n in PARAMS:
        _close(getattr(sc, n).grad, getattr(so, n).grad, "stat" + n, max_outlier_frac=1e-3)
    for n in DPARAMS:
        _close(getattr(dc, n).grad, getattr(do, n).grad, "dyn" + n, max_outlier_frac=1e-3)
    for pc, po in zip(dc.rgbdecoder.parameters(), do.rgbdecoder.parameters()):
        _close(pc.grad, po.grad, "decoder", max_outlier_frac=1e-3)


@pytest.mark.parametrize("delta,flow", [(None, False), (0.3, True), (-0.45, False)])
def test_render_matches_oracle(delta, flow):
    from mobgs_b200.gaussian_renderer import render
    (so, do, cam_o), (sc, dc, cam_c) = _pair()
    bg = torch.tensor([0.2, 0.5, 0.7, 1.0])
    kw = dict(get_static=True, get_dynamic=True, delta_exposure=delta, get_flow=flow)
    out_c = render(cam_c, sc, dc, None, bg.cuda(), **kw)
    out_o = M.render_ref(cam_o, so, do, None, bg, **kw)
    assert set(out_c.keys()) == set(out_o.keys())
    img_keys = ["render", "s_render", "d_render", "depth", "d_depth", "d_alpha", "s_alpha", "s_depth"]
    if flow:
        img_keys += ["ori_flow", "ori_coord_map"]
    for k in img_keys:
        assert out_c[k].shape == out_o[k].shape, k
        _close(out_c[k], out_o[k], k, atol=2e-4 if "depth" in k or "flow" in k or "coord" in k else 1e-4,
               scale_atol=False, max_outlier_frac=1e-3)
    for k in out_o:
        if out_o[k] is None:
            assert out_c[k] is None, k
    assert out_c["viewspace_points"].shape == out_o["viewspace_points"].shape == (1, 1000, 2)
    agree = (out_c["radii"].cpu() > 0) == (out_o["radii"] > 0)
    assert agree.float().mean() > 0.999
    assert torch.equal(out_c["visibility_filter"].cpu(), out_c["radii"].cpu() > 0)
    _close(out_c["means_3d_final"], out_o["means_3d_final"], "means_3d_final", atol=1e-3, scale_atol=False)
    _close(out_c["means_3d"], out_o["means_3d"], "means_3d", atol=1e-5, scale_atol=False)

    ws = _weighted(out_o, img_keys[:7] + (["ori_flow"] if flow else []))
    _loss(out_c, ws).backward()
    _loss(out_o, ws).backward()
    _check_param_grads(sc, dc, so, do)
    _close(out_c["viewspace_points"].grad, out_o["viewspace_points"].grad, "viewspace_points.grad",
           max_outlier_frac=1e-3)


@pytest.mark.parametrize("frozen", [False, True])
def test_render_pose_gradient_eval_path(frozen):
    """eval.py:120-150 optimises the pose through render(..., w2c=...) only.  frozen: every Gaussian tensor has
    requires_grad = False as eval.py:246-255 sets them — the backward then runs in pose-only mode (no parameter
    gradient buffer, no parameter stores) and must give the same pose gradient."""
    from mobgs_b200.gaussian_renderer import render
    (so, do, cam_o), (sc, dc, cam_c) = _pair(ns=500, nd=200, W=96, H=64)
    if frozen:
        sc.requires_grad_(False)
        dc.requires_grad_(False)
        for p in dc.rgbdecoder.parameters():
            p.requires_grad_(False)
    w2c_o = subframe_w2c(1, 3).clone().requires_grad_(True)
    w2c_c = subframe_w2c(1, 3, device="cuda").clone().requires_grad_(True)
    bg = torch.zeros(3)
    out_c = render(cam_c, sc, dc, None, bg.cuda(), w2c=w2c_c)
    out_o = M.render_ref(cam_o, so, do, None, bg, w2c=w2c_o)
    _close(out_c["render"], out_o["render"], "render", scale_atol=False)
    ws = _weighted(out_o, ["render", "depth"])
    _loss(out_c, ws).backward()
    _loss(out_o, ws).backward()
    _close(w2c_c.grad[:3], w2c_o.grad[:3], "v_w2c", rtol=5e-3)
    if frozen:
        assert all(p.grad is None for p in sc.parameters()) and all(p.grad is None for p in dc.parameters())


@pytest.mark.parametrize("delta", [0.6, -1.0])
def test_get_flow_matches_oracle(delta):
    from mobgs_b200.gaussian_renderer import get_flow
    (so, do, cam_o), (sc, dc, cam_c) = _pair(ns=600, nd=400, W=112, H=80, time=0.4)
    bg = torch.tensor([0.0, 0.0, 0.0])
    got = get_flow(cam_c, sc, dc, None, bg.cuda(), delta_exposure=delta)
    ref = M.get_flow_ref(cam_o, so, do, None, bg, delta_exposure=delta)
    names = ["exp2mid", "mid2exp", "latent_img", "latent_alpha"]
    for a, b, n in zip(got, ref, names):
        assert a.shape == b.shape, n
        _close(a, b, n, atol=2e-4, scale_atol=False, max_outlier_frac=1e-3)
    g = torch.Generator().manual_seed(1)
    ws = [torch.rand(b.shape, generator=g) for b in ref]
    sum((a * w.cuda()).sum() for a, w in zip(got, ws)).backward()
    sum((b * w).sum() for b, w in zip(ref, ws)).backward()
    _check_param_grads(sc, dc, so, do)


def test_get_flow_static_matches_oracle():
    from mobgs_b200.gaussian_renderer import get_flow_static
    (so, do, cam_o), (sc, dc, cam_c) = _pair(ns=800, nd=10, W=96, H=96)
    intr = synthetic_scene(1, 1, 96, 96)[2]
    cams_o = [make_camera(intr, subframe_w2c(k, 5)) for k in (0, 4, 2)]
    cams_c = [make_camera(intr, subframe_w2c(k, 5, device="cuda")) for k in (0, 4, 2)]
    bg = torch.zeros(3)
    with torch.no_grad():
        f_c, r_c = get_flow_static(*cams_c, sc, dc, None, bg.cuda())
        f_o, r_o = M.get_flow_static_ref(*cams_o, so, do, None, bg)
    vis = (f_o.abs().sum(-1) > 0) & (f_c.cpu().abs().sum(-1) > 0)
    _close(f_c.cpu()[vis], f_o[vis], "flow_2d", atol=1e-3, scale_atol=False)
    _close(r_c, r_o, "rendered_flow", atol=2e-4, scale_atol=False, max_outlier_frac=1e-3)


def test_blur_model_subframe_mean():
    """K sub-frames through render() + the pixel mean of train.py:540-541, gradients included."""
    from mobgs_b200.gaussian_renderer import render
    K = 3
    so, do, intr = synthetic_scene(400, 200, 96, 64, seed=5)
    sc, dc, _ = synthetic_scene(400, 200, 96, 64, seed=5, device="cuda")
    dc.rgbdecoder.load_state_dict(do.rgbdecoder.state_dict())
    bg = torch.zeros(3)
    deltas = torch.linspace(-1, 1, K) * 0.4
    imgs_c, imgs_o = [], []
    for k in range(K):
        cam_o = make_camera(intr, subframe_w2c(k, K))
        cam_c = make_camera(intr, subframe_w2c(k, K, device="cuda"))
        imgs_c.append(render(cam_c, sc, dc, None, bg.cuda(), delta_exposure=deltas[k].item())["render"])
        imgs_o.append(M.render_ref(cam_o, so, do, None, bg, delta_exposure=deltas[k].item())["render"])
    pc, po = M.blur_mean(imgs_c), M.blur_mean(imgs_o)
    _close(pc, po, "blurred", scale_atol=False)
    tgt = torch.rand(po.shape, generator=torch.Generator().manual_seed(2))
    (pc - tgt.cuda()).abs().mean().backward()
    (po - tgt).abs().mean().backward()
    _check_param_grads(sc, dc, so, do)


def test_render_subframes_batched_equals_looped_oracle():
    """One K-batched launch chain == K oracle render() calls + the blur mean, gradients included
    (parameters, decoder, per-sub-frame view matrices, densification gradient of the centre)."""
    from mobgs_b200.subframes import render_subframes
    K, W, H = 5, 112, 80
    so, do, intr = synthetic_scene(500, 300, W, H, seed=8)
    sc, dc, _ = synthetic_scene(500, 300, W, H, seed=8, device="cuda")
    bg = torch.tensor([0.1, 0.2, 0.3])
    deltas = torch.linspace(-1, 1, K) * 0.4
    w2c_o = [subframe_w2c(k, K).clone().requires_grad_(True) for k in range(K)]
    cams_o = [make_camera(intr, w.detach()) for w in w2c_o]
    imgs, depths = [], []
    vsp_o = None
    for k in range(K):
        out = M.render_ref(cams_o[k], so, do, None, bg, w2c=w2c_o[k], delta_exposure=deltas[k].item())
        imgs.append(out["render"]); depths.append(out["depth"])
        if k == K // 2:
            vsp_o = out["viewspace_points"]
    pred_o = M.blur_mean(imgs)

    view_c = torch.stack([w.detach() for w in w2c_o]).cuda().requires_grad_(True)
    t0 = torch.full((K,), 0.5)
    t_poly = (t0 + deltas / 23).cuda()
    t_spline = t_poly.clamp(0, 1)
    rays = torch.cat([c.cam_ray for c in cams_o]).cuda()
    out_c = render_subframes(sc, dc, view_c, cams_o[0].K.cuda(), t_spline, t_poly, rays, bg.cuda(), W, H)
    _close(out_c["render"], pred_o, "blurred render", scale_atol=False)
    _close(out_c["depth"], torch.cat(depths), "depth", atol=2e-4, scale_atol=False, max_outlier_frac=1e-3)
    g = torch.Generator().manual_seed(3)
    tgt = torch.rand(pred_o.shape, generator=g)
    wd = torch.rand(K, H, W, generator=g) * 0.1
    ((out_c["render"] - tgt.cuda()).abs().mean() + (out_c["depth"] * wd.cuda()).mean()).backward()
    ((pred_o - tgt).abs().mean() + (torch.cat(depths) * wd).mean()).backward()
    _check_param_grads(sc, dc, so, do)
    _close(view_c.grad[:, :3], torch.stack([w.grad for w in w2c_o])[:, :3], "v_viewmats", rtol=5e-3)
    _close(out_c["viewspace_points"].grad, vsp_o.grad, "viewspace grad", max_outlier_frac=1e-3)


@pytest.mark.parametrize("ns,nd,W,H,K", [(0, 300, 50, 34, 2), (300, 0, 50, 34, 2), (123, 77, 97, 61, 9),
                                         (40, 40, 16, 16, 1), (5, 3, 200, 40, 4)])
def test_render_subframes_edge_shapes(ns, nd, W, H, K):
    """Empty static / dynamic sets, image sizes that are not tile multiples, K = 1..9."""
    from mobgs_b200.subframes import render_subframes
    so, do, intr = synthetic_scene(ns, nd, W, H, seed=ns + nd + K)
    sc, dc, _ = synthetic_scene(ns, nd, W, H, seed=ns + nd + K, device="cuda")
    bg = torch.tensor([0.3, 0.1, 0.6])
    deltas = (torch.linspace(-1, 1, K) * 0.4).tolist() if K > 1 else [0.25]
    cams = [make_camera(intr, subframe_w2c(k, K), time=0.6) for k in range(K)]
    imgs, deps = [], []
    for k in range(K):
        out = M.render_ref(cams[k], so, do, None, bg, delta_exposure=deltas[k])
        imgs.append(out["render"]); deps.append(out["depth"])
    pred_o = M.blur_mean(imgs)
    view = torch.stack([subframe_w2c(k, K) for k in range(K)]).cuda()
    t_poly = torch.tensor([0.6 + d / 23 for d in deltas]).cuda()
    rays = torch.cat([c.cam_ray for c in cams]).cuda()
    out_c = render_subframes(sc, dc, view, cams[0].K.cuda(), t_poly.clamp(0, 1), t_poly, rays, bg.cuda(), W, H)
    assert out_c["render"].shape == (3, H, W) and out_c["radii"].shape == (K, ns + nd)
    _close(out_c["render"], pred_o, "blurred render", scale_atol=False, max_outlier_frac=2e-3)
    _close(out_c["depth"], torch.cat(deps), "depth", atol=2e-4, scale_atol=False, max_outlier_frac=2e-3)
    tgt = torch.rand(pred_o.shape, generator=torch.Generator().manual_seed(0))
    (out_c["render"] - tgt.cuda()).abs().mean().backward()
    (pred_o - tgt).abs().mean().backward()
    if ns:
        for n in PARAMS:
            _close(getattr(sc, n).grad, getattr(so, n).grad, "stat" + n, max_outlier_frac=3e-3)
    if nd:
        for n in DPARAMS:
            _close(getattr(dc, n).grad, getattr(do, n).grad, "dyn" + n, max_outlier_frac=3e-3)


@pytest.mark.parametrize("two_pass", [False, True])
def test_render_blurry_view_equals_reference_loop(two_pass):
    """render_blurry_view == the reference's per-view loop (train.py:441, :497-541): centre render with
    get_static / get_dynamic + K-1 warped renders + blur mean; outputs and all gradients.
    two_pass: the gradients are accumulated by TWO backward passes over the same graph, as train.py does
    (photo_loss.backward(retain_graph=True) at :629, then loss.backward() of the regularisers at :680) — the second
    pass reaches only the centre render's depth / alphas, so most tiles take the backward kernel's zero-gradient
    early exit."""
    from mobgs_b200.subframes import render_blurry_view
    K, W, H = 5, 96, 64
    so, do, intr = synthetic_scene(350, 250, W, H, seed=31)
    sc, dc, _ = synthetic_scene(350, 250, W, H, seed=31, device="cuda")
    bg = torch.tensor([0.15, 0.3, 0.45])
    expo = torch.linspace(-1, 1, K) * 0.4
    half = K // 2
    base_o = make_camera(intr, subframe_w2c(half, K), time=0.55)
    warped_o = [make_camera(intr, subframe_w2c(k, K) @ subframe_w2c(1, 3), time=0.55) for k in range(K)]
    base_c = make_camera(intr, subframe_w2c(half, K, device="cuda"), time=0.55)
    warped_c = [make_camera(intr, (subframe_w2c(k, K) @ subframe_w2c(1, 3)).cuda(), time=0.55) for k in range(K)]

    pkg = M.render_ref(base_o, so, do, None, bg, get_static=True, get_dynamic=True)
    imgs = []
    for k in range(K):
        imgs.append(pkg["render"] if k == half else
                    M.render_ref(warped_o[k], so, do, None, bg, get_static=True, get_dynamic=True,
                                 delta_exposure=expo[k])["render"])
    pred_o = M.blur_mean(imgs)

    out = render_blurry_view(base_c, warped_c, expo.cuda(), sc, dc, None, bg.cuda())
    _close(out["render"], pred_o, "blurred", scale_atol=False)
    for key in ("depth", "s_render", "d_render", "d_depth", "d_alpha", "s_alpha", "s_depth"):
        assert out[key].shape == pkg[key].shape, key
        _close(out[key], pkg[key], key, atol=2e-4, scale_atol=False, max_outlier_frac=1e-3)
    _close(out["render_center"], pkg["render"], "centre render", scale_atol=False)
    g = torch.Generator().manual_seed(5)
    tgt = torch.rand(pred_o.shape, generator=g)
    ws = {k: torch.rand(pkg[k].shape, generator=g) * 0.2 for k in ("depth", "d_alpha", "s_alpha", "s_render", "d_depth")}
    if two_pass:
        (out["render"] - tgt.cuda()).abs().mean().backward(retain_graph=True)
        vsp_photo = out["viewspace_points"].grad.clone()
        sum((out[k] * w.cuda()).mean() for k, w in ws.items()).backward()
        (pred_o - tgt).abs().mean().backward(retain_graph=True)
        vsp_photo_o = pkg["viewspace_points"].grad.clone()
        sum((pkg[k] * w).mean() for k, w in ws.items()).backward()
        _close(vsp_photo, vsp_photo_o, "viewspace grad after the photometric pass", max_outlier_frac=1e-3)
    else:
        ((out["render"] - tgt.cuda()).abs().mean() + sum((out[k] * w.cuda()).mean() for k, w in ws.items())).backward()
        ((pred_o - tgt).abs().mean() + sum((pkg[k] * w).mean() for k, w in ws.items())).backward()
    _check_param_grads(sc, dc, so, do)
    _close(out["viewspace_points"].grad, pkg["viewspace_points"].grad, "viewspace grad", max_outlier_frac=1e-3)


@pytest.mark.parametrize("K", [5, 9, 1])
def test_get_flow_batched_equals_k_reference_calls(K):
    """get_flow_batched == K oracle get_flow calls (train.py:563-579), outputs and gradients.  K = 5: all ten
    mid2exp flow channels in one walk; K = 9 (the reference's num_warp): two walks over one shared binning."""
    from mobgs_b200.gaussian_renderer import get_flow_batched
    (so, do, cam_o), (sc, dc, cam_c) = _pair(ns=500, nd=400, W=96, H=64, time=0.45)
    bg = torch.tensor([0.0, 0.0, 0.0])
    half = K // 2
    deltas = [(k - half) / max(half, 1) for k in range(K)] if K > 1 else [0.7]
    ref = [M.get_flow_ref(cam_o, so, do, None, bg, delta_exposure=d) for d in deltas]
    got = get_flow_batched(cam_c, sc, dc, None, bg.cuda(), deltas)
    want = [torch.cat([r[0] for r in ref]), torch.cat([r[1] for r in ref]),
            torch.stack([r[2] for r in ref]), torch.cat([r[3] for r in ref])]
    for a, b, n in zip(got, want, ["exp2mid", "mid2exp", "latent_img", "latent_alpha"]):
        assert a.shape == b.shape, (n, a.shape, b.shape)
        _close(a, b, n, atol=2e-4, scale_atol=False, max_outlier_frac=1e-3)
    g = torch.Generator().manual_seed(2)
    ws = [torch.rand(b.shape, generator=g) for b in want]
    sum((a * w.cuda()).sum() for a, w in zip(got, ws)).backward()
    sum((b * w).sum() for b, w in zip(want, ws)).backward()
    _check_param_grads(sc, dc, so, do)
