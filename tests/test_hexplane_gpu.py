"""a11: fused HexPlane + MLP forward (tcgen05, 3xTF32) vs the reference module's golden outputs
and vs the oracle restatement at a larger size.  Tolerance 1e-4 abs / 1e-3 rel (fp32 bar)."""
import numpy as np
import pytest
import torch

from test_oracle import _hexplane_args, load_hexplane_golden

pytestmark = pytest.mark.gpu


def _cmp(got, want, name, atol=1e-4, rtol=1e-3):
    got = got.detach().cpu().double()
    want = torch.as_tensor(want).detach().cpu().double()
    err = (got - want).abs()
    bad = err > atol + rtol * want.abs()
    assert not bad.any(), (name, float(err.max()), int(bad.sum()))


def test_fused_forward_matches_reference_golden():
    net, gold = load_hexplane_golden("cuda")
    ins = [torch.from_numpy(gold[k]).cuda() for k in ("in_pts", "in_scales", "in_rots", "in_t")]
    with torch.no_grad():
        p, s, r = net(*ins)
    _cmp(p, gold["out_pts"], "out_pts")
    _cmp(s, gold["out_scales"], "out_scales")
    _cmp(r, gold["out_rots"], "out_rots")


@pytest.mark.parametrize("n,base_res", [(5000, 64), (128, 16), (1, 16), (129, 32)])
def test_fused_forward_matches_oracle(n, base_res):
    from mobgs_b200.deformation import HexPlaneMLP
    from oracle.hexplane_ref import deform_forward_ref
    torch.manual_seed(n)
    net = HexPlaneMLP(_hexplane_args(base_res))
    with torch.no_grad():
        for name, p in net.named_parameters():
            if p.dim() == 1:
                p.uniform_(-0.2, 0.2)
            elif "grids" in name:
                p.add_(0.1 * torch.randn_like(p))
            elif p.requires_grad:
                torch.nn.init.xavier_uniform_(p)
    net.set_aabb([1.3, 1.1, 1.2], [-1.2, -1.0, -1.4])
    pts = torch.rand(n, 3) * 3 - 1.5
    scales = torch.randn(n, 3) * 0.3 - 3
    rots = torch.randn(n, 4)
    t = torch.rand(n, 1)
    with torch.no_grad():
        want = deform_forward_ref(net, pts, scales, rots, t)
        net.cuda()
        got = net(pts.cuda(), scales.cuda(), rots.cuda(), t.cuda())
    for g, w, name in zip(got, want, ("pts", "scales", "rots")):
        _cmp(g, w, name)


def test_gradients_match_reference_golden():
    """d loss / d (inputs, planes, MLP weights) through the fused forward vs the reference module's own
    autograd (golden `grad_pts`, `pgrad::*`)."""
    net, gold = load_hexplane_golden("cuda")
    ins = [torch.from_numpy(gold[k]).cuda() for k in ("in_pts", "in_scales", "in_rots", "in_t")]
    ins[0].requires_grad_(True)
    p, s, r = net(*ins)
    _cmp(p, gold["out_pts"], "out_pts")
    w = torch.from_numpy(gold["loss_w"]).cuda()
    ((p * w[:, :3]).sum() + (s * w[:, 3:6]).sum() + (r * w[:, 6:]).sum()).backward()
    _cmp(ins[0].grad, gold["grad_pts"], "grad_pts", atol=2e-4)
    own = dict(net.named_parameters())
    checked = 0
    for k in gold.files:
        if k.startswith("pgrad::") and k[7:] in own and own[k[7:]].grad is not None:
            g = gold[k]
            _cmp(own[k[7:]].grad, g, k, atol=1e-4 * max(1.0, float(np.abs(g).max())))
            checked += 1
    assert checked >= 20


def test_native_feature_gather_and_vjp_match_grid_sample():
    """csrc/hexplane_grid.cu vs the oracle's F.grid_sample restatement: features, plane gradients,
    coordinate gradients (incl. points outside the AABB and on the border)."""
    from mobgs_b200.deformation import HexPlaneMLP, hexplane_features, hexplane_features_vjp
    from oracle.hexplane_ref import hexplane_features as ref_features
    torch.manual_seed(3)
    net = HexPlaneMLP(_hexplane_args(16))
    with torch.no_grad():
        for name, p in net.named_parameters():
            if "grids" in name:
                p.add_(0.2 * torch.randn_like(p))
    net.set_aabb([1.1, 1.3, 0.9], [-1.0, -1.2, -1.1])
    n = 1000
    pts = (torch.rand(n, 3) * 3 - 1.5).requires_grad_(True)
    t = torch.rand(n, 1)
    t[:4, 0] = torch.tensor([0.0, 1.0, 1.2, -0.1])
    grids = [[p for p in level] for level in net.deformation_net.grid.grids]
    aabb = net.deformation_net.grid.aabb
    feat0 = ref_features(pts, t, aabb, grids)
    w = torch.randn(feat0.shape)
    (feat0 * w).sum().backward()
    net.cuda()
    cg = [[p for p in level] for level in net.deformation_net.grid.grids]
    feat = hexplane_features(pts.detach().cuda(), t.cuda(), net.deformation_net.grid.aabb, cg)
    _cmp(feat, feat0.detach(), "feat", atol=2e-6, rtol=1e-5)
    g_planes, g_pts, _ = hexplane_features_vjp(pts.detach().cuda(), t.cuda(), net.deformation_net.grid.aabb, cg, w.cuda())
    _cmp(g_pts, pts.grad, "g_pts", atol=1e-4 * float(pts.grad.abs().max()))
    flat0 = [p for level in grids for p in level]
    for i, (g, p0) in enumerate(zip(g_planes, flat0)):
        _cmp(g, p0.grad, f"g_plane{i}", atol=1e-5 * max(1.0, float(p0.grad.abs().max())))


def test_fused_adam_step_invalidates_the_packed_operand_cache():
    """FusedAdam writes parameters through raw pointers; the packed TF32 weights /
    channels-last planes of the fused forward are cached on (tensor, version), so the step must bump the
    version.  forward -> FusedAdam.step -> forward has to see the new weights (compared with the oracle
    on the updated module)."""
    from mobgs_b200.deformation import HexPlaneMLP
    from mobgs_b200.optim import FusedAdam
    from oracle.hexplane_ref import deform_forward_ref
    torch.manual_seed(11)
    net = HexPlaneMLP(_hexplane_args(16)).cuda()
    net.set_aabb([1.3, 1.1, 1.2], [-1.2, -1.0, -1.4])
    n = 300
    ins = [torch.rand(n, 3, device="cuda") * 2 - 1, torch.randn(n, 3, device="cuda") * 0.3 - 3,
           torch.randn(n, 4, device="cuda"), torch.rand(n, 1, device="cuda")]
    params = [p for p in net.parameters() if p.requires_grad]
    opt = FusedAdam(params, lr=5e-2, eps=1e-15)
    v0 = [p._version for p in params]
    p0, s0, r0 = net(*ins)
    (p0.sum() + s0.sum() + r0.sum()).backward()
    opt.step()
    assert all(p._version > v for p, v in zip(params, v0) if p.grad is not None)
    with torch.no_grad():
        got = net(*ins)
        ref_net = HexPlaneMLP(_hexplane_args(16))
        ref_net.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
        ref_net.set_aabb([1.3, 1.1, 1.2], [-1.2, -1.0, -1.4])
        want = deform_forward_ref(ref_net, *[t.cpu() for t in ins])
    assert (got[0] - p0.detach()).abs().max() > 1e-3          # the step moved the output ...
    for g, w, name in zip(got, want, ("pts", "scales", "rots")):
        _cmp(g, w, name)                                      # ... to what the updated weights give
