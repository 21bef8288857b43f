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


@pytest.mark.parametrize("ncol,transpose,n_pts", [(128, False, 1000), (96, False, 257), (16, True, 5000), (64, False, 128)])
def test_wgrad_split_k_gemm_matches_fp64(ncol, transpose, n_pts):
    """mobgs_hexplane_wgrad alone: C[m][n] += sum_k A[m][k] B[n][k] (3xTF32 on tcgen05, split-K with TMEM-resident
    accumulators), row sums as bias gradients — against a float64 matmul; 1e-5 relative to the result's scale."""
    from mobgs_b200 import _lib as L
    g = torch.Generator().manual_seed(ncol + n_pts)
    ld = (n_pts + 127) // 128 * 128
    A = torch.zeros(128, ld)
    B = torch.zeros(ncol, ld)
    A[:, :n_pts] = torch.randn(128, n_pts, generator=g)
    B[:, :n_pts] = torch.randn(ncol, n_pts, generator=g)
    want = A.double() @ B.double().t()
    Ac, Bc = A.cuda(), B.cuda()
    C = torch.zeros((ncol, 128) if transpose else (128, ncol), device="cuda")
    bias_a, bias_b = torch.zeros(128, device="cuda"), torch.zeros(ncol, device="cuda")
    wg = L.HexWgrad()
    wg.n_problems, wg.ld = 2, ld
    for i, (bias, bfrom) in enumerate(((bias_a, 1), (bias_b, 2))):
        wg.A[i], wg.B[i], wg.n_cols[i] = Ac.data_ptr(), Bc.data_ptr(), ncol
        wg.C[i], wg.ldc[i], wg.transpose_out[i] = C.data_ptr(), (128 if transpose else ncol), int(transpose)
        wg.bias[i], wg.bias_from[i] = bias.data_ptr(), bfrom
    L.call("mobgs_hexplane_wgrad", wg, L.current_stream())
    torch.cuda.synchronize()
    got = (C.t() if transpose else C).cpu().double() / 2          # both problems accumulate into the same C
    scale = float(want.abs().max())
    assert (got - want).abs().max() <= 1e-5 * scale, float((got - want).abs().max() / scale)
    assert (bias_a.cpu().double() - A.double().sum(1)).abs().max() <= 1e-4 * float(A.double().sum(1).abs().max())
    assert (bias_b.cpu().double() - B.double().sum(1)).abs().max() <= 1e-4 * float(B.double().sum(1).abs().max())


@pytest.mark.parametrize("n,base_res", [(3000, 16), (129, 16), (1, 16)])
def test_tcgen05_backward_matches_oracle_autograd(n, base_res):
    """Every gradient of the fused module (inputs, planes, all MLP weights and biases) through the tcgen05 backward
    (mobgs_hexplane_mlp_bwd + mobgs_hexplane_wgrad + the native plane scatter) against autograd of the oracle
    restatement on the CPU; 1e-4 of the tensor's max / 1e-3 relative."""
    from mobgs_b200.deformation import HexPlaneMLP
    from oracle.hexplane_ref import deform_forward_ref
    torch.manual_seed(n + 5)
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
    pts = (torch.rand(n, 3) * 3 - 1.5).requires_grad_(True)
    scales = (torch.randn(n, 3) * 0.3 - 3).requires_grad_(True)
    rots = torch.randn(n, 4).requires_grad_(True)
    t = torch.rand(n, 1)
    w = [torch.randn(n, 3), torch.randn(n, 3), torch.randn(n, 4)]
    outs = deform_forward_ref(net, pts, scales, rots, t)
    sum((o * wi).sum() for o, wi in zip(outs, w)).backward()
    ref = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    ref_in = [pts.grad.clone(), scales.grad.clone(), rots.grad.clone()]
    net.zero_grad(set_to_none=True)
    net.cuda()
    ins = [x.detach().cuda().requires_grad_(True) for x in (pts, scales, rots)]
    got = net(ins[0], ins[1], ins[2], t.cuda())
    sum((o * wi.cuda()).sum() for o, wi in zip(got, w)).backward()
    for g, r, name in zip(ins, ref_in, ("pts", "scales", "rots")):
        _cmp(g.grad, r, "grad_" + name, atol=1e-4 * max(1.0, float(r.abs().max())))
    checked = 0
    for k, p in net.named_parameters():
        if k in ref:
            assert p.grad is not None, k
            # plane gradients are fp32 atomic sums over up to thousands of points (order not reproducible): 3e-4
            tol = (3e-4 if "grids" in k else 1e-4) * max(1.0, float(ref[k].abs().max()))
            if "grids" in k:
                _cmp(p.grad, ref[k], k, atol=tol)
            else:
                # A ReLU whose pre-activation lies within rounding of zero may gate differently in two correct fp32
                # implementations (3xTF32 tensor-core sums vs the CPU's SGEMM); one such flip at (point, unit o)
                # perturbs row o of that layer's weight gradient and element o of its bias gradient by one point's
                # contribution.  Allow at most two such rows / elements per tensor; everything else must hold 1e-4.
                err = (p.grad.detach().cpu().double() - ref[k].double()).abs()
                bad = err > tol + 1e-3 * ref[k].double().abs()
                rows = bad.reshape(bad.shape[0], -1).any(1)
                assert int(rows.sum()) <= 2, (k, float(err.max()), int(bad.sum()), int(rows.sum()))
            checked += 1
    assert checked >= 30


# ---- the torch formulation of the operand layouts (what the kernels' headers specify), as the checker of the
# ---- device-side packing launch (mobgs_pack_operands)
def _tile_ref(w):
    """[rows, K] row-major -> {hi, lo} x [K/4][rows][4]"""
    rows, K = w.shape
    hi = (w.contiguous().view(torch.int32) & -8192).view(torch.float32)     # keep the 10 mantissa bits of tf32
    lo = w - hi
    t = lambda m: m.reshape(rows, K // 4, 4).permute(1, 0, 2).contiguous().reshape(-1)  # noqa: E731
    return torch.cat([t(hi), t(lo)])


def _pad16(m):
    out = torch.zeros(16, *m.shape[1:], device=m.device)
    out[:m.shape[0]] = m
    return out


@pytest.mark.parametrize("base_res,levels", [(16, 3), (32, 2), (8, 1)])
def test_device_operand_pack_is_bit_exact(base_res, levels):
    from mobgs_b200 import deformation as D
    args = _hexplane_args(base_res)
    args.multires = [1, 2, 4][:levels]
    torch.manual_seed(base_res + levels)
    net = D.HexPlaneMLP(args).cuda()
    with torch.no_grad():
        for p in net.parameters():
            if p.requires_grad:
                p.copy_(torch.randn_like(p))
    d = net.deformation_net
    heads = [(s[1].weight, s[1].bias, s[3].weight, s[3].bias) for s in (d.pos_deform, d.scales_deform, d.rotations_deform)]
    planes = [[p for p in level] for level in d.grid.grids]
    w0, b0 = d.feature_out[0].weight, d.feature_out[0].bias
    D._PACK_CACHE.clear()
    pack = D._operand_pack(planes, w0, b0, heads)

    def check():
        K0 = w0.shape[1]
        want = {"w0": torch.cat([_tile_ref(w0.detach()[h * 64:(h + 1) * 64]) for h in range(2)]),
                "b0": b0.detach(),
                "wa": torch.cat([_tile_ref(Wa.detach()[h * 64:(h + 1) * 64]) for Wa, _, _, _ in heads for h in range(2)]),
                "ba": torch.stack([ba.detach() for _, ba, _, _ in heads]),
                "wb": torch.cat([_tile_ref(_pad16(Wb.detach())) for _, _, Wb, _ in heads]),
                "bb": torch.cat([_pad16(bb.detach()) for _, _, _, bb in heads])}
        w0t = w0.detach().t().contiguous()
        want["w0_t"] = torch.cat([_tile_ref(w0t[:min(64, K0)])] + ([_tile_ref(w0t[64:])] if K0 > 64 else []))
        want["wa_t"] = torch.cat([_tile_ref(Wa.detach().t().contiguous()[h * 64:(h + 1) * 64])
                                  for Wa, _, _, _ in heads for h in range(2)])
        want["wb_t"] = torch.cat([_tile_ref(_pad16(Wb.detach()).t().contiguous()[h * 64:(h + 1) * 64])
                                  for _, _, Wb, _ in heads for h in range(2)])
        for name, w in want.items():
            got = getattr(pack, name)
            assert got.numel() == w.numel(), (name, got.shape, w.shape)
            assert torch.equal(got.reshape(-1).view(torch.int32), w.reshape(-1).view(torch.int32)), name
        flat = [p for level in planes for p in level]
        assert len(pack.cl) == len(flat)
        for cl, g in zip(pack.cl, flat):
            assert torch.equal(cl, g.detach()[0].permute(1, 2, 0).contiguous())

    check()
    # a raw-pointer parameter update (FusedAdam) bumps the version counters: the next use repacks, same plan object
    from mobgs_b200.optim import FusedAdam
    params = [p for p in net.parameters() if p.requires_grad]
    for p in params:
        p.grad = torch.randn_like(p)
    FusedAdam([{"params": params, "lr": 1e-2, "name": "deformation"}], lr=0.0, eps=1e-15).step()
    assert D._operand_pack(planes, w0, b0, heads) is pack
    check()
    # replacing a parameter's storage invalidates the plan
    with torch.no_grad():
        d.feature_out[0].weight.data = d.feature_out[0].weight.data.clone()
    pack2 = D._operand_pack(planes, d.feature_out[0].weight, b0, heads)
    assert pack2 is not pack
    pack = pack2
    w0 = d.feature_out[0].weight
    check()
