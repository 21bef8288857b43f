"""f4: densify / prune gather-compaction.

  * CPU, needs /root/reference: oracle.densify_ref == the reference's own GaussianModel._prune_optimizer /
    cat_tensors_to_optimizer executed on a real GaussianModel (bit-exact) — pins the oracle.
  * GPU: mobgs_b200.densify.{prune_optimizer, cat_tensors_to_optimizer} (one mobgs_compact_rows launch) ==
    the oracle, bit for bit, on the reference's 17-group optimiser layout incl. the int64 current_control_num,
    with edge cases (keep all / keep none / one row)."""
import copy
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_env as E  # noqa: E402

# (group name, trailing shape, dtype, trainable) — scene/gaussian_model.py:598-618 (deformation / grid / decoder
# groups hold several tensors and are skipped by the reference's surgery, :1047)
GROUPS = [("xyz", (3,), torch.float32, True), ("control_xyz", (12, 3), torch.float32, True),
          ("current_control_num", (1,), torch.int64, False), ("f_dc", (6,), torch.float32, True),
          ("f_rest", (16, 3), torch.float32, True), ("f_t", (3,), torch.float32, True), ("opacity", (1,), torch.float32, True),
          ("scaling", (3,), torch.float32, True), ("rotation", (4,), torch.float32, True), ("omega", (4,), torch.float32, True),
          ("zeta", (1,), torch.float32, True), ("trbf_center", (1,), torch.float32, True), ("trbf_scale", (1,), torch.float32, True),
          ("motion", (9,), torch.float32, True)]


def _make_optimizer(n, device, seed=0, with_state=True):
    g = torch.Generator().manual_seed(seed)
    groups = []
    for name, shape, dt, train in GROUPS:
        if dt == torch.int64:
            p = torch.nn.Parameter(torch.randint(4, 13, (n,) + shape, generator=g).to(device), requires_grad=False)
        else:
            p = torch.nn.Parameter(torch.randn((n,) + shape, generator=g).to(device))
        groups.append({"params": [p], "lr": 1e-3, "name": name})
    dec = [torch.nn.Parameter(torch.randn(6, 12, generator=g).to(device)), torch.nn.Parameter(torch.randn(3, 6, generator=g).to(device))]
    groups.append({"params": dec, "lr": 1e-4, "name": "decoder"})
    opt = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    if with_state:
        for grp in opt.param_groups:
            for p in grp["params"]:
                if p.requires_grad and grp["name"] not in ("f_rest", "motion"):      # some groups never get a gradient
                    p.grad = torch.randn(p.shape, generator=g).to(device)
        opt.step()
        opt.zero_grad(set_to_none=True)
    return opt


def _to_cuda(opt):
    """a CUDA twin of a CPU optimiser: same group layout, bit-identical parameters and Adam moments (stepping the
    two on their own devices would not give identical bits)"""
    groups, pairs = [], []
    for g in opt.param_groups:
        ps = [torch.nn.Parameter(p.detach().cuda(), requires_grad=p.requires_grad) for p in g["params"]]
        pairs += list(zip(g["params"], ps))
        groups.append({"params": ps, "lr": g["lr"], "name": g["name"]})
    twin = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    for p_cpu, p_gpu in pairs:
        st = opt.state.get(p_cpu)
        if st:
            twin.state[p_gpu] = {"step": st["step"].clone(), "exp_avg": st["exp_avg"].cuda(), "exp_avg_sq": st["exp_avg_sq"].cuda()}
    return twin


def _same(a_opt, b_opt, a_out, b_out):
    assert list(a_out) == list(b_out)
    for k in a_out:
        assert a_out[k].dtype == b_out[k].dtype and a_out[k].shape == b_out[k].shape, k
        assert a_out[k].requires_grad == b_out[k].requires_grad, k
        assert torch.equal(a_out[k].detach().cpu(), b_out[k].detach().cpu()), k
    for ga, gb in zip(a_opt.param_groups, b_opt.param_groups):
        assert ga["name"] == gb["name"]
        for pa, pb in zip(ga["params"], gb["params"]):
            sa, sb = a_opt.state.get(pa), b_opt.state.get(pb)
            assert (sa is None or len(sa) == 0) == (sb is None or len(sb) == 0), ga["name"]
            if sa:
                for key in ("exp_avg", "exp_avg_sq"):
                    assert torch.equal(sa[key].cpu(), sb[key].cpu()), (ga["name"], key)


@pytest.mark.skipif(not E.reference_available(), reason="/root/reference not present on this machine")
def test_oracle_matches_reference_optimizer_surgery():
    from oracle import densify_ref as O
    E.setup_paths()
    with E.cuda_to_cpu():
        from arguments import OptimizationParams

        def fresh():
            """two real GaussianModels with a populated Adam state — deterministic, so calling it twice gives twins"""
            torch.manual_seed(0)          # create_from_pcd draws the static control points from the global RNG
            stat, dyn, _, _ = E.synthetic_reference_scene(n_static=120, n_dynamic=90)
            opt_args = E.group_args(OptimizationParams)
            g = torch.Generator().manual_seed(11)
            for pc in (stat, dyn):
                pc.training_setup(opt_args)
                for grp in pc.optimizer.param_groups:
                    grp["lr"] = 1e-3
                    for p in grp["params"]:
                        if p.requires_grad and p.dtype == torch.float32 and grp["name"] in (
                                "xyz", "control_xyz", "f_dc", "opacity", "scaling", "rotation", "omega", "f_t"):
                            p.grad = torch.randn(p.shape, generator=g)
                pc.optimizer.step()
                pc.optimizer.zero_grad(set_to_none=True)
            return stat, dyn

        def twins(which):
            """(a, b): two models with bit-identical parameters and Adam state (b is overwritten from a: the
            dynamic model's lstsq spline fit is not run-to-run reproducible)"""
            a, b = fresh()[which], fresh()[which]
            with torch.no_grad():
                for ga, gb in zip(a.optimizer.param_groups, b.optimizer.param_groups):
                    for pa, pb in zip(ga["params"], gb["params"]):
                        pb.copy_(pa)
                        sa, sb = a.optimizer.state.get(pa), b.optimizer.state.get(pb)
                        if sa:
                            for key in ("exp_avg", "exp_avg_sq"):
                                sb[key].copy_(sa[key])
            return a, b

        for which in (0, 1):
            g = torch.Generator().manual_seed(1)
            n = fresh()[which].get_xyz.shape[0]
            mask = torch.rand(n, generator=g) > 0.35
            ref_pc, ora_pc = twins(which)
            out_ref = ref_pc._prune_optimizer(mask)                 # the reference method, unmodified
            out_ora = O.prune_optimizer(ora_pc.optimizer, mask)
            _same(ref_pc.optimizer, ora_pc.optimizer, out_ref, out_ora)
            # append: extension rows for every single-parameter group
            ref_pc, ora_pc = twins(which)
            ext = {grp["name"]: (torch.randint(4, 13, (7,) + tuple(grp["params"][0].shape[1:]), generator=g)
                                 if grp["params"][0].dtype == torch.int64 else
                                 torch.randn((7,) + tuple(grp["params"][0].shape[1:]), generator=g))
                   for grp in ref_pc.optimizer.param_groups if len(grp["params"]) == 1}
            out_ref = ref_pc.cat_tensors_to_optimizer(ext)
            out_ora = O.cat_tensors_to_optimizer(ora_pc.optimizer, ext)
            _same(ref_pc.optimizer, ora_pc.optimizer, out_ref, out_ora)


@pytest.mark.gpu
@pytest.mark.parametrize("n,keep", [(5000, 0.6), (1, 1.0), (257, 0.0), (300, 1.0), (100_000, 0.9)])
def test_prune_optimizer_matches_oracle_bit_exact(n, keep):
    from mobgs_b200.densify import prune_optimizer
    from oracle import densify_ref as O
    cpu = _make_optimizer(n, "cpu", seed=n)
    gpu = _to_cuda(cpu)
    mask = torch.rand(n, generator=torch.Generator().manual_seed(7)) < keep
    stats = [torch.rand(n, 1), torch.rand(n), torch.rand(n, 3)]          # xyz_gradient_accum / max_radii2D / _deformation_accum
    want = O.prune_optimizer(cpu, mask)
    got, extra = prune_optimizer(gpu, mask.cuda(), extra=[s.cuda() for s in stats])
    _same(cpu, gpu, want, got)
    for s, e in zip(stats, extra):
        assert torch.equal(s[mask], e.cpu())
    # the rebuilt optimiser keeps stepping (state re-keyed to the new Parameters)
    for grp in gpu.param_groups:
        for p in grp["params"]:
            if p.requires_grad:
                p.grad = torch.ones_like(p)
    gpu.step()


@pytest.mark.gpu
@pytest.mark.parametrize("n,n_new", [(4000, 900), (10, 0), (0, 5)])
def test_cat_tensors_to_optimizer_matches_oracle_bit_exact(n, n_new):
    from mobgs_b200.densify import cat_tensors_to_optimizer
    from oracle import densify_ref as O
    cpu = _make_optimizer(n, "cpu", seed=n + 1, with_state=n > 0)
    gpu = _to_cuda(cpu)
    g = torch.Generator().manual_seed(3)
    ext = {name: (torch.randint(4, 13, (n_new,) + shape, generator=g) if dt == torch.int64 else torch.randn((n_new,) + shape, generator=g))
           for name, shape, dt, _ in GROUPS}
    want = O.cat_tensors_to_optimizer(cpu, ext)
    got = cat_tensors_to_optimizer(gpu, {k: v.cuda() for k, v in ext.items()})
    _same(cpu, gpu, want, got)
