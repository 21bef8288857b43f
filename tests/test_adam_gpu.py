"""f4 parity: mobgs_b200.optim.FusedAdam (one mobgs_adam_step launch) against torch.optim.Adam — the
optimiser the reference builds at scene/gaussian_model.py:641 — and against oracle/adam_ref.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 2e-6                    # fp32 tolerance of this row, written here: 2e-6 relative + 2 ulp of the largest
ULP2 = 2.4e-7                  # operand of the update (cancellation in m + w (g - m) and p - step_size * q)


def _close(x, y, scale, what=""):
    err = (x - y).abs()
    ok = err <= RTOL * y.abs() + ULP2 * scale
    assert bool(ok.all()), (what, float(err.max()), scale)


def _groups(gen, dev, n_dyn=5003, n_stat=7001):
    """Parameter groups shaped like the reference's (names, per-group lr, one frozen group, one without grad)."""
    def P(*s):
        return torch.randn(*s, generator=gen).to(dev).requires_grad_(True)
    return [
        {"params": [P(n_stat, 3)], "lr": 1.6e-4, "name": "xyz"},
        {"params": [P(n_dyn, 12, 3)], "lr": 1.6e-3, "name": "control_xyz"},
        {"params": [P(n_dyn, 6)], "lr": 2.5e-3, "name": "f_dc"},
        {"params": [P(n_dyn, 3)], "lr": 2.5e-3, "name": "f_t"},
        {"params": [P(n_dyn, 1)], "lr": 0.05, "name": "opacity"},
        {"params": [P(n_dyn, 3)], "lr": 5e-3, "name": "scaling"},
        {"params": [P(n_dyn, 4)], "lr": 1e-3, "name": "rotation"},
        {"params": [P(n_dyn, 4)], "lr": 1e-4, "name": "omega"},
        {"params": [P(n_dyn, 1)], "lr": 0.0, "name": "trbf_center"},
        {"params": [P(6, 12), P(3, 6)], "lr": 1e-4, "name": "decoder"},
        {"params": [P(70001)], "lr": 1e-3, "name": "long_odd"},          # > one chunk, not a multiple of 4
        {"params": [P(17)], "lr": 1e-3, "name": "never_gets_a_grad"},
    ]


def _clone_groups(groups):
    return [{**g, "params": [p.detach().clone().requires_grad_(True) for p in g["params"]]} for g in groups]


def _set_grads(groups_a, groups_b, gen, step):
    for ga, gb in zip(groups_a, groups_b):
        if ga["name"] == "never_gets_a_grad":
            continue
        for pa, pb in zip(ga["params"], gb["params"]):
            scale = 10.0 ** float(torch.randint(-7, 1, (1,), generator=gen))
            g = (torch.randn(pa.shape, generator=gen) * scale).to(pa.device)
            if step == 2 and ga["name"] == "f_t":
                g.zero_()                           # all-zero gradient: eps = 1e-15 must not produce NaN
            pa.grad, pb.grad = g.clone(), g.clone()


def test_fused_adam_matches_torch_adam():
    from mobgs_b200.optim import FusedAdam
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(0)
    ga = _groups(gen, dev)
    gb = _clone_groups(ga)
    ours = FusedAdam(ga, lr=0.0, eps=1e-15)
    ref = torch.optim.Adam(gb, lr=0.0, eps=1e-15)
    for step in range(1, 8):
        _set_grads(ga, gb, gen, step)
        ours.step()
        ref.step()
        for A, B in zip(ga, gb):
            for pa, pb in zip(A["params"], B["params"]):
                _close(pa, pb, max(float(pb.abs().max()), A["lr"]), f"{A['name']} step {step}")
                if pb in ref.state and len(ref.state[pb]):
                    gmax = float(pb.grad.abs().max())
                    m_ref, v_ref = ref.state[pb]["exp_avg"], ref.state[pb]["exp_avg_sq"]
                    _close(ours.state[pa]["exp_avg"], m_ref, max(gmax, float(m_ref.abs().max())), "exp_avg")
                    _close(ours.state[pa]["exp_avg_sq"], v_ref, max(gmax * gmax, float(v_ref.abs().max())), "exp_avg_sq")
                    assert float(ours.state[pa]["step"]) == float(ref.state[pb]["step"]) == step
                else:
                    assert len(ours.state[pa]) == 0          # no gradient -> no state, parameter untouched
    # bit-exactness is the usual outcome (same operation order as torch's foreach kernels); report it
    same = all(torch.equal(pa, pb) for A, B in zip(ga, gb) for pa, pb in zip(A["params"], B["params"]))
    print("bit-identical to torch.optim.Adam after 7 steps:", same)


def test_fused_adam_matches_oracle_and_survives_densification_edits():
    """State stays an ordinary torch.optim.Adam state: the reference's densification replaces
    param / exp_avg / exp_avg_sq by concatenated tensors (gaussian_model.py:1044-1123); stepping on."""
    from mobgs_b200.optim import FusedAdam
    from oracle.adam_ref import adam_step
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(1)
    p = torch.randn(1001, 3, generator=gen).to(dev).requires_grad_(True)
    opt = FusedAdam([{"params": [p], "lr": 1e-2, "name": "xyz"}], lr=0.0, eps=1e-15)
    P, M, V = p.detach().cpu().numpy().copy(), np.zeros((1001, 3), np.float32), np.zeros((1001, 3), np.float32)
    for step in (1, 2, 3):
        g = torch.randn(1001, 3, generator=gen)
        p.grad = g.to(dev)
        opt.step()
        adam_step(P, g.numpy(), M, V, step, 1e-2)
    np.testing.assert_allclose(p.detach().cpu().numpy(), P, rtol=2e-6, atol=1e-6)
    # densify: 99 new rows with zero moments, exactly like cat_tensors_to_optimizer
    group = opt.param_groups[0]
    st = opt.state.pop(group["params"][0])
    new = torch.randn(99, 3, generator=gen).to(dev)
    st["exp_avg"] = torch.cat((st["exp_avg"], torch.zeros_like(new)), 0)
    st["exp_avg_sq"] = torch.cat((st["exp_avg_sq"], torch.zeros_like(new)), 0)
    p2 = torch.nn.Parameter(torch.cat((p.detach(), new), 0).requires_grad_(True))
    group["params"][0] = p2
    opt.state[p2] = st
    P = np.concatenate([P, new.cpu().numpy()]); M = np.concatenate([M, np.zeros((99, 3), np.float32)])
    V = np.concatenate([V, np.zeros((99, 3), np.float32)])
    g = torch.randn(1100, 3, generator=gen)
    p2.grad = g.to(dev)
    opt.step()
    adam_step(P, g.numpy(), M, V, 4, 1e-2)
    np.testing.assert_allclose(p2.detach().cpu().numpy(), P, rtol=2e-6, atol=1e-6)


def test_fused_step_over_two_optimisers_is_one_launch():
    from mobgs_b200 import _lib
    from mobgs_b200.optim import FusedAdam, fused_step
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(2)
    ga, gb = _groups(gen, dev, 301, 401), _groups(gen, dev, 211, 97)
    ra, rb = _clone_groups(ga), _clone_groups(gb)
    oa, ob = FusedAdam(ga, lr=0.0, eps=1e-15), FusedAdam(gb, lr=0.0, eps=1e-15)
    ta, tb = torch.optim.Adam(ra, lr=0.0, eps=1e-15), torch.optim.Adam(rb, lr=0.0, eps=1e-15)
    for step in (1, 2):
        _set_grads(ga, ra, gen, step)
        _set_grads(gb, rb, gen, step)
        before = _lib.LAUNCH_COUNT
        fused_step([oa, ob])
        assert _lib.LAUNCH_COUNT - before == 1
        ta.step(); tb.step()
    for G, R in ((ga, ra), (gb, rb)):
        for A, B in zip(G, R):
            for pa, pb in zip(A["params"], B["params"]):
                _close(pa, pb, max(float(pb.abs().max()), A["lr"]), A["name"])


def test_fused_adam_rejects_cpu_parameters():
    from mobgs_b200.optim import FusedAdam
    p = torch.zeros(4, requires_grad=True)
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        FusedAdam([p], lr=1e-3).step()
