"""Host path: the CUDA-graphed step (mobgs_b200.graphs.GraphedStep) replays exactly what the eager K-batched step
computes — outputs, parameter gradients, pose gradients, densification gradient — for new inputs, and a replay whose
tile lists overflow the capacity baked into the graph is detected, recomputed exactly and re-captured."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rays(intr, views):
    from mobgs_b200.scene import make_camera
    return torch.cat([make_camera(intr, v).cam_ray for v in views])


def _views(K, shift_x=0.0, yaw=0.0, device="cuda"):
    from mobgs_b200.scene import subframe_w2c
    vs = []
    for k in range(K):
        w = subframe_w2c(k, K)
        w[0, 3] += shift_x
        w[1, 3] += 0.02 * yaw
        vs.append(w)
    return torch.stack(vs).to(device)


def _make(ns, nd, W, H, K, seed):
    from mobgs_b200.losses import l1_loss
    from mobgs_b200.scene import synthetic_scene
    from mobgs_b200.subframes import render_subframes
    sc, dc, intr = synthetic_scene(ns, nd, W, H, seed=seed, device="cuda")
    Kmat = torch.tensor([[intr.fx, 0, intr.cx], [0, intr.fy, intr.cy], [0, 0, 1]], device="cuda")
    bg = torch.tensor([0.1, 0.2, 0.3], device="cuda")
    params = [p for pc in (sc, dc) for p in pc.parameters() if p.requires_grad]

    def fn(view, tpoly, rays, tgt):
        out = render_subframes(sc, dc, view, Kmat, tpoly.clamp(0, 1), tpoly, rays, bg, W, H)
        loss = l1_loss(out["render"], tgt) + 0.1 * out["depth"].mean()
        loss.backward()
        return loss.detach(), out["render"].detach(), out["viewspace_points"].grad

    return sc, dc, intr, params, fn


def _eager(fn, params, inputs):
    for p in params:
        p.grad = None
    ins = [t.detach().clone().requires_grad_(t.requires_grad) for t in inputs]
    outs = fn(*ins)
    torch.cuda.synchronize()
    return [o.clone() for o in outs], [None if p.grad is None else p.grad.clone() for p in params], ins[0].grad.clone()


def _close(a, b, name):
    scale = float(b.abs().max()) + 1e-12
    err = float((a - b).abs().max())
    assert err <= 2e-4 * scale + 1e-7, (name, err, scale)      # atomics reorder sums between runs: not bit-equal


def _compare(step, params, ref, name):
    outs, grads, vgrad = ref
    got = step.outputs
    for i, (a, b) in enumerate(zip(got, outs)):
        _close(a, b, f"{name} output {i}")
    for i, (p, g) in enumerate(zip(params, grads)):
        if g is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
        else:
            _close(p.grad, g, f"{name} grad {i}")
    _close(step.input_grads[0], vgrad, name + " pose grad")


def test_graphed_step_replays_the_eager_step():
    from mobgs_b200.graphs import GraphedStep
    K, W, H = 3, 112, 80
    sc, dc, intr, params, fn = _make(700, 500, W, H, K, seed=21)
    g = torch.Generator().manual_seed(4)

    def inputs(yaw):
        v = _views(K, yaw=yaw)
        tp = (0.5 + torch.linspace(-1, 1, K) * 0.4 / 23 + 0.01 * yaw).cuda()
        return [v.requires_grad_(True), tp, _rays(intr, v.detach().cpu()).cuda(), torch.rand(3, H, W, generator=g).cuda()]

    first = inputs(0.0)
    step = GraphedStep(fn, first, params)
    assert step.captures == 1
    for yaw in (0.0, 1.0, -2.0):
        ins = inputs(yaw)
        ref = _eager(fn, params, ins)
        step(*ins)
        assert step.validate() and step.overflows == 0
        _compare(step, params, ref, f"yaw {yaw}")
        n, cap = step.intersection_counts()[0]
        assert 0 < n <= cap
    assert step.replays == 3 and step.captures == 1
    # the caller may drop the gradients between steps (optimizer.zero_grad(set_to_none=True)): they come back
    for p in params:
        p.grad = None
    step(*first)
    assert step.validate()
    assert [p.grad is not None for p in params] == [g is not None for g in step.static_grads]
    assert sum(p.grad is not None for p in params) >= 10


def test_list_overflow_is_detected_recomputed_and_recaptured():
    from mobgs_b200.graphs import GraphedStep
    K, W, H = 3, 160, 112
    sc, dc, intr, params, fn = _make(4000, 2500, W, H, K, seed=22)
    tgt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(6)).cuda()
    tp = torch.full((K,), 0.5).cuda()

    def inputs(shift):
        v = _views(K, shift_x=shift)
        return [v.requires_grad_(True), tp, _rays(intr, v.detach().cpu()).cuda(), tgt]

    # captured while most of the scene is outside the frustum: the capacity baked into the graph is small
    step = GraphedStep(fn, inputs(9.0), params)
    cap = step.intersection_counts()[0][1]
    ins = inputs(0.0)
    ref = _eager(fn, params, ins)
    step(*ins)
    ok = step.validate()
    assert not ok and step.overflows == 1 and step.captures == 2, (ok, step.overflows, step.captures, cap)
    assert step.intersection_counts()[0][1] > cap
    _compare(step, params, ref, "overflowed replay (recomputed)")
    # the re-captured graph holds the larger lists: the same inputs now replay cleanly
    step(*ins)
    assert step.validate() and step.overflows == 1
    _compare(step, params, ref, "replay after re-capture")
