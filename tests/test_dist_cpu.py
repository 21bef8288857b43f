"""world_size-2 gloo test of the multi-GPU host logic (view sharding + flat gradient all-reduce).
The CPU oracle stands in for the kernels: two ranks each render one view of the same replicated
scene; after FlatGradients.reduce() both hold the gradient a single process gets from both views."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_items_partitions_exactly():
    from mobgs_b200.dist import shard_items, shard_subframes
    for n in (0, 1, 7, 14, 18):
        for world in (1, 2, 3, 4, 8):
            got = [i for r in range(world) for i in shard_items(n, r, world)]
            assert got == list(range(n))
            assert max(len(shard_items(n, r, world)) for r in range(world)) - \
                min(len(shard_items(n, r, world)) for r in range(world)) <= 1
    # 2 views x K=7 on 4 ranks: every (view, k) exactly once
    seen = []
    for r in range(4):
        for v, ks in shard_subframes(2, 7, r, 4):
            seen += [(v, k) for k in ks]
    assert sorted(seen) == [(v, k) for v in range(2) for k in range(7)]


def _view_loss(stat, dyn, intr, view_idx):
    from oracle import mobgs_ref as M
    from mobgs_b200.scene import make_camera, subframe_w2c
    cam = make_camera(intr, subframe_w2c(view_idx, 3), time=0.3 + 0.2 * view_idx)
    out = M.render_ref(cam, stat, dyn, None, torch.zeros(3))
    tgt = torch.rand(out["render"].shape, generator=torch.Generator().manual_seed(view_idx))
    return (out["render"] - tgt).abs().mean()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mobgs_b200.dist import FlatGradients, shard_items, trainable
    from mobgs_b200.scene import synthetic_scene
    torch.set_num_threads(2)
    stat, dyn, intr = synthetic_scene(60, 40, 32, 32, seed=9)
    params = trainable(stat.parameters()) + trainable(dyn.parameters())
    for v in shard_items(2, rank, world):
        _view_loss(stat, dyn, intr, v).backward()
    params = [p for p in params if p.grad is not None]
    fg = FlatGradients(params)
    fg.reduce()
    q.put((rank, [p.grad.detach().numpy().copy() for p in params]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_single_process():
    sys.path.insert(0, ROOT)
    from mobgs_b200.dist import trainable
    from mobgs_b200.scene import synthetic_scene
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    stat, dyn, intr = synthetic_scene(60, 40, 32, 32, seed=9)
    params = trainable(stat.parameters()) + trainable(dyn.parameters())
    (_view_loss(stat, dyn, intr, 0) + _view_loss(stat, dyn, intr, 1)).backward()
    want = [p.grad for p in params if p.grad is not None]
    for r in (0, 1):
        assert len(results[r]) == len(want)
        for a, b in zip(results[r], want):
            torch.testing.assert_close(torch.from_numpy(a), b, atol=1e-6, rtol=1e-5)


def _sub_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mobgs_b200.dist import FlatGradients, blur_from_partial_sums, shard_items, trainable
    from mobgs_b200.scene import make_camera, subframe_w2c, synthetic_scene
    from oracle import mobgs_ref as M
    torch.set_num_threads(2)
    K = 3
    stat, dyn, intr = synthetic_scene(60, 40, 32, 32, seed=4)
    deltas = (torch.linspace(-1, 1, K) * 0.4).tolist()
    mine = shard_items(K, rank, world)
    imgs = [M.render_ref(make_camera(intr, subframe_w2c(k, K)), stat, dyn, None, torch.zeros(3),
                         delta_exposure=deltas[k])["render"] for k in mine]
    local = torch.stack(imgs) if imgs else torch.zeros(0, 3, 32, 32)
    pred = blur_from_partial_sums(local, K)
    tgt = torch.rand(pred.shape, generator=torch.Generator().manual_seed(0))
    (pred - tgt).abs().mean().backward()
    params = [p for p in trainable(stat.parameters()) + trainable(dyn.parameters()) if p is not stat.control_xyz]
    for p in params:       # a rank with no dynamic contribution still needs a zero gradient to reduce
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    FlatGradients(params).reduce()
    q.put((rank, pred.detach().numpy().copy(), [p.grad.numpy().copy() for p in params]))
    dist.barrier()
    dist.destroy_process_group()


def test_subframe_sharding_matches_single_process():
    """K sub-frames of ONE view split across 2 ranks: partial-sum all-reduce + gradient all-reduce
    reproduce the single-process blurred image and gradients."""
    sys.path.insert(0, ROOT)
    from mobgs_b200.dist import trainable
    from mobgs_b200.scene import make_camera, subframe_w2c, synthetic_scene
    from oracle import mobgs_ref as M
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sub_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: (pred, grads) for r, pred, grads in (q.get(timeout=240) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    K = 3
    stat, dyn, intr = synthetic_scene(60, 40, 32, 32, seed=4)
    deltas = (torch.linspace(-1, 1, K) * 0.4).tolist()
    imgs = [M.render_ref(make_camera(intr, subframe_w2c(k, K)), stat, dyn, None, torch.zeros(3),
                         delta_exposure=deltas[k])["render"] for k in range(K)]
    pred = M.blur_mean(imgs)
    tgt = torch.rand(pred.shape, generator=torch.Generator().manual_seed(0))
    (pred - tgt).abs().mean().backward()
    params = [p for p in trainable(stat.parameters()) + trainable(dyn.parameters()) if p is not stat.control_xyz]
    for r in (0, 1):
        torch.testing.assert_close(torch.from_numpy(res[r][0]), pred.detach(), atol=1e-6, rtol=1e-5)
        for a, p in zip(res[r][1], params):
            want = p.grad if p.grad is not None else torch.zeros_like(p)
            torch.testing.assert_close(torch.from_numpy(a), want, atol=1e-6, rtol=1e-5)


def _inplace_worker(rank, world, port, q):
    """Gradients that are views into one flat buffer (what fused.synth_project's backward returns) are
    all-reduced in place; a gradient with its own storage goes through the packed path in the same call."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mobgs_b200.dist import FlatGradients
    g = torch.Generator().manual_seed(100 + rank)
    shapes = [(7, 3), (5, 4), (11,)]
    params = [torch.zeros(s, requires_grad=True) for s in shapes] + [torch.zeros(6, 12, requires_grad=True)]
    flat = torch.zeros(sum((torch.Size(s).numel() + 3) // 4 * 4 for s in shapes))
    off = 0
    for p, s in zip(params, shapes):
        n = torch.Size(s).numel()
        flat[off:off + n] = torch.randn(n, generator=g)
        p.grad = flat[off:off + n].view(s)
        off += (n + 3) // 4 * 4
    params[3].grad = torch.randn(6, 12, generator=g)
    ptrs = [p.grad.data_ptr() for p in params[:3]]
    fg = FlatGradients(params, inplace_shared=True)
    fg.reduce()
    assert [p.grad.data_ptr() for p in params[:3]] == ptrs          # reduced where they were
    assert fg.last_collective_elems == flat.numel() + 72
    q.put((rank, [p.grad.detach().numpy().copy() for p in params]))
    dist.barrier()
    dist.destroy_process_group()


def test_inplace_reduction_of_shared_gradient_storage():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_inplace_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = []
    for i, s in enumerate([(7, 3), (5, 4), (11,), (6, 12)]):
        tot = torch.zeros(s)
        for rank in (0, 1):
            g = torch.Generator().manual_seed(100 + rank)
            vals = [torch.randn(torch.Size(t).numel(), generator=g).view(t) for t in [(7, 3), (5, 4), (11,)]]
            vals.append(torch.randn(6, 12, generator=g))
            tot += vals[i]
        want.append(tot)
    for r in (0, 1):
        for a, b in zip(results[r], want):
            torch.testing.assert_close(torch.from_numpy(a), b, atol=1e-6, rtol=1e-6)
