"""Multi-GPU plumbing: one process per GPU, parameters replicated, (view, sub-frame) work items
sharded across ranks, one all-reduce of the flat gradient buffer per step.

The reference is single-process / single-GPU (no torch.distributed anywhere); the loss of one
optimiser step is a mean over the batch's views (train.py:621-629), i.e. linear across views, so
rendering different views on different ranks and summing gradients reproduces the single-process
gradient exactly.  Gaussians themselves are never partitioned (per-tile sorting needs all
Gaussians of a tile): "replicas only" along N.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_items(n_items: int, rank: int, world: int) -> range:
    """Contiguous block partition of work items (views, or (view, sub-frame) pairs flattened as
    v*K+k so that the sub-frames of one view stay on neighbouring ranks)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_subframes(n_views: int, K: int, rank: int, world: int) -> List[Tuple[int, range]]:
    """[(view, sub-frame range)] owned by `rank` when V*K work items are split across ranks."""
    out = []
    items = shard_items(n_views * K, rank, world)
    if len(items) == 0:
        return out
    v0, v1 = items.start // K, (items.stop - 1) // K
    for v in range(v0, v1 + 1):
        lo = max(items.start, v * K) - v * K
        hi = min(items.stop, (v + 1) * K) - v * K
        out.append((v, range(lo, hi)))
    return out


class FlatGradients:
    """Flat fp32 buffer over the gradients of a fixed parameter list; `.reduce()` = one collective."""

    def __init__(self, params: Sequence[torch.Tensor], inplace_shared: bool = False):
        """inplace_shared: reduce gradients that share one storage in place (see shared_storages).  Only
        valid when every rank builds its gradients through the same autograd graph (view sharding), so
        that all ranks find the same shared buffers; with sub-frame sharding idle ranks have no gradients
        and every rank must take the packed path."""
        self.inplace_shared = inplace_shared
        self.params = list(params)
        self.sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(self.sizes), dtype=torch.float32, device=self.params[0].device)

    def pack(self) -> torch.Tensor:
        off = 0
        for p, n in zip(self.params, self.sizes):
            if p.grad is not None:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            else:
                self.flat[off:off + n].zero_()
            off += n
        return self.flat

    def unpack(self) -> None:
        off = 0
        for p, n in zip(self.params, self.sizes):
            g = self.flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n

    def shared_storages(self):
        """Gradients that already live in one flat buffer (fused.synth_project's backward returns views
        into a single allocation, which autograd adopts as `.grad`): [(alias over the whole storage,
        parameter indices)] for every storage shared by more than one gradient — such a buffer is reduced
        in place, without pack / unpack.  Deterministic across ranks (order of first appearance)."""
        groups = {}
        for i, p in enumerate(self.params):
            g = p.grad
            if g is None or g.dtype != torch.float32 or not g.is_contiguous():
                continue
            groups.setdefault(g.untyped_storage().data_ptr(), []).append(i)
        out = []
        for idx in groups.values():
            if len(idx) < 2:
                continue
            g0 = self.params[idx[0]].grad
            st = g0.untyped_storage()
            alias = torch.empty(0, dtype=torch.float32, device=g0.device).set_(st, 0, (st.nbytes() // 4,))
            out.append((alias, idx))
        return out

    def reduce(self, group=None, average: bool = False) -> torch.Tensor:
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        shared = self.shared_storages() if self.inplace_shared else []
        covered = {i for _, idx in shared for i in idx}
        self.last_collective_elems = 0
        from . import fused
        sink = fused.GRAD_SINK
        for alias, _ in shared:                      # in place: these gradients need no copy at all
            if sink is not None and sink.reduced_storage == alias.untyped_storage().data_ptr():
                sink.reduced_storage = None          # already summed over the ranks inside the backward (GradSink)
                self.last_collective_elems += sink.last_collective_elems
                continue
            if multi:
                if not (_SYMMETRIC is not None and _SYMMETRIC.all_reduce(alias)):      # NVLS multimem path, else NCCL
                    dist.all_reduce(alias, group=group)
                if average:
                    alias.div_(dist.get_world_size(group))
            self.last_collective_elems += alias.numel()
        if len(covered) == len(self.params):
            return self.flat
        # everything else goes through the packed buffer (segments of covered parameters stay zero)
        off = 0
        for i, (p, n) in enumerate(zip(self.params, self.sizes)):
            if i not in covered:
                if p.grad is not None:
                    self.flat[off:off + n].copy_(p.grad.reshape(-1))
                else:
                    self.flat[off:off + n].zero_()
            off += n
        rest = [i for i in range(len(self.params)) if i not in covered]
        lo = sum(self.sizes[:rest[0]])
        hi = sum(self.sizes[:rest[-1] + 1])
        if multi:
            dist.all_reduce(self.flat[lo:hi], group=group)
            if average:
                self.flat[lo:hi].div_(dist.get_world_size(group))
        self.last_collective_elems += hi - lo
        off = 0
        for i, (p, n) in enumerate(zip(self.params, self.sizes)):
            if i not in covered:
                g = self.flat[off:off + n].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
            off += n
        return self.flat


class SymmetricGradients:
    """The flat Gaussian-gradient buffer in NVLS symmetric memory (torch.distributed._symmetric_memory), reduced with
    the multimem (in-switch) all-reduce instead of NCCL: measured on 8 x B200 for the 116 MB buffer of the 1 M scene,
    0.285 ms vs 0.36 ms (profiles/r2_allreduce_n8.json).  `install()` makes fused._SynthProject.backward write its
    parameter gradients straight into the symmetric buffer (one persistent allocation per size, re-zeroed per step —
    the `.grad` views of successive steps therefore alias each other, which is what an optimiser loop wants and what
    a caller that keeps old gradients must know); `FlatGradients.reduce()` recognises the buffer and reduces it in
    place with `torch.ops.symm_mem.multimem_all_reduce_`.  Falls back to NCCL (returns False) when symmetric memory or
    multicast is unavailable."""

    def __init__(self, group=None):
        self.group = group if group is not None else dist.group.WORLD
        self.buf = None
        self.storage_ptr = None
        self.free = False
        self.slice_ok = None

    def begin_step(self):
        """Call once per optimiser step, before the backward: the NEXT flat gradient buffer the backward asks for is
        the symmetric one.  Further requests within the same step (a second backward pass over a retained graph,
        whose gradients autograd accumulates into the first pass's) get ordinary memory and the NCCL path — the
        persistent buffer is never re-zeroed under live gradients."""
        self.free = True

    def available(self) -> bool:
        try:
            import torch.distributed._symmetric_memory as symm_mem  # noqa: F401
            return hasattr(torch.ops.symm_mem, "multimem_all_reduce_")
        except Exception:  # noqa: BLE001
            return False

    def alloc(self, n: int, device) -> torch.Tensor:
        import torch.distributed._symmetric_memory as symm_mem
        if not self.free:
            return torch.zeros(n, device=device)
        self.free = False
        if self.buf is None or self.buf.numel() != n:          # collective: every rank reaches it with the same n
            self.buf = symm_mem.empty(n, dtype=torch.float32, device=device)
            hdl = symm_mem.rendezvous(self.buf, self.group.group_name)
            if not getattr(hdl, "multicast_ptr", 0):
                raise RuntimeError("symmetric memory without multicast (NVLS) support")
            self.storage_ptr = self.buf.untyped_storage().data_ptr()
        self.buf.zero_()
        return self.buf

    def install(self) -> bool:
        from . import fused
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1 and self.available()):
            return False
        import torch.distributed._symmetric_memory as symm_mem
        probe = symm_mem.empty(1024, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        if not getattr(symm_mem.rendezvous(probe, self.group.group_name), "multicast_ptr", 0):
            return False                                       # no NVLS multicast on this system: stay on NCCL
        fused.FLAT_ALLOCATOR = self.alloc
        global _SYMMETRIC
        _SYMMETRIC = self
        return True

    def uninstall(self):
        from . import fused
        global _SYMMETRIC
        fused.FLAT_ALLOCATOR = None
        _SYMMETRIC = None

    def all_reduce_slice(self, part: torch.Tensor) -> bool:
        """in-place sum over ranks of a contiguous, 16-byte aligned slice of the symmetric buffer (same order on every
        rank); False if `part` is not such a slice or the op refuses views"""
        if (self.buf is None or self.slice_ok is False or part.untyped_storage().data_ptr() != self.storage_ptr
                or not part.is_contiguous() or (part.storage_offset() * 4) % 16 or (part.numel() * 4) % 16):
            return False
        try:
            torch.ops.symm_mem.multimem_all_reduce_(part, "sum", self.group.group_name)
            self.slice_ok = True
            return True
        except Exception:  # noqa: BLE001   (every rank fails alike: same op, same arguments)
            self.slice_ok = False
            return False

    def all_reduce(self, alias: torch.Tensor) -> bool:
        """in-place sum over ranks if `alias` is (a view over) the symmetric buffer"""
        if self.buf is None or alias.untyped_storage().data_ptr() != self.storage_ptr:
            return False
        torch.ops.symm_mem.multimem_all_reduce_(self.buf, "sum", self.group.group_name)
        return True


_SYMMETRIC = None


def overlap_gradient_allreduce(enable: bool = True, n_chunks: int = 2, group=None):
    """Install (or remove) the overlapped Gaussian-gradient all-reduce: the projection backward then runs in
    `n_chunks` Gaussian ranges and all-reduces each range's gradient block on a side stream while the next range's
    kernel runs (mobgs_b200.fused.GradSink).  The gradients autograd delivers are then ALREADY summed over the
    ranks; `FlatGradients.reduce()` recognises that buffer and only reduces what is left (decoder weights, poses).
    No-op semantics on a single rank (the collective is skipped, the chunked launches still run)."""
    from . import fused
    if not enable:
        fused.GRAD_SINK = None
        return None
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def reduce(t):
        if not multi:
            return None
        # a slice of the NVLS symmetric gradient buffer: in-switch (multimem) reduction of just that slice, on the
        # side stream the caller has made current; anything else: NCCL
        if _SYMMETRIC is not None and _SYMMETRIC.all_reduce_slice(t):
            return None
        return dist.all_reduce(t, group=group, async_op=True)

    # high priority: the collective's CTAs must get SM slots while the next range's projection kernel still has CTAs queued
    fused.GRAD_SINK = fused.GradSink(reduce, torch.cuda.Stream(priority=-1), n_chunks)
    return fused.GRAD_SINK


def trainable(params: Iterable[torch.Tensor]) -> List[torch.Tensor]:
    return [p for p in params if p.is_floating_point() and p.requires_grad]


class _AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x.  Every rank goes on to compute the SAME scalar loss from y, and the
    per-rank parameter gradients are summed afterwards (FlatGradients.reduce), so the backward
    hands each rank d loss / d y for its own addend unchanged."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(y, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


def all_reduce_sum(x: torch.Tensor, group=None) -> torch.Tensor:
    return _AllReduceSum.apply(x, group)


def blur_from_partial_sums(local_subframes: torch.Tensor, K: int, group=None) -> torch.Tensor:
    """Sub-frame sharding (collective 1): each rank holds the decoded images of its
    share of the K latent sub-frames, `local_subframes` [k_local,3,H,W] (k_local may be 0).  One
    all-reduce of the [3,H,W] partial sum gives every rank the blurred prediction
    mean_k(rgb_k) + 1e-10 (train.py:540-541)."""
    partial = local_subframes.sum(dim=0)
    return all_reduce_sum(partial, group) / K + 1e-10
