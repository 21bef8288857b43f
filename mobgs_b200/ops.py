"""torch.autograd.Function wrappers over the C ABI (include/mobgs_b200.h).

PyTorch is plumbing here: it owns device memory, streams and the autograd graph; every
arithmetic step of the path runs in libmobgs_b200.so.  All inputs must be CUDA fp32 tensors —
there is no CPU path and none is attempted.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch

from . import _lib as L

EPS2D, NEAR, FAR, RADIUS_CLIP = 0.3, 0.01, 1e10, 0.0   # gsplat.rendering.rasterization defaults


def _stream() -> int:
    return L.current_stream()


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("mobgs_b200 ops need CUDA tensors (no CPU fallback exists)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _cams(viewmats, Ks, width, height, eps2d=EPS2D, near=NEAR, far=FAR, radius_clip=RADIUS_CLIP):
    return L.Cameras(viewmats.shape[0], int(width), int(height), _p(viewmats), _p(Ks),
                     float(eps2d), float(near), float(far), float(radius_clip))


def _strided(g: Optional[torch.Tensor], N: int, inner: int):
    """Returns (tensor kept alive, pointer, per-Gaussian stride in floats) for a gradient of shape
    [K,N,inner] / [K,N] whose layout is (k*N+g)*stride + j — e.g. a view into packed records."""
    if g is None:
        return None, None, 0
    st = g.stride()
    if inner > 1:
        ok = g.dim() == 3 and st[2] == 1 and st[0] == N * st[1] and st[1] >= inner
    else:
        ok = g.dim() == 2 and st[0] == N * st[1] and st[1] >= 1
    if g.dtype != torch.float32 or not ok:
        g = g.float().contiguous()
        return g, g.data_ptr(), inner
    return g, g.data_ptr(), st[1]


class _Project(torch.autograd.Function):
    """gsplat.rendering.fully_fused_projection (pinhole, packed=False, covars=None)."""

    @staticmethod
    def forward(ctx, means, quats, scales, viewmats, Ks, width, height, eps2d, near, far, radius_clip):
        means, quats, scales = _f32c(means), _f32c(quats), _f32c(scales)
        viewmats, Ks = _f32c(viewmats), _f32c(Ks)
        K, N = viewmats.shape[0], means.shape[0]
        dev = means.device
        radii = torch.empty(K, N, dtype=torch.int32, device=dev)
        means2d = torch.empty(K, N, 2, device=dev)
        depths = torch.empty(K, N, device=dev)
        conics = torch.empty(K, N, 3, device=dev)
        cams = _cams(viewmats, Ks, width, height, eps2d, near, far, radius_clip)
        a = L.ProjectFwd(cams, N, _p(means), _p(quats), _p(scales), _p(radii), _p(means2d), _p(depths), _p(conics))
        L.call("mobgs_project_fwd", a, _stream())
        ctx.save_for_backward(means, quats, scales, viewmats, Ks, radii)
        ctx.cfg = (width, height, eps2d, near, far, radius_clip)
        ctx.mark_non_differentiable(radii)
        return radii, means2d, depths, conics

    @staticmethod
    def backward(ctx, _g_radii, g_means2d, g_depths, g_conics):
        means, quats, scales, viewmats, Ks, radii = ctx.saved_tensors
        width, height, eps2d, near, far, radius_clip = ctx.cfg
        K, N = viewmats.shape[0], means.shape[0]
        dev = means.device
        v_means = torch.empty(N, 3, device=dev)
        v_quats = torch.empty(N, 4, device=dev)
        v_scales = torch.empty(N, 3, device=dev)
        v_viewmats = torch.zeros(K, 4, 4, device=dev) if ctx.needs_input_grad[3] else None
        k1, p1, s1 = _strided(g_means2d, N, 2)
        k2, p2, s2 = _strided(g_depths, N, 1)
        k3, p3, s3 = _strided(g_conics, N, 3)
        cams = _cams(viewmats, Ks, width, height, eps2d, near, far, radius_clip)
        a = L.ProjectBwd(cams, N, _p(means), _p(quats), _p(scales), _p(radii), p1, s1, p2, s2, p3, s3,
                         _p(v_means), _p(v_quats), _p(v_scales), _p(v_viewmats))
        L.call("mobgs_project_bwd", a, _stream())
        del k1, k2, k3
        return v_means, v_quats, v_scales, v_viewmats, None, None, None, None, None, None, None


def project(means, quats, scales, viewmats, Ks, width, height, eps2d=EPS2D, near_plane=NEAR,
            far_plane=FAR, radius_clip=RADIUS_CLIP):
    return _Project.apply(means, quats, scales, viewmats, Ks, int(width), int(height),
                          float(eps2d), float(near_plane), float(far_plane), float(radius_clip))


class TileLists:
    """Per-(sub-frame, tile) depth-sorted Gaussian lists."""

    def __init__(self, tile_offsets, sorted_ids, n_isect, K, width, height):
        self.tile_offsets = tile_offsets
        self.sorted_ids = sorted_ids
        self.n_isect = n_isect
        self.capacity = n_isect
        self.K, self.width, self.height = K, width, height


# Speculative list sizing: the number of tile/Gaussian intersections I is only known on the device
# after mobgs_tile_count.  Reading it back before sizing the lists stalls the launch pipeline in
# the middle of the forward.  Instead the lists are sized from the last I seen for the same launch
# shape (+25 %), emit/sort/blend are enqueued immediately (the kernels never write or read beyond
# the capacity they are given), I travels to pinned host memory asynchronously, and only after
# everything is enqueued does the host check it — redoing emit/sort/blend with the exact size in the
# (rare) overflow case, so results are always those of the exactly-sized lists.
_CAP_CACHE = {}
SPECULATIVE_LISTS = os.environ.get("MOBGS_SPECULATIVE_LISTS", "1") != "0"
# 0: always the two-pass count / emit kernels (ablation; the default records entries in the counting pass)
RECORD_ENTRIES = os.environ.get("MOBGS_RECORD_ENTRIES", "1") != "0"


# Set by mobgs_b200.graphs.GraphedStep while it captures a CUDA graph (a graphs._Capture): the tile lists are then sized
# from the capacity guess only, the intersection count travels to a pre-allocated pinned word by a captured copy, and the
# "did the lists fit" check — the one host synchronisation of the eager path — moves behind the replay
# (GraphedStep.validate), which redoes an overflowed step eagerly and re-captures.
CAPTURE = None


class BinPlan:
    """Counting pass fused into the projection (MobgsSynthFwd.bin_tile_counts): created by the caller that knows which
    lists the blend will walk BEFORE it projects, handed to `fused.synth_project(..., binning=plan)` — which lets the
    projection kernel count and record the tile intersections while each projected Gaussian is still in registers — and
    then to `build_tile_lists(..., plan=plan)`, which only runs the prefix sum, the scatter and the sorts.  Speculative
    like the lists themselves: it needs the entry capacity guessed from the previous launch of the same shape, so the
    first call of a shape (and an overflow redo) take the stand-alone counting kernels."""

    def __init__(self, specs, width, height, tight=True):
        self.specs = tuple(tuple(int(v) for v in s) for s in specs)
        self.width, self.height, self.tight = int(width), int(height), bool(tight)
        self.counts = self.entries = self.cursor = None
        self.N = None

    def key(self, N, dev):
        return (len(self.specs), int(N), self.width, self.height, self.specs, self.tight, dev.index)

    def prepare(self, N, dev):
        """-> (lists struct, counts, entries, cursor) for the projection launch, or None (no capacity guess yet)"""
        if not (SPECULATIVE_LISTS and RECORD_ENTRIES and FUSED_COUNT) or len(self.specs) > L.MAX_K:
            return None
        guess = _CAP_CACHE.get(self.key(N, dev))
        if guess is None:
            return None
        # The projection kernel owns one Gaussian per thread for all K sub-frames, so its slot atomics are issued one
        # after the other: fine at ~1.4 tiles per (list, Gaussian) (150 k / 512x288 / K = 9: count 0.19 -> +0.11 ms on the
        # projection, step 1.41 -> 1.31 ms; 1 M / 1080p: neutral), hopeless at 35 (heavy-footprint run: 6.8 -> 22 ms).
        if guess > FUSED_COUNT_MAX_TILES * len(self.specs) * max(int(N), 1) + 4096:
            return None
        if len(self.specs) * int(N) > FUSED_COUNT_MAX_PAIRS:
            return None
        tiles = math.ceil(self.width / L.TILE) * math.ceil(self.height / L.TILE)
        self.counts = torch.empty(len(self.specs) * tiles, dtype=torch.int32, device=dev)
        self.entries = torch.empty(max(int(guess), 1), 4, dtype=torch.int32, device=dev)
        self.cursor = torch.empty(1, dtype=torch.int32, device=dev)
        self.N = int(N)
        return L.make_lists(self.specs), self.counts, self.entries, self.cursor


# 0: never fuse the counting pass into the projection (ablation)
FUSED_COUNT = os.environ.get("MOBGS_FUSED_COUNT", "1") != "0"
FUSED_COUNT_MAX_TILES = 2.5      # average tile entries per (list, Gaussian) above which the stand-alone pass is used
FUSED_COUNT_MAX_PAIRS = 3_000_000   # (list, Gaussian) pairs above which it is used too: at 7 M pairs (1 M / 1080p / K = 7) the
                                    # fused form measured 0.70 vs 0.69 ms for projection + binning, at 1.35 M pairs 0.22 vs 0.29 ms


def build_tile_lists(records, radii, depths, width, height, tight=True, specs=None, consume=None, tile_list=None, plan=None):
    """mobgs_tile_count -> mobgs_tile_emit_sort (-> consume(lists)).

    specs: [(record_set, g_begin, g_end)] per list; default = one full-range list per record set.
    tile_list: optional per-list index of the tile lists the list walks — lists with the same geometry and
    index range (but different colour payloads) name the same index and are binned / sorted ONCE (the index
    of a group is the position of its first member among the distinct values).  Default: every list its own.
    consume: optional callable(lists) that enqueues the kernels using the lists (blend); when given,
    returns (lists, consume(lists)) and sizes the lists speculatively (see above); otherwise reads I
    back synchronously and returns lists.
    plan: a BinPlan that the projection launch has already filled (counts + recorded entries): only the prefix sum,
    the scatter and the sorts run here."""
    Kr, N = radii.shape
    if specs is None:
        specs = tuple((k, 0, N) for k in range(Kr))
    if tile_list is None:
        bin_specs, tl_idx = tuple(specs), None
    else:
        firsts = {}
        for i, t in enumerate(tile_list):
            firsts.setdefault(t, i)
        order = sorted(firsts, key=lambda t: firsts[t])
        bin_specs = tuple(specs[firsts[t]] for t in order)
        tl_idx = tuple(order.index(t) for t in tile_list)
        for i, t in enumerate(tl_idx):      # shared lists must agree on the index range (the geometry is the caller's promise)
            if tuple(specs[i][1:]) != tuple(bin_specs[t][1:]):
                raise ValueError("lists that share a tile binning must cover the same Gaussian index range")
    K = len(bin_specs)
    lists = L.make_lists(bin_specs)
    blend_lists = L.make_lists(specs, tl_idx)
    dev = records.device
    tiles = math.ceil(width / L.TILE) * math.ceil(height / L.TILE)
    nt = K * tiles
    counts = torch.empty(nt, dtype=torch.int32, device=dev)
    offsets = torch.empty(nt + 1, dtype=torch.int32, device=dev)
    key = (K, N, width, height, tuple(bin_specs), bool(tight), dev.index)
    guess = _CAP_CACHE.get(key)
    speculative = consume is not None and guess is not None and SPECULATIVE_LISTS
    if CAPTURE is not None and not speculative:
        raise RuntimeError("CUDA-graph capture needs the speculative list path: run the step eagerly once for this launch "
                           "shape (capacity guess) and give build_tile_lists a `consume` callable")
    # Speculative calls also let the counting pass RECORD the intersections it finds (MobgsTileCount.entries), which turns
    # the emit pass into a streaming scatter; the exact-size (first / overflow) calls use the two-pass kernels.
    entries = cursor = None
    precounted = (plan is not None and plan.counts is not None and speculative and tile_list is None and plan.N == N
                  and plan.specs == tuple(bin_specs) and plan.tight == bool(tight)
                  and (plan.width, plan.height) == (width, height))
    if precounted:
        counts, entries, cursor = plan.counts, plan.entries, plan.cursor
        plan.counts = plan.entries = plan.cursor = None          # consumed
    elif speculative and RECORD_ENTRIES:
        entries = torch.empty(max(int(guess), 1), 4, dtype=torch.int32, device=dev)
        cursor = torch.empty(1, dtype=torch.int32, device=dev)
    a = L.TileCount(K, N, width, height, _p(records), _p(radii), int(tight), lists, _p(counts), _p(offsets),
                    _p(depths), _p(entries), 0 if entries is None else entries.shape[0], _p(cursor), int(precounted))
    L.call("mobgs_tile_count", a, _stream())

    def emit_sort(cap, use_entries=False):
        cap = max(int(cap), 1)
        keys = torch.empty(cap, dtype=torch.int64, device=dev)
        keys_tmp = torch.empty(cap, dtype=torch.int64, device=dev)
        sorted_ids = torch.empty(cap, dtype=torch.int32, device=dev)
        b = L.TileSort(K, N, width, height, _p(records), _p(radii), _p(depths), int(tight), lists, _p(offsets),
                       _p(counts), cap, _p(keys), _p(keys_tmp), _p(sorted_ids),
                       *((_p(entries), entries.shape[0], _p(cursor)) if use_entries else (None, 0, None)))
        L.call("mobgs_tile_emit_sort", b, _stream())
        tl = TileLists(offsets, sorted_ids, None, K, width, height)
        tl.lists, tl.capacity, tl.specs = blend_lists, cap, tuple(specs)
        return tl

    if not speculative:
        n_isect = int(offsets[-1].item())
        tl = emit_sort(n_isect)
        tl.n_isect = n_isect
        _CAP_CACHE[key] = int(n_isect * 1.25) + 4096
        return tl if consume is None else (tl, consume(tl))
    if CAPTURE is not None:
        # being captured: no event, no host wait — the count lands in the capture's pinned word on every replay and
        # GraphedStep.validate() compares it with the capacity baked into the graph
        host_n = CAPTURE.pinned_word()
        host_n.copy_(offsets[-1:], non_blocking=True)
        tl = emit_sort(guess, use_entries=entries is not None)
        out = consume(tl)
        CAPTURE.checks.append((host_n, tl.capacity, key))
        tl.n_isect = None
        return tl, out
    host_n = torch.empty(1, dtype=torch.int32, pin_memory=True)
    host_n.copy_(offsets[-1:], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    tl = emit_sort(guess, use_entries=entries is not None)
    out = consume(tl)
    ev.synchronize()
    n_isect = int(host_n[0])
    if n_isect > tl.capacity:           # overflow: redo with the exact size (two-pass emit: the entry list overflowed too)
        tl = emit_sort(n_isect)
        out = consume(tl)
    tl.n_isect = n_isect
    _CAP_CACHE[key] = max(int(n_isect * 1.25) + 4096, int(guess * 0.9))
    return tl, out


class _Rasterize(torch.autograd.Function):
    """isect_tiles + sort + rasterize_to_pixels of gsplat.rendering.rasterization, operating on
    gsplat's SoA tensors.  `depths` (if given and append_depth) becomes colour channel D — the
    torch.cat of render_mode='RGB+ED'."""

    @staticmethod
    def forward(ctx, means2d, conics, colors, opacities, backgrounds, depths, radii, width, height,
                append_depth, tight):
        means2d, conics, colors, opacities = _f32c(means2d), _f32c(conics), _f32c(colors), _f32c(opacities)
        depths = _f32c(depths)
        K, N = radii.shape
        dev = means2d.device
        per_cam = colors.dim() == 3
        D0 = colors.shape[-1]
        D = D0 + (1 if append_depth else 0)
        if D > L.MAX_COLORS:
            raise RuntimeError(f"mobgs_b200 blend kernels carry at most {L.MAX_COLORS} channels, got {D}")
        records = torch.empty(K, N, L.REC, device=dev)
        pk = L.Pack(K, N, D0, _p(means2d), _p(conics), _p(opacities), _p(colors), int(per_cam),
                    _p(depths) if append_depth else None, _p(records))
        L.call("mobgs_pack_records", pk, _stream())
        bg = None
        if backgrounds is not None:
            bg = _f32c(backgrounds)
            if append_depth:
                bg = torch.cat([bg, torch.zeros_like(bg[:, :1])], dim=-1).contiguous()

        def blend(lists):
            out_c = torch.empty(K, height, width, D, device=dev)
            out_a = torch.empty(K, height, width, device=dev)
            last = torch.empty(K, height, width, dtype=torch.int32, device=dev)
            a = L.BlendFwd(K, N, D, width, height, lists.lists, lists.capacity, _p(records), _p(lists.tile_offsets),
                           _p(lists.sorted_ids), _p(bg), _p(out_c), _p(out_a), _p(last))
            L.call("mobgs_blend_fwd", a, _stream())
            return out_c, out_a, last

        lists, (out_c, out_a, last) = build_tile_lists(records, radii, depths, width, height, tight, consume=blend)
        ctx.capacity = lists.capacity
        ctx.save_for_backward(records, lists.tile_offsets, lists.sorted_ids, bg, out_a, last)
        ctx.lists = lists.lists
        ctx.meta = (K, N, D0, D, width, height, per_cam, append_depth)
        ctx.n_isect = lists.n_isect
        return out_c, out_a.unsqueeze(-1)

    @staticmethod
    def backward(ctx, g_c, g_a):
        records, offsets, sorted_ids, bg, out_a, last = ctx.saved_tensors
        K, N, D0, D, width, height, per_cam, append_depth = ctx.meta
        dev = records.device
        v_rec = torch.zeros(K, N, L.REC, device=dev)
        g_c = _f32c(g_c)
        g_a = _f32c(g_a) if g_a is not None else None
        a = L.BlendBwd(K, N, D, width, height, ctx.lists, ctx.capacity, _p(records), _p(offsets), _p(sorted_ids),
                       _p(bg), _p(out_a), _p(last), _p(g_c), _p(g_a), _p(v_rec), -1, None)
        L.call("mobgs_blend_bwd", a, _stream())
        v_means2d = v_rec[..., 0:2]
        v_conics = v_rec[..., 3:6]
        v_op = v_rec[..., 2]
        v_op = v_op[0] if K == 1 else v_op.sum(0)
        v_col = v_rec[..., 6:6 + D0]
        if not per_cam:
            v_col = v_col[0] if K == 1 else v_col.sum(0)
        v_depths = v_rec[..., 6 + D0] if append_depth else None
        v_bg = None
        if ctx.needs_input_grad[4] and bg is not None:
            T_fin = (1.0 - out_a).unsqueeze(-1)
            v_bg = (g_c * T_fin).sum(dim=(1, 2))[:, :D0]
        return v_means2d, v_conics, v_col, v_op, v_bg, v_depths, None, None, None, None, None


def rasterize(means2d, conics, colors, opacities, backgrounds, depths, radii, width, height,
              append_depth=False, tight=True):
    return _Rasterize.apply(means2d, conics, colors, opacities, backgrounds, depths, radii,
                            int(width), int(height), bool(append_depth), bool(tight))
