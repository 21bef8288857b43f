"""The fused MoBGS path: K latent sub-frames per launch.

Stage A  `synth_project`  — attribute synthesis (spline / activations / time-linear terms) +
         projection for all K sub-frames in one launch -> packed 64-byte records.
Stage B  `blend_records`  — tile binning + per-tile sort + alpha compositing of the records
         (optionally restricted to the static or the dynamic index range, which is how the
         s_render / d_render / s_alpha / d_alpha renders of render() share one projection).

The gradient between the two stages travels as one packed [K,N,16] tensor (the layout the
backward blend kernel scatters into), never as gsplat's five SoA tensors.
"""
from __future__ import annotations

from typing import Optional, Tuple

import os
import weakref

import torch

from . import _lib as L
from . import ops as _ops
from .ops import _cams, _f32c, _p, _stream, build_tile_lists, EPS2D, NEAR, FAR, RADIUS_CLIP

class GradSink:
    """Data-parallel gradient sink for `_SynthProject.backward` ("the all-reduce launched as soon as
    the last backward block retires").  While installed (`fused.GRAD_SINK = GradSink(...)`, done by
    mobgs_b200.dist.overlap_gradient_allreduce) the projection backward runs as `n_chunks` launches over Gaussian
    ranges; each range writes its parameter gradients into one contiguous block of a chunk-major buffer, and the
    block's NCCL all-reduce is issued on `comm_stream` as soon as its launch is enqueued — it overlaps the next
    range's kernel.  One mobgs_copy_segments launch then unpacks the reduced blocks into the per-parameter gradient
    tensors.  With n_chunks = 2 (default) the two ranges are "all dynamic" and "all static" Gaussians, which are
    contiguous slices of the flat gradient buffer itself: no chunk-major buffer and no unpack (_split_backward).
    `reduce(tensor)` is the collective (sum over ranks, in place, async handle or None)."""

    def __init__(self, reduce, comm_stream, n_chunks=2):
        self.reduce, self.comm_stream, self.n_chunks = reduce, comm_stream, int(n_chunks)
        self.reduced_storage = None      # untyped-storage pointer of the last flat gradient buffer delivered reduced
        self.last_collective_elems = 0


GRAD_SINK = None

# Optional allocator of the flat parameter-gradient buffer: fn(n_floats, device) -> zeroed fp32 tensor.  A data-parallel
# caller installs one that hands out a persistent NVLS symmetric-memory buffer (mobgs_b200.dist.SymmetricGradients), so
# that the gradient all-reduce is a multimem (in-switch) reduction on the very buffer the backward kernel wrote.
FLAT_ALLOCATOR = None

# Gradient-record buffers [record sets, N, 16] are kept across steps: the blend backward scatters into an all-zero buffer
# and the projection backward — its only reader — zeroes the whole buffer behind its reads (MobgsSynthBwd.zero_v_records:
# each warp clears the 2 KB chunk it has just consumed, coalesced), so the buffer comes back clean and the per-step 448 MB
# allocation + memset (1 M Gaussians, K = 7) disappears.  Only buffers handed out by `_take_grad_records` and still alive
# when they reach `_SynthProject.backward` are recycled; anything else (a user-supplied gradient, an autograd-accumulated
# sum in a fresh tensor) is read without being touched.
_GRAD_REC_POOL = {}      # (record sets, N, device index) -> clean buffer not in use
_GRAD_REC_OUT = {}       # data_ptr -> (key, weakref to the tensor handed out)
_GRAD_REC_USES = {}      # key -> number of reuses
RECYCLE_GRAD_RECORDS = os.environ.get("MOBGS_RECYCLE_GRAD_RECORDS", "1") != "0"


def _take_grad_records(Kr, N, dev):
    """an all-zero [Kr,N,16] gradient-record buffer (recycled when possible)"""
    key = (int(Kr), int(N), dev.index)
    cap = _ops.CAPTURE
    if cap is not None and RECYCLE_GRAD_RECORDS:
        # A CUDA graph owns its gradient-record buffers: an all-zero pool buffer (allocated before the capture) when
        # there is one, else zeros from the graph's own memory (re-zeroed by a captured memset on every replay).  Either
        # way the projection backward zeroes it behind its reads, so every replay starts from zeros; it is never
        # returned to the shared pool (_recycle).
        t = _GRAD_REC_POOL.pop(key, None)
        if t is None:
            t = torch.zeros(Kr, N, L.REC, device=dev)
        _GRAD_REC_OUT[t.data_ptr()] = (key, weakref.ref(t))
        cap.keep.append(t)
        return t
    t = _GRAD_REC_POOL.pop(key, None) if RECYCLE_GRAD_RECORDS else None
    if t is not None:
        _GRAD_REC_USES[key] = _GRAD_REC_USES.get(key, 0) + 1
    if t is None:
        t = torch.zeros(Kr, N, L.REC, device=dev)
    if RECYCLE_GRAD_RECORDS:
        if len(_GRAD_REC_OUT) > 16:
            for ptr in [p for p, (_k, r) in _GRAD_REC_OUT.items() if r() is None]:
                del _GRAD_REC_OUT[ptr]
        _GRAD_REC_OUT[t.data_ptr()] = (key, weakref.ref(t))
    return t


def _recyclable(g_rec) -> bool:
    """True if g_rec is a buffer of `_take_grad_records` (still the same live allocation): the projection backward may
    zero it behind its reads; `_recycle` then returns it to the pool."""
    ent = _GRAD_REC_OUT.get(g_rec.data_ptr())
    if ent is None:
        return False
    key, ref = ent
    orig = ref()
    return (orig is not None and orig.data_ptr() == g_rec.data_ptr() and tuple(g_rec.shape) == (key[0], key[1], L.REC)
            and g_rec.is_contiguous() and g_rec.device.index == key[2])


def _recycle(g_rec):
    key, _ = _GRAD_REC_OUT.pop(g_rec.data_ptr())
    if _ops.CAPTURE is None:              # (a buffer baked into a CUDA graph stays the graph's)
        _GRAD_REC_POOL[key] = g_rec


# argument tails of L.BlendFwd / L.BlendBwd for launches without the fused decoder, up to (not including) list_masks
_NO_DEC_FWD = (None, 0, None, None, None, None, 0, None, None, 0.0, 0.0, 0.0, 0.0)
_NO_DEC_BWD = (None, 0, None, None, None, None, None, None, None, 0, None, None, 0, None, None, 0.0, 0.0, 0.0, 0.0, None)


STATIC_KEYS = ("xyz", "rotation", "scaling", "opacity", "features_dc")
DYNAMIC_KEYS = ("control_xyz", "rotation", "omega", "scaling", "opacity", "features_dc", "features_t",
                "trbf_center")


def _static_struct(ts):
    xyz, rot, sc, op, fdc = ts
    return L.StaticParams(xyz.shape[0], _p(xyz), _p(rot), _p(sc), _p(op), _p(fdc))


def _dynamic_struct(ts, control_num, offset):
    ctrl, rot, om, sc, op, fdc, ft, trbf = ts
    return L.DynamicParams(ctrl.shape[0], ctrl.shape[1] if ctrl.dim() == 3 else 12, _p(ctrl), _p(control_num),
                           _p(rot), _p(om), _p(sc), _p(op), _p(fdc), _p(ft), _p(trbf), _p(offset))


class _SynthProject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, viewmats, Ks, t_spline, t_poly, control_num, width, height, want_means3d,
                s_xyz, s_rot, s_sc, s_op, s_fdc,
                d_ctrl, d_rot, d_om, d_sc, d_op, d_fdc, d_ft, d_trbf, d_off, binning=None):
        # gradients of outputs nobody used arrive as None instead of freshly zero-filled tensors (radii / depths / means3d
        # are K*N each: 56 MB of fills per backward at 1 M Gaussians, K = 7)
        ctx.set_materialize_grads(False)
        st = [_f32c(t) for t in (s_xyz, s_rot, s_sc, s_op, s_fdc)]
        dy = [_f32c(t) for t in (d_ctrl, d_rot, d_om, d_sc, d_op, d_fdc, d_ft, d_trbf)]
        viewmats, Ks = _f32c(viewmats), _f32c(Ks)
        t_spline, t_poly = _f32c(t_spline), _f32c(t_poly)
        control_num = control_num.to(torch.int64).contiguous()
        off = _f32c(d_off) if d_off is not None else None
        K = viewmats.shape[0]
        Ns, Nd = st[0].shape[0], dy[0].shape[0]
        N = Ns + Nd
        dev = viewmats.device
        records = torch.empty(K, N, L.REC, device=dev)
        radii = torch.empty(K, N, dtype=torch.int32, device=dev)
        depths = torch.empty(K, N, device=dev)
        means3d = torch.empty(K, N, 3, device=dev) if want_means3d else None
        cams = _cams(viewmats, Ks, width, height)
        # fused counting pass (ops.BinPlan): the lists the blend will walk are counted / recorded by this launch
        bin_args = ()
        if binning is not None and (binning.width, binning.height) == (width, height):
            prep = binning.prepare(N, dev)
            if prep is not None:
                lists, counts, entries, cursor = prep
                bin_args = (len(binning.specs), int(binning.tight), lists, _p(counts), _p(entries), entries.shape[0], _p(cursor))
        a = L.SynthFwd(cams, _static_struct(st), _dynamic_struct(dy, control_num, off), _p(t_spline),
                       _p(t_poly), _p(records), _p(radii), _p(depths), _p(means3d), *bin_args)
        L.call("mobgs_synth_project_fwd", a, _stream())
        ctx.save_for_backward(viewmats, Ks, t_spline, t_poly, control_num, radii, off, *st, *dy)
        ctx.size = (width, height)
        ctx.mark_non_differentiable(radii, depths)
        if means3d is None:
            means3d = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(means3d)
        return records, radii, depths, means3d

    @staticmethod
    def backward(ctx, g_rec, _g_radii, _g_depths, _g_m3d):
        viewmats, Ks, t_spline, t_poly, control_num, radii, off, *rest = ctx.saved_tensors
        st, dy = rest[:5], rest[5:]
        width, height = ctx.size
        K = viewmats.shape[0]
        dev = viewmats.device
        if g_rec is None:                     # (only radii / depths were used downstream: nothing to propagate)
            return (None,) * 23
        g_rec = _f32c(g_rec)
        recycle = _recyclable(g_rec)          # our own buffer: zero it behind the reads and keep it for the next step
        # All Gaussian-parameter gradients are views into ONE zeroed flat buffer (segments padded to 16 B):
        # autograd adopts the views as `.grad`, so a data-parallel caller all-reduces the buffer in place
        # (dist.FlatGradients) instead of packing / unpacking 12 tensors.  control_xyz needs the zeros
        # (accumulated in place); trbf_center gets no gradient: dt is detached in the reference (render():102).
        v_off = torch.empty_like(off) if off is not None else None
        v_view = torch.zeros(K, 4, 4, device=dev) if ctx.needs_input_grad[0] else None
        cams = _cams(viewmats, Ks, width, height)
        if not any(ctx.needs_input_grad[8:]):
            # pose-only backward (eval.py:120-150: every Gaussian tensor is frozen, only w2c is optimised): no flat
            # gradient buffer, no parameter-gradient stores
            if v_view is not None:
                a = L.SynthBwd(cams, _static_struct(st), _dynamic_struct(dy, control_num, off), _p(t_spline),
                               _p(t_poly), _p(radii), _p(g_rec), *([None] * 13), _p(v_view), 0, 0, int(recycle))
                L.call("mobgs_synth_project_bwd", a, _stream())
                if recycle:
                    _recycle(g_rec)
            return (v_view,) + (None,) * 22
        outs = list(st) + list(dy[:7])
        starts, tot = [], 0
        for t in outs:
            starts.append(tot)
            tot += (t.numel() + 3) // 4 * 4
        flat = FLAT_ALLOCATOR(tot, dev) if FLAT_ALLOCATOR is not None else torch.zeros(tot, device=dev)
        views = [flat[o:o + t.numel()].view(t.shape) for o, t in zip(starts, outs)]
        v_st, v_dy = views[:5], views[5:] + [None]
        sink = GRAD_SINK
        if sink is not None and off is None and sink.n_chunks > 1:
            if sink.n_chunks == 2:
                _split_backward(sink, cams, st, dy, control_num, t_spline, t_poly, radii, g_rec, flat, starts, views, v_view,
                                int(recycle))
            else:
                _chunked_backward(sink, cams, st, dy, control_num, t_spline, t_poly, radii, g_rec, outs, views, v_view,
                                  int(recycle))
            if recycle:
                _recycle(g_rec)
            sink.reduced_storage = flat.untyped_storage().data_ptr()
            return (v_view, None, None, None, None, None, None, None, *v_st, *v_dy[:7], None, None, None)
        a = L.SynthBwd(cams, _static_struct(st), _dynamic_struct(dy, control_num, off), _p(t_spline),
                       _p(t_poly), _p(radii), _p(g_rec),
                       _p(v_st[0]), _p(v_st[1]), _p(v_st[2]), _p(v_st[3]), _p(v_st[4]),
                       _p(v_dy[0]), _p(v_dy[1]), _p(v_dy[2]), _p(v_dy[3]), _p(v_dy[4]), _p(v_dy[5]),
                       _p(v_dy[6]), _p(v_off), _p(v_view), 0, 0, int(recycle))
        L.call("mobgs_synth_project_bwd", a, _stream())
        if recycle:
            _recycle(g_rec)
        return (v_view, None, None, None, None, None, None, None, *v_st, *v_dy[:7], None, v_off, None)


def _split_backward(sink, cams, st, dy, control_num, t_spline, t_poly, radii, g_rec, flat, starts, views, v_view, zero_rec=0):
    """GradSink with two ranges that need no repacking: the flat gradient buffer holds the 5 static tensors, then
    the 7 dynamic ones, so "all dynamic Gaussians" and "all static Gaussians" are each one contiguous slice of it.
    The dynamic range runs first; its all-reduce (59 % of the bytes at 700 k / 300 k) overlaps the static range's
    kernel; only the static slice's collective is exposed."""
    Ns, Nd = st[0].shape[0], dy[0].shape[0]
    main = torch.cuda.current_stream()
    stream = _stream()
    split = starts[5] if len(starts) > 5 else flat.numel()          # first float of the dynamic tensors
    works = []
    for kind in ("d", "s"):
        lo, hi = (Ns, Ns + Nd) if kind == "d" else (0, Ns)
        if hi <= lo:
            continue
        ptr = [_p(v) for v in views]
        a = L.SynthBwd(cams, _static_struct(st), _dynamic_struct(dy, control_num, None), _p(t_spline), _p(t_poly),
                       _p(radii), _p(g_rec), *ptr[:5], *ptr[5:12], None, _p(v_view), lo, hi, zero_rec)
        L.call("mobgs_synth_project_bwd", a, stream)
        ev = torch.cuda.Event()
        ev.record(main)
        part = flat[split:] if kind == "d" else flat[:split]
        with torch.cuda.stream(sink.comm_stream):
            sink.comm_stream.wait_event(ev)
            works.append(sink.reduce(part))
    sink.last_collective_elems = flat.numel()
    for w in works:
        if w is not None:
            w.wait()
    main.wait_stream(sink.comm_stream)
    flat.record_stream(sink.comm_stream)


def _chunked_backward(sink, cams, st, dy, control_num, t_spline, t_poly, radii, g_rec, outs, views, v_view, zero_rec=0):
    """see GradSink.  outs = the 12 parameter tensors (5 static, 7 dynamic), views = their gradient tensors."""
    dev = g_rec.device
    Ns, Nd = st[0].shape[0], dy[0].shape[0]
    nc = sink.n_chunks
    n_s = max(1, min(nc, round(nc * 17 * Ns / max(1, 17 * Ns + 57 * Nd)))) if Ns > 0 else 0     # chunks ~ equal bytes
    n_d = max(1, nc - n_s) if Nd > 0 else 0
    ranges = [("s", Ns * i // n_s, Ns * (i + 1) // n_s) for i in range(n_s)] + \
             [("d", Nd * i // n_d, Nd * (i + 1) // n_d) for i in range(n_d)]
    ranges = [r for r in ranges if r[2] > r[1]]
    row = [int(t[0].numel()) if t.shape[0] else int(torch.Size(t.shape[1:]).numel()) for t in outs]
    # chunk-major buffer: per range, the rows of its 5 (static) or 7 (dynamic) tensors back to back (16-byte aligned)
    plan, tot = [], 0
    for kind, lo, hi in ranges:
        idx = range(0, 5) if kind == "s" else range(5, 12)
        blocks, start = [], tot
        for t in idx:
            blocks.append((t, tot))
            tot += ((hi - lo) * row[t] + 3) // 4 * 4
        plan.append((kind, lo, hi, blocks, start, tot))
    cm = torch.zeros(tot, device=dev)
    main = torch.cuda.current_stream()
    works = []
    stream = _stream()
    base_ptr = cm.data_ptr()
    for kind, lo, hi, blocks, start, end in plan:
        ptr = [None] * 12
        for t, o in blocks:
            ptr[t] = base_ptr + 4 * (o - lo * row[t])          # "row g of the full tensor" addressing
        g_lo, g_hi = (lo, hi) if kind == "s" else (Ns + lo, Ns + hi)
        a = L.SynthBwd(cams, _static_struct(st), _dynamic_struct(dy, control_num, None), _p(t_spline), _p(t_poly),
                       _p(radii), _p(g_rec), *ptr[:5], *ptr[5:12], None, _p(v_view), g_lo, g_hi, zero_rec)
        L.call("mobgs_synth_project_bwd", a, stream)
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(sink.comm_stream):
            sink.comm_stream.wait_event(ev)
            works.append(sink.reduce(cm[start:end]))
    sink.last_collective_elems = tot
    for w in works:
        if w is not None:
            w.wait()                                  # the main stream waits for the collective (no host block)
    main.wait_stream(sink.comm_stream)
    cm.record_stream(sink.comm_stream)
    # unpack: block (tensor t, rows lo..hi) -> views[t][lo:hi]
    chunk = L.load().mobgs_compact_chunk_words()
    seg = [(base_ptr + 4 * o, views[t].data_ptr() + 4 * lo * row[t], (hi - lo) * row[t])
           for kind, lo, hi, blocks, _s, _e in plan for t, o in blocks]
    for s0 in range(0, len(seg), L.COPY_MAX_SEGMENTS):
        part = seg[s0:s0 + L.COPY_MAX_SEGMENTS]
        c = L.CopySegments()
        c.n_segments = len(part)
        chunks = 0
        for i, (src, dst, nw) in enumerate(part):
            c.src[i], c.dst[i], c.n_words[i], c.chunk_begin[i] = src, dst, nw, chunks
            chunks += (nw + chunk - 1) // chunk
        c.chunk_begin[len(part)] = chunks
        L.call("mobgs_copy_segments", c, stream)


def synth_project(static_params, dynamic_params, control_num, viewmats, Ks, t_spline, t_poly,
                  width, height, offset=None, want_means3d=False, binning=None):
    """static_params: (xyz[Ns,3], rotation[Ns,4], scaling[Ns,3], opacity[Ns(,1)], features_dc[Ns,6]);
    dynamic_params: (control_xyz[Nd,P,3], rotation, omega, scaling, opacity, features_dc,
    features_t[Nd,3], trbf_center[Nd(,1)]); control_num int64 [Nd(,1)].
    binning: optional ops.BinPlan — the tile-counting pass of the lists it names is fused into this launch.
    -> records [K,N,16], radii i32 [K,N], depths [K,N], means3d [K,N,3] or empty."""
    return _SynthProject.apply(viewmats, Ks, t_spline, t_poly, control_num, int(width), int(height),
                               bool(want_means3d), *static_params, *dynamic_params, offset, binning)


class _BlendRecords(torch.autograd.Function):
    @staticmethod
    def forward(ctx, records, radii, depths, backgrounds, vsp, D, width, height, specs, tight, vsp_list, tile_list=None):
        ctx.set_materialize_grads(False)
        records = _f32c(records)
        Kr, N = radii.shape
        dev = records.device
        K = Kr if specs is None else len(specs)
        bg = _f32c(backgrounds) if backgrounds is not None else None

        def blend(lists):
            out_c = torch.empty(K, height, width, D, device=dev)
            out_a = torch.empty(K, height, width, device=dev)
            last = torch.empty(K, height, width, dtype=torch.int32, device=dev)
            masks = torch.empty(lists.capacity, dtype=torch.int16, device=dev)       # unit masks, kept for the backward
            a = L.BlendFwd(K, N, D, width, height, lists.lists, lists.capacity, _p(records), _p(lists.tile_offsets),
                           _p(lists.sorted_ids), _p(bg), _p(out_c), _p(out_a), _p(last), *_NO_DEC_FWD, _p(masks))
            L.call("mobgs_blend_fwd", a, _stream())
            return out_c, out_a, last, masks

        lists, (out_c, out_a, last, masks) = build_tile_lists(records, radii, depths, width, height, tight, specs,
                                                              consume=blend, tile_list=tile_list)
        ctx.save_for_backward(records, lists.tile_offsets, lists.sorted_ids, bg, out_a, last, masks)
        ctx.lists, ctx.capacity = lists.lists, lists.capacity
        ctx.meta = (K, Kr, N, D, width, height, vsp_list, vsp is not None)
        ctx.n_isect = lists.n_isect
        ctx.mark_non_differentiable(last)
        return out_c, out_a, last

    @staticmethod
    def backward(ctx, g_c, g_a, _g_last):
        records, offsets, sorted_ids, bg, out_a, last, masks = ctx.saved_tensors
        K, Kr, N, D, width, height, vsp_list, has_vsp = ctx.meta
        dev = records.device
        g_c = _f32c(g_c) if g_c is not None else torch.zeros(K, height, width, D, device=dev)
        g_a = _f32c(g_a) if g_a is not None else None
        v_rec = _take_grad_records(Kr, N, dev)
        v_vsp = torch.zeros(1, N, 2, device=dev) if has_vsp else None
        a = L.BlendBwd(K, N, D, width, height, ctx.lists, ctx.capacity, _p(records), _p(offsets), _p(sorted_ids),
                       _p(bg), _p(out_a), _p(last), _p(g_c), _p(g_a), _p(v_rec), vsp_list if has_vsp else -1, _p(v_vsp),
                       *_NO_DEC_BWD, _p(masks))
        L.call("mobgs_blend_bwd", a, _stream())
        return v_rec, None, None, None, v_vsp, None, None, None, None, None, None, None


def blend_records(records, radii, depths, backgrounds, D, width, height, specs=None, g_range=None, tight=True,
                  vsp: Optional[torch.Tensor] = None, vsp_k: int = 0, tile_list=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (colors [K,H,W,D] incl. background, alphas [K,H,W]) for K lists.

    specs: [(record_set, g_begin, g_end)] per list (default: one full-range list per record set;
    `g_range=(g0, g1)` is shorthand for restricting every record set to one range).
    `vsp` is an optional leaf [1,N,2] that receives d loss / d means2d of list `vsp_k` as its .grad
    (densification statistics, reference train.py:634-648).  `tile_list`: lists naming the same value share one
    tile binning / sort (same geometry and range, different colours; ops.build_tile_lists)."""
    if specs is None and g_range is not None:
        specs = [(k, g_range[0], g_range[1]) for k in range(radii.shape[0])]
    if specs is not None:
        specs = tuple(tuple(int(v) for v in s) for s in specs)
    out_c, out_a, _ = _BlendRecords.apply(records, radii, depths, backgrounds, vsp, int(D), int(width),
                                          int(height), specs, bool(tight), int(vsp_k),
                                          None if tile_list is None else tuple(int(t) for t in tile_list))
    return out_c, out_a


class _Decode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img10, alpha, rays, w1, w2, want_mean):
        img10, alpha, rays, w1, w2 = (_f32c(t) for t in (img10, alpha, rays, w1, w2))
        K, H, W, D = img10.shape
        assert D == 10 and rays.shape[1] == 6 and rays.shape[0] in (1, K)
        per_k = int(rays.shape[0] == K and K > 1)
        dev = img10.device
        rgb = torch.empty(K, 3, H, W, device=dev)
        depth = torch.empty(K, H, W, device=dev)
        mean = torch.empty(3, H, W, device=dev) if want_mean else None
        a = L.DecodeFwd(K, W, H, _p(img10), _p(alpha), _p(rays), per_k, _p(w1), _p(w2), _p(rgb), _p(depth), _p(mean))
        L.call("mobgs_decode_fwd", a, _stream())
        ctx.save_for_backward(img10, alpha, rays, w1, w2)
        ctx.per_k = per_k
        if mean is None:
            mean = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(mean)
        return rgb, depth, mean

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_mean):
        img10, alpha, rays, w1, w2 = ctx.saved_tensors
        K, H, W, _ = img10.shape
        dev = img10.device
        g_rgb = _f32c(g_rgb) if g_rgb is not None else None
        g_depth = _f32c(g_depth) if g_depth is not None else None
        g_mean = _f32c(g_mean) if (g_mean is not None and g_mean.numel() > 0) else None
        v_img = torch.empty_like(img10)
        v_alpha = torch.empty_like(alpha)
        v_rays = None
        if ctx.needs_input_grad[2]:
            v_rays = torch.empty_like(rays) if ctx.per_k else torch.zeros_like(rays)
        v_w1, v_w2 = torch.zeros_like(w1), torch.zeros_like(w2)
        a = L.DecodeBwd(K, W, H, _p(img10), _p(alpha), _p(rays), ctx.per_k, _p(w1), _p(w2), _p(g_rgb),
                        _p(g_depth), _p(g_mean), _p(v_img), _p(v_alpha), _p(v_rays), _p(v_w1), _p(v_w2))
        L.call("mobgs_decode_bwd", a, _stream())
        return v_img, v_alpha, v_rays, v_w1, v_w2, None


def decode(img10, alpha, rays, w1, w2, want_mean=False):
    """img10 [K,H,W,10], alpha [K,H,W], rays [K|1,6,H,W], w1 [6,12], w2 [3,6]
    -> rgb [K,3,H,W], expected depth [K,H,W], blur-model mean [3,H,W] (or empty)."""
    return _Decode.apply(img10, alpha, rays, w1, w2, bool(want_mean))


class _BlendDecode(torch.autograd.Function):
    """Stage B + epilogue in one kernel pair: binning + sort + blend with the expected-depth
    division and the Sandwich decoder applied while the pixel is in registers (forward), and their
    VJP evaluated in the blend backward's prologue — img10 gradients never exist in HBM."""

    @staticmethod
    def forward(ctx, records, radii, depths, backgrounds, vsp, rays, w1, w2, width, height, specs, tight,
                vsp_list, want_mean, mean_K=0, flow_ref=-1, ray_intr=None, plan=None):
        # Unused outputs' gradients arrive as None, not as zero-filled tensors: a step whose loss reads only the blur mean
        # would otherwise allocate, fill and then READ (in the backward prologue) zeros for rgb [K,3,H,W], depth and
        # alpha — 290 MB each way at 1080p, K = 7.
        ctx.set_materialize_grads(False)
        records = _f32c(records)
        rays, w1, w2 = _f32c(rays), _f32c(w1), _f32c(w2)
        Kr, N = radii.shape
        dev = records.device
        K = Kr if specs is None else len(specs)
        # ray_intr = (ppx, ppy, sfx, sfy): `rays` is then the [n_cam,12] pose tensor (R row-major | centre) and every
        # pixel's ray is generated in registers (MobgsBlendFwd.dec_pose) instead of read from a [n_cam,6,H,W] image
        assert rays.shape[1] == (6 if ray_intr is None else 12)
        if rays.shape[0] == 1:
            per_k = 0
        elif rays.shape[0] == K:
            per_k = 1
        elif rays.shape[0] == Kr:
            per_k = 2          # one camera per record set, shared by the lists over it
        else:
            raise ValueError(f"rays has {rays.shape[0]} cameras for {K} lists over {Kr} record sets")
        mK = mean_K if mean_K > 0 else K
        bg = _f32c(backgrounds) if backgrounds is not None else None

        def blend(lists):
            img10 = torch.empty(K, height, width, 10, device=dev)
            alpha = torch.empty(K, height, width, device=dev)
            last = torch.empty(K, height, width, dtype=torch.int32, device=dev)
            rgb = torch.empty(K, 3, height, width, device=dev)
            depth = torch.empty(K, height, width, device=dev)
            flow = torch.empty(K, height, width, 2, device=dev) if flow_ref >= 0 else None
            masks = torch.empty(lists.capacity, dtype=torch.int16, device=dev)       # unit masks, kept for the backward
            a = L.BlendFwd(K, N, 10, width, height, lists.lists, lists.capacity, _p(records), _p(lists.tile_offsets),
                           _p(lists.sorted_ids), _p(bg), _p(img10), _p(alpha), _p(last),
                           _p(rays) if ray_intr is None else None, per_k, _p(w1), _p(w2), _p(rgb), _p(depth),
                           max(flow_ref, 0), _p(flow), *((None, 0.0, 0.0, 0.0, 0.0) if ray_intr is None else
                                                         (_p(rays), *ray_intr)), _p(masks))
            L.call("mobgs_blend_fwd", a, _stream())
            return img10, alpha, last, rgb, depth, flow, masks

        lists, (img10, alpha, last, rgb, depth, flow, masks) = build_tile_lists(records, radii, depths, width, height, tight,
                                                                                specs, consume=blend, plan=plan)
        if want_mean:
            mean = torch.empty(3, height, width, device=dev)
            L.subframe_mean(_p(rgb), _p(mean), mK, 3 * height * width, _stream())
        else:
            mean = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(mean)
        ctx.save_for_backward(records, lists.tile_offsets, lists.sorted_ids, bg, img10, alpha, last, rays, w1, w2, masks)
        ctx.lists, ctx.capacity = lists.lists, lists.capacity
        ctx.meta = (K, Kr, N, width, height, vsp_list, vsp is not None, per_k, mK, flow_ref, ray_intr)
        ctx.n_isect = lists.n_isect
        if flow is None:
            flow = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(flow)
        return rgb, depth, alpha, mean, flow

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_alpha, g_mean, g_flow):
        records, offsets, sorted_ids, bg, img10, alpha, last, rays, w1, w2, masks = ctx.saved_tensors
        K, Kr, N, width, height, vsp_list, has_vsp, per_k, mK, flow_ref, ray_intr = ctx.meta
        if flow_ref >= 0:
            g_flow = _f32c(g_flow) if g_flow is not None else torch.zeros(K, height, width, 2, device=records.device)
        else:
            g_flow = None
        dev = records.device
        g_rgb = _f32c(g_rgb) if g_rgb is not None else None
        g_depth = _f32c(g_depth) if g_depth is not None else None
        g_alpha = _f32c(g_alpha) if g_alpha is not None else None
        g_mean = _f32c(g_mean) if (g_mean is not None and g_mean.numel() > 0) else None
        v_rec = _take_grad_records(Kr, N, dev)
        v_vsp = torch.zeros(1, N, 2, device=dev) if has_vsp else None
        v_rays = v_pose = None
        # No gradient through the blur mean: this is a pass that reaches few of the lists (train.py's second backward,
        # :680, touches the centre render's depth / alphas / image only) — tiles test their upstream gradients first and
        # leave before the decoder prologue (MobgsBlendBwd.sparse_grads).
        sparse = g_mean is None and K > 1
        if ctx.needs_input_grad[5]:
            if ray_intr is not None:
                v_pose = torch.zeros(rays.shape[0], L.POSE_SLOTS, 12, device=dev)
            else:
                v_rays = torch.empty_like(rays) if (per_k == 1 and not sparse) else torch.zeros_like(rays)
        v_wp = torch.zeros(L.DEC_SLOTS, 90, device=dev)
        a = L.BlendBwd(K, N, 10, width, height, ctx.lists, ctx.capacity, _p(records), _p(offsets), _p(sorted_ids),
                       _p(bg), _p(alpha), _p(last), None, None, _p(v_rec), vsp_list if has_vsp else -1, _p(v_vsp),
                       _p(rays) if ray_intr is None else None, per_k, _p(w1), _p(w2), _p(img10), _p(g_rgb), _p(g_depth),
                       _p(g_alpha), _p(g_mean), mK, _p(v_rays), _p(v_wp), max(flow_ref, 0), _p(g_flow),
                       *((None, 0.0, 0.0, 0.0, 0.0, None) if ray_intr is None else (_p(rays), *ray_intr, _p(v_pose))),
                       _p(masks), int(sparse))
        L.call("mobgs_blend_bwd", a, _stream())
        if v_pose is not None:
            v_rays = v_pose.sum(1)
        v_w = v_wp.sum(0)
        return (v_rec, None, None, None, v_vsp, v_rays, v_w[:72].reshape(6, 12), v_w[72:].reshape(3, 6),
                None, None, None, None, None, None, None, None, None, None)


def blend_decode(records, radii, depths, backgrounds, rays, w1, w2, width, height, specs=None, tight=True,
                 vsp: Optional[torch.Tensor] = None, vsp_k: int = 0, want_mean: bool = False, mean_K: int = 0,
                 flow_ref: Optional[int] = None, plan=None):
    """-> (rgb [K,3,H,W], expected depth [K,H,W], alpha [K,H,W], mean [3,H,W] or empty[, flow [K,H,W,2]]).
    rays: [1,6,H,W] shared, [K,6,H,W] per list, or [record sets,6,H,W] per projection — or a
    mobgs_b200.cameras.RayPose (same leading-axis conventions, 12 pose floats per camera): the kernels then generate
    every pixel's ray in registers and reduce the pose gradient themselves, no ray image exists.
    mean_K: the blur mean is taken over the first mean_K lists (0 = all).
    flow_ref: record set whose projected means define two extra colour channels composited in the same walk,
    records[flow_ref][g].xy - records[own set][g].xy, no background (get_flow's exp2mid render, renderer :426-441);
    when given, a fifth output `flow` is returned."""
    if specs is not None:
        specs = tuple(tuple(int(v) for v in s) for s in specs)
    ray_intr = None
    if not isinstance(rays, torch.Tensor):          # cameras.RayPose
        rays, ray_intr = rays.pose, tuple(float(v) for v in rays.intr)
    out = _BlendDecode.apply(records, radii, depths, backgrounds, vsp, rays, w1, w2, int(width), int(height),
                             specs, bool(tight), int(vsp_k), bool(want_mean), int(mean_K),
                             -1 if flow_ref is None else int(flow_ref), ray_intr, plan)
    return out if flow_ref is not None else out[:4]


class _FlowRecords(torch.autograd.Function):
    """records [K+1,N,16] (set 0 = mid time, 1..K = exposure times) -> flow records [2K,N,16]."""

    @staticmethod
    def forward(ctx, records):
        records = _f32c(records)
        K, N = records.shape[0] - 1, records.shape[1]
        out = torch.empty(2 * K, N, L.REC, device=records.device)
        L.call("mobgs_flow_records_fwd", L.FlowRecFwd(K, N, _p(records), _p(out)), _stream())
        ctx.shape = (K, N)
        return out

    @staticmethod
    def backward(ctx, g):
        K, N = ctx.shape
        g = _f32c(g)
        v = torch.empty(K + 1, N, L.REC, device=g.device)
        L.call("mobgs_flow_records_bwd", L.FlowRecBwd(K, N, _p(g), _p(v), 0), _stream())
        return v


def flow_records(records):
    return _FlowRecords.apply(records)


class _MidFlowRecords(torch.autograd.Function):
    """records [K+1,N,16] (set 0 = mid time, 1..K = exposure times) -> [ceil(2K/10),N,16]: mid geometry carrying
    all 2K mid2exp flow channels, ten per record set (mobgs_midflow_records_fwd)."""

    @staticmethod
    def forward(ctx, records):
        records = _f32c(records)
        K, N = records.shape[0] - 1, records.shape[1]
        M = (2 * K + 9) // 10
        out = torch.empty(M, N, L.REC, device=records.device)
        L.call("mobgs_midflow_records_fwd", L.FlowRecFwd(K, N, _p(records), _p(out)), _stream())
        ctx.shape = (K, N)
        return out

    @staticmethod
    def backward(ctx, g):
        K, N = ctx.shape
        g = _f32c(g)
        v = torch.empty(K + 1, N, L.REC, device=g.device)
        L.call("mobgs_midflow_records_bwd", L.FlowRecBwd(K, N, _p(g), _p(v), 0), _stream())
        return v


def midflow_records(records):
    return _MidFlowRecords.apply(records)


class _FlowRender(torch.autograd.Function):
    """All rasterisations of K get_flow() calls that share ONE projection (records [K+1,N,16]: set 0 = mid time,
    sets 1..K = exposure times) as ONE autograd node, so their three backward passes accumulate into a single
    gradient-record buffer instead of three that autograd then has to add (0.5 GB each at the headline size):

      exposure lists  (k+1, all)      latent image (decoded) + exp2mid flow, one walk      gaussian_renderer :437, :473
      dynamic lists   (k+1, Ns..N)    latent alpha, one channel                             :379
      mid lists       (0, all) x M    all 2K mid2exp flow channels over one shared binning  :456

    -> rgb [K,3,H,W], flow_e2m [K,H,W,2], alpha_dyn [K,H,W], mid [M,H,W,10]."""

    @staticmethod
    def forward(ctx, records, radii, depths, bg10, rays, w1, w2, width, height, Ns, tight, ray_intr=None):
        ctx.set_materialize_grads(False)
        records = _f32c(records)
        rays, w1, w2, bg10 = _f32c(rays), _f32c(w1), _f32c(w2), _f32c(bg10)
        Kr, N = radii.shape
        K = Kr - 1
        dev = records.device
        assert rays.shape[0] == 1 and rays.shape[1] == (6 if ray_intr is None else 12)   # ray image or pose (_BlendDecode)
        st = _stream()

        def exp_walk(lists):
            img10 = torch.empty(K, height, width, 10, device=dev)
            alpha = torch.empty(K, height, width, device=dev)
            last = torch.empty(K, height, width, dtype=torch.int32, device=dev)
            rgb = torch.empty(K, 3, height, width, device=dev)
            flow = torch.empty(K, height, width, 2, device=dev)
            masks = torch.empty(lists.capacity, dtype=torch.int16, device=dev)
            a = L.BlendFwd(K, N, 10, width, height, lists.lists, lists.capacity, _p(records), _p(lists.tile_offsets),
                           _p(lists.sorted_ids), _p(bg10), _p(img10), _p(alpha), _p(last),
                           _p(rays) if ray_intr is None else None, 0, _p(w1), _p(w2), _p(rgb), None, 0, _p(flow),
                           *((None, 0.0, 0.0, 0.0, 0.0) if ray_intr is None else (_p(rays), *ray_intr)), _p(masks))
            L.call("mobgs_blend_fwd", a, st)
            return img10, alpha, last, rgb, flow, masks

        le, (img10, alpha_e, last_e, rgb, flow, masks_e) = build_tile_lists(
            records, radii, depths, width, height, tight, tuple((k + 1, 0, N) for k in range(K)), consume=exp_walk)

        def plain_walk(recs, D):
            def walk(lists):
                Kl = len(lists.specs)
                out_c = torch.empty(Kl, height, width, D, device=dev)
                out_a = torch.empty(Kl, height, width, device=dev)
                last = torch.empty(Kl, height, width, dtype=torch.int32, device=dev)
                masks = torch.empty(lists.capacity, dtype=torch.int16, device=dev)
                a = L.BlendFwd(Kl, N, D, width, height, lists.lists, lists.capacity, _p(recs), _p(lists.tile_offsets),
                               _p(lists.sorted_ids), None, _p(out_c), _p(out_a), _p(last), *_NO_DEC_FWD, _p(masks))
                L.call("mobgs_blend_fwd", a, st)
                return out_c, out_a, last, masks
            return walk

        ld, (_junk, alpha_d, last_d, masks_d) = build_tile_lists(
            records, radii, depths, width, height, tight, tuple((k + 1, Ns, N) for k in range(K)), consume=plain_walk(records, 1))

        M = (2 * K + 9) // 10
        frec = torch.empty(M, N, L.REC, device=dev)
        L.call("mobgs_midflow_records_fwd", L.FlowRecFwd(K, N, _p(records), _p(frec)), st)
        radii_m, depths_m = radii[0:1].expand(M, -1).contiguous(), depths[0:1].expand(M, -1).contiguous()
        lm, (mid, alpha_m, last_m, masks_m) = build_tile_lists(
            frec, radii_m, depths_m, width, height, tight, tuple((m, 0, N) for m in range(M)),
            consume=plain_walk(frec, 10), tile_list=(0,) * M)

        ctx.save_for_backward(records, bg10, rays, w1, w2, img10, alpha_e, last_e, le.tile_offsets, le.sorted_ids,
                              alpha_d, last_d, ld.tile_offsets, ld.sorted_ids, frec, alpha_m, last_m, lm.tile_offsets,
                              lm.sorted_ids, masks_e, masks_d, masks_m)
        ctx.lists = (le.lists, le.capacity, ld.lists, ld.capacity, lm.lists, lm.capacity)
        ctx.meta = (K, M, N, width, height, ray_intr)
        return rgb, flow, alpha_d, mid

    @staticmethod
    def backward(ctx, g_rgb, g_flow, g_alpha_d, g_mid):
        (records, bg10, rays, w1, w2, img10, alpha_e, last_e, off_e, ids_e, alpha_d, last_d, off_d, ids_d, frec, alpha_m,
         last_m, off_m, ids_m, masks_e, masks_d, masks_m) = ctx.saved_tensors
        le, cap_e, ld, cap_d, lm, cap_m = ctx.lists
        K, M, N, width, height, ray_intr = ctx.meta
        dev = records.device
        st = _stream()
        v_rec = _take_grad_records(K + 1, N, dev)
        v_rays = v_pose = None
        if ctx.needs_input_grad[4]:
            if ray_intr is None:
                v_rays = torch.zeros_like(rays)
            else:
                v_pose = torch.zeros(1, L.POSE_SLOTS, 12, device=dev)
        v_wp = torch.zeros(L.DEC_SLOTS, 90, device=dev)
        g_rgb = _f32c(g_rgb) if g_rgb is not None else None
        g_flow = _f32c(g_flow) if g_flow is not None else torch.zeros(K, height, width, 2, device=dev)
        a = L.BlendBwd(K, N, 10, width, height, le, cap_e, _p(records), _p(off_e), _p(ids_e), _p(bg10), _p(alpha_e),
                       _p(last_e), None, None, _p(v_rec), -1, None, _p(rays) if ray_intr is None else None, 0, _p(w1),
                       _p(w2), _p(img10), _p(g_rgb), None, None, None, 0, _p(v_rays), _p(v_wp), 0, _p(g_flow),
                       *((None, 0.0, 0.0, 0.0, 0.0, None) if ray_intr is None else (_p(rays), *ray_intr, _p(v_pose))),
                       _p(masks_e))
        L.call("mobgs_blend_bwd", a, st)
        if v_pose is not None:
            v_rays = v_pose.sum(1)
        if g_alpha_d is not None:
            zeros = torch.zeros(K, height, width, 1, device=dev)
            a = L.BlendBwd(K, N, 1, width, height, ld, cap_d, _p(records), _p(off_d), _p(ids_d), None, _p(alpha_d),
                           _p(last_d), _p(zeros), _p(_f32c(g_alpha_d)), _p(v_rec), -1, None, *_NO_DEC_BWD, _p(masks_d))
            L.call("mobgs_blend_bwd", a, st)
        if g_mid is not None:
            v_frec = torch.zeros(M, N, L.REC, device=dev)
            a = L.BlendBwd(M, N, 10, width, height, lm, cap_m, _p(frec), _p(off_m), _p(ids_m), None, _p(alpha_m),
                           _p(last_m), _p(_f32c(g_mid)), None, _p(v_frec), -1, None, *_NO_DEC_BWD, _p(masks_m))
            L.call("mobgs_blend_bwd", a, st)
            L.call("mobgs_midflow_records_bwd", L.FlowRecBwd(K, N, _p(v_frec), _p(v_rec), 1), st)
        v_w = v_wp.sum(0)
        return (v_rec, None, None, None, v_rays, v_w[:72].reshape(6, 12), v_w[72:].reshape(3, 6), None, None, None, None, None)


def flow_render(records, radii, depths, bg10, rays, w1, w2, width, height, n_static, tight=True):
    """see _FlowRender; bg10 [K,10] (the exposure lists' background), rays [1,6,H,W] (one camera for all lists) or a
    one-camera mobgs_b200.cameras.RayPose."""
    ray_intr = None
    if not isinstance(rays, torch.Tensor):
        rays, ray_intr = rays.pose, tuple(float(v) for v in rays.intr)
    return _FlowRender.apply(records, radii, depths, bg10, rays, w1, w2, int(width), int(height), int(n_static), bool(tight),
                             ray_intr)
