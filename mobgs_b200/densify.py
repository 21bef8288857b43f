"""Densify / prune gather-compaction behind the reference's own method names.

    prune_optimizer(optimizer, mask)                  scene/gaussian_model.py:1044-1069  `_prune_optimizer`
    cat_tensors_to_optimizer(optimizer, tensors_dict) scene/gaussian_model.py:1094-1123
    compact_rows(tensors, idx, n_old, ext=None)       the primitive: one launch for any number of tensors

Same results, bit for bit, as the reference's per-group `param[mask]` / `torch.cat` statements — parameters,
`exp_avg`, `exp_avg_sq` and (via `extra=`) the densification statistics `prune_points` indexes next to them
(:1086-1090) — but ONE `mobgs_compact_rows` launch and one `nonzero` for the whole model instead of ~45 of each.
The optimizer is edited in place exactly as the reference does (new nn.Parameter objects, state moved to them),
so `optimizer.param_groups` / `optimizer.state` keep the layout FusedAdam and torch.optim.Adam expect.
CUDA tensors only; there is no fallback path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib


def compact_rows(tensors: Sequence[torch.Tensor], idx: torch.Tensor, n_old: int,
                 ext: Optional[Sequence[Optional[torch.Tensor]]] = None) -> List[torch.Tensor]:
    """out[t][i] = tensors[t][idx[i]] if idx[i] < n_old else ext[t][idx[i] - n_old] (zeros when ext[t] is None).
    Every tensor has n_old rows (dim 0), any trailing shape, dtype of 4 or 8 bytes; idx int64 on the device."""
    if not tensors:
        return []
    dev = tensors[0].device
    if idx.dtype != torch.int64 or not idx.is_cuda:
        raise RuntimeError("compact_rows needs a CUDA int64 index list (there is no CPU fallback)")
    idx = idx.contiguous()
    n_out = idx.numel()
    ext = list(ext) if ext is not None else [None] * len(tensors)
    chunk = _lib.load().mobgs_compact_chunk_words()
    stream = _lib.current_stream()
    outs, keep = [], []
    for s in range(0, len(tensors), _lib.COMPACT_MAX_TENSORS):
        a = _lib.CompactRows()
        part = tensors[s:s + _lib.COMPACT_MAX_TENSORS]
        a.n_tensors, a.n_old, a.n_out, a.idx = len(part), int(n_old), n_out, idx.data_ptr()
        chunks = 0
        for i, t in enumerate(part):
            if not t.is_cuda or t.device != dev or t.element_size() not in (4, 8):
                raise RuntimeError("compact_rows needs CUDA tensors of 4- or 8-byte elements on one device")
            if t.shape[0] != n_old:
                raise RuntimeError(f"tensor {s + i} has {t.shape[0]} rows, expected {n_old}")
            tc = t.detach().contiguous()
            e = ext[s + i]
            if e is not None:
                if e.dtype != t.dtype or e.shape[1:] != t.shape[1:]:
                    raise RuntimeError(f"extension {s + i}: dtype / row shape differs from its tensor")
                e = e.detach().contiguous()
            out = torch.empty((n_out,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            rw = int(torch.Size(t.shape[1:]).numel()) * t.element_size() // 4        # 32-bit words per row
            a.src[i], a.ext[i], a.dst[i] = tc.data_ptr(), (e.data_ptr() if e is not None else None), out.data_ptr()
            a.row_words[i] = max(int(rw), 1)
            a.chunk_begin[i] = chunks
            chunks += (n_out * a.row_words[i] + chunk - 1) // chunk
            keep += [tc, e]
            outs.append(out)
        a.chunk_begin[len(part)] = chunks
        _lib.call("mobgs_compact_rows", a, stream)
    del keep
    return outs


def _groups(optimizer):
    for group in optimizer.param_groups:
        if len(group["params"]) > 1 or group["name"] == "focal":      # gaussian_model.py:1047, :1097
            continue
        yield group


def _install(optimizer, group, new_param, new_state):
    """the reference's state hand-over (gaussian_model.py:1053-1056): drop the old key, re-key to the new Parameter"""
    old = group["params"][0]
    if new_state is not None:
        del optimizer.state[old]
        group["params"][0] = new_param
        optimizer.state[new_param] = new_state
    else:
        group["params"][0] = new_param
    return new_param


def _rebuild(optimizer, idx, n_old, ext_of: Dict[str, Optional[torch.Tensor]], extra, extra_ext):
    groups = list(_groups(optimizer))
    srcs, exts, slots = [], [], []
    for gi, group in enumerate(groups):
        p = group["params"][0]
        st = optimizer.state.get(p, None)
        e = ext_of.get(group["name"]) if ext_of is not None else None
        srcs.append(p); exts.append(e); slots.append((gi, "param"))
        if st is not None:
            for key in ("exp_avg", "exp_avg_sq"):
                srcs.append(st[key]); exts.append(None); slots.append((gi, key))      # the moments are extended with zeros
    n_opt = len(srcs)
    extra = list(extra or [])
    srcs += extra
    exts += list(extra_ext) if extra_ext is not None else [None] * len(extra)
    outs = compact_rows(srcs, idx, n_old, exts)
    new_state = {}
    new_param = {}
    for (gi, key), out in zip(slots, outs[:n_opt]):
        if key == "param":
            new_param[gi] = out
        else:
            new_state.setdefault(gi, {})[key] = out
    optimizable = {}
    for gi, group in enumerate(groups):
        p = group["params"][0]
        st = optimizer.state.get(p, None)
        if st is not None:
            st["exp_avg"], st["exp_avg_sq"] = new_state[gi]["exp_avg"], new_state[gi]["exp_avg_sq"]
            q = nn.Parameter(new_param[gi].requires_grad_(True))
        elif group["name"] == "current_control_num":                  # int64, never trained (:1060-1062)
            q = nn.Parameter(new_param[gi], requires_grad=False)
        else:
            q = nn.Parameter(new_param[gi].requires_grad_(True))
        optimizable[group["name"]] = _install(optimizer, group, q, st)
    return optimizable, outs[n_opt:]


@torch.no_grad()
def prune_optimizer(optimizer, mask: torch.Tensor, extra: Optional[Sequence[torch.Tensor]] = None):
    """`GaussianModel._prune_optimizer(mask)` (mask = rows to KEEP): -> {group name: new Parameter}.
    extra: further per-Gaussian tensors to compact in the same launch (prune_points' `xyz_gradient_accum`,
    `denom`, `max_radii2D`, `_deformation_accum`, `_deformation_table`, :1086-1090); when given, returns
    (optimizable_tensors, [compacted extras])."""
    idx = torch.nonzero(mask.reshape(-1), as_tuple=False).reshape(-1)
    out, ex = _rebuild(optimizer, idx, mask.numel(), None, extra, None)
    return out if extra is None else (out, ex)


@torch.no_grad()
def cat_tensors_to_optimizer(optimizer, tensors_dict: Dict[str, torch.Tensor],
                             extra: Optional[Sequence[torch.Tensor]] = None,
                             extra_ext: Optional[Sequence[Optional[torch.Tensor]]] = None):
    """`GaussianModel.cat_tensors_to_optimizer(tensors_dict)`: appends the new rows of every group (Adam moments
    extended with zeros).  -> {group name: new Parameter} (and the extended extras)."""
    groups = list(_groups(optimizer))
    n_old = groups[0]["params"][0].shape[0]
    n_new = tensors_dict[groups[0]["name"]].shape[0]
    idx = torch.arange(n_old + n_new, dtype=torch.int64, device=groups[0]["params"][0].device)
    out, ex = _rebuild(optimizer, idx, n_old, tensors_dict, extra, extra_ext)
    return out if extra is None else (out, ex)
