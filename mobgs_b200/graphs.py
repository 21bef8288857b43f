"""CUDA-graphed training / evaluation step (host path).

At the reference's real training shape (150 k Gaussians, 512x288, K = 9) one blurry-view step is ~1.2 ms of kernels
under ~1.3 ms of Python, autograd-engine and launch overhead: the step is host-bound.  `GraphedStep` captures the whole
step — the K-batched forward, the loss and the backward, ~25 launches — into ONE CUDA graph per launch shape and
replays it with new inputs copied into static buffers, so the host cost of a step is a few small copies and one
`cudaGraphLaunch`.

What makes the path capturable:
  * every kernel is launched through the C ABI on torch's current stream (the capture stream inside
    `torch.cuda.graph`), none synchronises, and all workspaces come from torch's allocator (the graph's private pool);
  * the one host round trip of the eager path — reading the intersection count back to size the tile lists — is
    already speculative (ops.build_tile_lists): under capture the lists are sized from the last count seen for the
    shape (+25 %), the count is copied to a pinned word by a captured memcpy, and the "did it fit" check runs AFTER
    the replay (`validate`).  An overflowed replay produced results from truncated lists; `validate` then recomputes
    the step eagerly (exact, as the eager path's own redo does), hands those results out instead and re-captures with
    the larger capacity — results are always those of exactly-sized lists;
  * the gradient-record buffer a graph scatters into is owned by the graph (fused._take_grad_records).

Use:
    step = GraphedStep(fn, inputs, params)         # fn(*inputs) -> tensor or tuple of tensors; runs fwd (+ loss.backward())
    outs = step(*new_inputs)                       # replay; outputs / parameter .grad are static tensors, valid until the next call
    step.validate()                                # before consuming them when exactness on list overflow matters
                                                   # (the next call validates the previous one automatically)
`fn` must take every per-step value as a CUDA tensor argument (Python numbers are frozen into the graph), must not
read tensors back to the host and must produce gradients only in `params`' .grad.
"""
from __future__ import annotations

import gc
from typing import Callable, Sequence

import torch

from . import _lib as L
from . import ops


class _Capture:
    """what ops / fused hand to the GraphedStep that is capturing (ops.CAPTURE)"""

    def __init__(self, n_words: int = 64):
        self._pinned = torch.zeros(n_words, dtype=torch.int32).pin_memory()      # allocated BEFORE the capture starts
        self._used = 0
        self.checks = []          # (pinned count word, list capacity baked into the graph, _CAP_CACHE key)
        self.keep = []            # buffers the graph owns (gradient records)

    def pinned_word(self) -> torch.Tensor:
        if self._used >= self._pinned.numel():
            raise RuntimeError("too many tile binnings in one captured step")
        w = self._pinned[self._used:self._used + 1]
        self._used += 1
        return w


class GraphedStep:
    def __init__(self, fn: Callable, inputs: Sequence[torch.Tensor], params: Sequence[torch.Tensor] = (),
                 warmup: int = 2):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedStep needs a CUDA device (there is no CPU path)")
        self.fn = fn
        self.params = [p for p in params]
        dev = torch.device("cuda", torch.cuda.current_device())
        # static device copies (inputs may be pinned host tensors: a replay then starts with their H2D copies)
        self.static_inputs = [torch.empty_like(t, device=dev).copy_(t.detach()).requires_grad_(t.requires_grad)
                              for t in inputs]
        self.warmup = max(1, int(warmup))
        self.replays = self.overflows = self.captures = 0
        self._pending = False
        self._graph = None
        self._stream = None
        self._capture(first=True)

    # ------------------------------------------------------------------------------------------
    def _zero_grads(self):
        for p in self.params:
            p.grad = None
        for t in self.static_inputs:
            t.grad = None

    def _eager(self):
        self._zero_grads()
        out = self.fn(*self.static_inputs)
        return out

    def _capture(self, first: bool):
        if L.TIMING is not None:
            raise RuntimeError("per-call device timing (_lib.TIMING) records events and cannot run under capture")
        # Warm-up and capture run on ONE side stream: autograd gives every leaf's AccumulateGrad node the stream it was
        # created on, and a node created on another stream would make the engine synchronise the capturing stream with
        # it, which invalidates the capture.  (The same happens when the caller still holds an autograd graph of an
        # earlier eager step over these parameters — drop such references before constructing a GraphedStep.)
        if self._stream is None:
            self._stream = torch.cuda.Stream()
        self._stream.wait_stream(torch.cuda.current_stream())
        if first:
            # eager runs settle the capacity guesses of every launch shape in the step (and the gradient-record pool)
            with torch.cuda.stream(self._stream):
                for _ in range(self.warmup):
                    self._eager()
        torch.cuda.synchronize()
        self._zero_grads()
        gc.collect()
        cap = _Capture()
        graph = torch.cuda.CUDAGraph()
        ops.CAPTURE = cap
        try:
            with torch.cuda.graph(graph, stream=self._stream):
                out = self.fn(*self.static_inputs)
        finally:
            ops.CAPTURE = None
        torch.cuda.current_stream().wait_stream(self._stream)
        self._graph, self._cap = graph, cap
        self._single = torch.is_tensor(out)
        self.static_outputs = [out] if self._single else list(out)
        self.static_grads = [p.grad for p in self.params]
        self.static_input_grads = [t.grad for t in self.static_inputs]
        self._event = torch.cuda.Event()
        self.captures += 1

    # ------------------------------------------------------------------------------------------
    def __call__(self, *inputs):
        if len(inputs) != len(self.static_inputs):
            raise ValueError(f"expected {len(self.static_inputs)} inputs, got {len(inputs)}")
        self.validate()                       # the previous replay (no-op when already validated)
        with torch.no_grad():
            for s, t in zip(self.static_inputs, inputs):
                if t is not s:
                    s.copy_(t, non_blocking=True)
        self._graph.replay()
        self._event.record()
        self._pending = True
        self.replays += 1
        for p, g in zip(self.params, self.static_grads):       # the caller may have dropped them (zero_grad)
            p.grad = g
        return self.outputs

    @property
    def outputs(self):
        return self.static_outputs[0] if self._single else tuple(self.static_outputs)

    @property
    def input_grads(self):
        return list(self.static_input_grads)

    def intersection_counts(self):
        """[(count of the last validated replay, capacity baked into the graph)] per tile binning of the step"""
        return [(int(h[0]), int(c)) for h, c, _ in self._cap.checks]

    def validate(self) -> bool:
        """Waits for the last replay and checks that every tile-list capacity baked into the graph held the step's
        intersections.  True: the replay's results stand.  False: a list overflowed — the step has been recomputed
        eagerly (outputs / gradients now hold the exact results) and the graph re-captured with the larger capacity."""
        if not self._pending:
            return True
        self._event.synchronize()
        self._pending = False
        over = [(int(h[0]), c, key) for h, c, key in self._cap.checks if int(h[0]) > c]
        if not over:
            return True
        self.overflows += 1
        for n, _c, key in over:
            ops._CAP_CACHE[key] = int(n * 1.25) + 4096
        self._stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._stream):      # (the stream of the capture: see _capture)
            out = self._eager()               # exact: the eager path redoes its own overflows
            outs = [out] if torch.is_tensor(out) else list(out)
            exact_out = [o.detach().clone() for o in outs]
            exact_g = [None if p.grad is None else p.grad.detach().clone() for p in self.params]
            exact_ig = [None if t.grad is None else t.grad.detach().clone() for t in self.static_inputs]
        del out, outs
        self._capture(first=False)
        with torch.no_grad():
            for dst, src in zip(self.static_outputs, exact_out):
                dst.copy_(src)
            for dst, src in zip(self.static_grads + self.static_input_grads, exact_g + exact_ig):
                if dst is not None and src is not None:
                    dst.copy_(src)
        torch.cuda.current_stream().synchronize()      # the side-stream temporaries are released only after their last use
        return False
