"""Re-created reference path on the GPU: the reference renderer's *call pattern* —
per sub-frame one `render()` = torch-op attribute synthesis (≈25-kernel Hermite chain, activations,
5 `torch.cat`s) + up to five independent `gsplat.rendering.rasterization()` pipelines + cuDNN-style
decoder, K sub-frames issued sequentially from Python, `mean(stack())` — running on THIS repo's
gsplat-compatible operators (mobgs_b200.rendering; gsplat itself cannot be installed here).

The orchestration is oracle/mobgs_ref.py's restatement of gaussian_renderer/__init__.py (pinned to
the reference source by tests/golden) with its operator module swapped for the CUDA one.  It
measures what fusion + K-batching buy on identical kernels; it is NOT gsplat's own kernels.

    python tests/perf/bench_unfused.py [workload] [steps]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from mobgs_b200 import rendering as R  # noqa: E402
from mobgs_b200.scene import make_camera, subframe_w2c, synthetic_scene  # noqa: E402
from mobgs_b200.subframes import render_subframes  # noqa: E402
from oracle import mobgs_ref as M  # noqa: E402

M.G = R     # the two gsplat operators now resolve to libmobgs_b200.so


def timed(fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    ns, nd, W, H, K = bench.WORKLOADS[name]
    dev = torch.device("cuda")
    stat, dyn, intr = synthetic_scene(ns, nd, W, H, seed=1234, device=dev)
    params = [p for pc in (stat, dyn) for p in pc.parameters() if p.requires_grad]
    cams = [make_camera(intr, subframe_w2c(k, K, device=dev), time=0.5) for k in range(K)]
    deltas = (torch.linspace(-1, 1, K) * 0.4).tolist() if K > 1 else [0.0]
    bg = torch.zeros(3, device=dev)
    tgt = torch.rand(3, H, W, device=dev)

    def reference_pattern(full):
        def step():
            for p in params:
                p.grad = None
            imgs = []
            for k in range(K):
                out = M.render_ref(cams[k], stat, dyn, None, bg, get_static=full, get_dynamic=full,
                                   delta_exposure=None if k == K // 2 else deltas[k])
                imgs.append(out["render"])
            loss = (M.blur_mean(imgs) - tgt).abs().mean()
            loss.backward()
        return step

    view = torch.stack([c.world_view_transform.t() for c in cams]).contiguous()
    tpoly = torch.tensor([0.5 + d / 23 for d in deltas], device=dev)
    rays = torch.cat([c.cam_ray for c in cams])

    def fused():
        for p in params:
            p.grad = None
        out = render_subframes(stat, dyn, view, cams[0].K, tpoly.clamp(0, 1), tpoly, rays, bg, W, H)
        (out["render"] - tgt).abs().mean().backward()

    from mobgs_b200.gaussian_renderer import render as dropin_render

    def dropin_loop():
        """unchanged train.py on the drop-in renderer: K sequential render() calls (train.py:441,:512)"""
        for p in params:
            p.grad = None
        imgs = []
        for k in range(K):
            out = dropin_render(cams[k], stat, dyn, None, bg, get_static=True, get_dynamic=True,
                                delta_exposure=None if k == K // 2 else deltas[k])
            imgs.append(out["render"])
        (M.blur_mean(imgs) - tgt).abs().mean().backward()

    from mobgs_b200.subframes import render_blurry_view
    expo = torch.tensor(deltas, device=dev)

    def blurry_view():
        """same outputs as the train.py pattern (blurred image + the centre's s/d renders), one launch chain"""
        for p in params:
            p.grad = None
        out = render_blurry_view(cams[K // 2], cams, expo, stat, dyn, None, bg)
        (out["render"] - tgt).abs().mean().backward()

    ms_bview = timed(blurry_view, steps)

    # the flow half of the step (train.py:563-579): K get_flow calls per view
    from mobgs_b200.gaussian_renderer import get_flow as dropin_get_flow, get_flow_batched
    half = max(K // 2, 1)
    fdeltas = [(k - K // 2) / half for k in range(K)]

    def flow_loss(outs):
        return sum(o.mean() for o in outs)

    def ref_flows():
        for p in params:
            p.grad = None
        tot = 0
        for d in fdeltas:
            tot = tot + flow_loss(M.get_flow_ref(cams[K // 2], stat, dyn, None, bg, delta_exposure=d))
        tot.backward()

    def dropin_flows():
        for p in params:
            p.grad = None
        tot = 0
        for d in fdeltas:
            tot = tot + flow_loss(dropin_get_flow(cams[K // 2], stat, dyn, None, bg, delta_exposure=d))
        tot.backward()

    def batched_flows():
        for p in params:
            p.grad = None
        flow_loss(get_flow_batched(cams[K // 2], stat, dyn, None, bg, fdeltas)).backward()

    ms_flow_ref, ms_flow_dropin, ms_flow_batched = timed(ref_flows, steps), timed(dropin_flows, steps), timed(batched_flows, steps)
    ms_dropin = timed(dropin_loop, steps)
    ms_full = timed(reference_pattern(True), steps)
    ms_min = timed(reference_pattern(False), steps)
    ms_fused = timed(fused, steps)
    print(json.dumps({
        "workload": name, "K": K,
        "reference_call_pattern_train_py_ms": ms_full,       # 5 rasterisations per sub-frame (train.py:441,:512)
        "reference_call_pattern_minimal_ms": ms_min,         # 1 rasterisation per sub-frame
        "dropin_render_loop_ms": ms_dropin,                  # mobgs_b200.gaussian_renderer.render x K, train.py unchanged
        "fused_blurry_view_ms": ms_bview,                    # render_blurry_view: K sub-frames + centre s/d lists
        "fused_k_batched_ms": ms_fused,
        "get_flow_reference_pattern_ms": ms_flow_ref,        # K x get_flow: 4 rasterisations + 2 projections each
        "get_flow_dropin_loop_ms": ms_flow_dropin,
        "get_flow_batched_ms": ms_flow_batched,
        "speedup_get_flow_batched_vs_reference_pattern": ms_flow_ref / ms_flow_batched,
        "speedup_blurry_view_vs_train_py_pattern": ms_full / ms_bview,
        "speedup_vs_train_py_pattern": ms_full / ms_fused, "speedup_vs_minimal_pattern": ms_min / ms_fused,
        "note": "same sm_100a kernels underneath both arms; the reference arm keeps the reference's per-call "
                "structure (sequential K, 5 pipelines per render, torch-op attribute synthesis)"}))


if __name__ == "__main__":
    main()
