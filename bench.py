#!/usr/bin/env python
"""bench.py — the MoBGS render + deblur hot path on B200.

One "step" = one blurry training view: K latent sub-frames rendered in one launch chain
(synth+project -> tile bin/sort -> blend -> decode+mean), L1 loss against a target image, full
backward to every Gaussian / decoder / pose gradient.  At N>1 every rank renders its own view
(data parallel over views; parameters replicated) and the flat Gaussian-gradient buffer is
all-reduced over NCCL — weak scaling.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every key.

`--impl reference` times the CPU oracle port of the reference path (oracle/) on a BOUNDED SAMPLE of
the SAME workload: all Gaussians are synthesised + projected for all K sub-frames, one 64x64-pixel
window of the frame is rasterised / decoded / averaged / differentiated, and the per-pixel cost is
scaled to the frame (oracle/bench_ref.py).  Both arms print the same `config`.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name: (n_static, n_dynamic, width, height, K, footprint)
#   footprint "px128": projected 1-sigma size 0.46-2.3 px at every resolution (SURVEY §8d's "~2-10 px radius")
#   footprint "literal": SURVEY §8d's log-scale ~ U(ln 0.004 z, ln 0.02 z) taken literally with fx = 0.9 W,
#                        i.e. the px128 footprint scaled by W / 128 (6.9-34.6 px sigma at 1080p)
WORKLOADS = {
    "c1_1k_128_K1": (700, 300, 128, 128, 1, "px128"),                 # BASELINE configs[0] (CPU-runnable)
    "c2_200k_960x540_K7": (140_000, 60_000, 960, 540, 7, "px128"),     # configs[1]
    "c3_500k_960x540_K7": (350_000, 150_000, 960, 540, 7, "px128"),    # configs[2]
    "c4_1M_1080p_K7": (700_000, 300_000, 1920, 1080, 7, "px128"),      # the metric's "1M Gaussians K=7"
    "c4L_1M_1080p_K7": (700_000, 300_000, 1920, 1080, 7, "literal"),   # same, 15x larger splats (sensitivity run)
    "c4_1M_1080p_K9": (700_000, 300_000, 1920, 1080, 9, "px128"),      # configs[3]
    "sb_150k_512x288_K9": (100_000, 50_000, 512, 288, 9, "px128"),     # the reference's real training shape (Stereo-Blur loader size, num_warp=9)
}
DEFAULT_WORKLOAD = "c4_1M_1080p_K7"
REFERENCE_CONFIG = "c1_1k_128_K1"            # BASELINE configs[0]: the one config both arms can run in full
TRAINING_SHAPE = "sb_150k_512x288_K9"        # what the reference actually trains at (host-bound without a CUDA graph)
METRIC = "rendered_Mpix_per_s_train_step"   # K*H*W / (fwd+loss+bwd time); ms_per_step = train-step ms
HBM_PEAK_FALLBACK = 6650.0                  # GB/s, B200_PROFILING.md fallback
CPU_WINDOW = 64                             # side of the window the CPU arm rasterises (tile aligned)


def footprint_px(kind, W):
    base = (0.4608, 2.304)
    return base if kind == "px128" else (base[0] * W / 128.0, base[1] * W / 128.0)


def make_config(workload, world=1, shard_sub=False):
    """The `config` object — identical for both arms of one workload (the driver compares them)."""
    ns, nd, W, H, K, fp = WORKLOADS[workload]
    lo, hi = footprint_px(fp, W)
    return {"workload": workload, "gaussians": ns + nd, "static": ns, "dynamic": nd, "width": W, "height": H,
            "subframes": K, "footprint_sigma_px": [round(lo, 3), round(hi, 3)],
            "views_per_step": 1 if shard_sub else world,
            "parallelism": f"subframes_over_{world}" if shard_sub else f"dp{world}_views",
            "l2": "flushed between timed iterations (192 MB memset, untimed)",
            "step": "K-sub-frame render + decode + blur mean + L1 + full backward"
                    + (" + NCCL all-reduce of Gaussian gradients" if world > 1 else "")}


def _peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:  # noqa: BLE001
        return HBM_PEAK_FALLBACK, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled every 10 ms from a thread
    (the timed region of a default run is a fraction of a second — `nvidia-smi -lms` delivers too few rows,
    none at all when 8 ranks query at once); nvidia-smi only if NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index, self.mark = [], None, index, 0
        self.nvml, self.handle, self.stop_flag, self.th = None, None, False, None

    def mark_start(self):
        """rows sampled from here on belong to the timed region"""
        self.mark = len(self.rows)

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:  # noqa: BLE001
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _poll(self):
        nv, h = self.nvml, self.handle
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                flags = ["Active" if r & bits[n] else "Not Active"
                         for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")]
                self.rows.append([str(self.index), str(sm), str(mx), "0"] + flags)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.01)

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:  # noqa: BLE001
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.th.join(timeout=1)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()
        sm, mx, reasons = [], None, set()
        rows = self.rows[self.mark:] or self.rows[-1:]
        for r in rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (oracle/), bounded sample of the SAME workload
# ---------------------------------------------------------------------------------------------
class CpuSample:
    """The oracle's train step (oracle/bench_ref.blurry_view_step) on the scene of `workload`.

    Frames up to 128x128 are rendered in full.  Larger frames: every Gaussian is synthesised + projected for
    all K sub-frames (as the full step would), but only one CPU_WINDOW^2 window per step is rasterised (three
    window positions, cycled); the full-frame step time is t_gaussian + t_pixel * (W H) / window_pixels."""

    def __init__(self, workload):
        import torch
        from mobgs_b200.scene import make_camera, subframe_w2c, synthetic_scene
        ns, nd, W, H, K, fp = WORKLOADS[workload]
        self.workload, self.K, self.W, self.H, self.N = workload, K, W, H, ns + nd
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.stat, self.dyn, intr = synthetic_scene(ns, nd, W, H, seed=1234, footprint_px=footprint_px(fp, W))
        self.cams = [make_camera(intr, subframe_w2c(k, K)) for k in range(K)]
        self.deltas = (torch.linspace(-1, 1, K) * 0.4).tolist() if K > 1 else [0.0]
        self.bg = torch.zeros(3)
        self.tgt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(7))
        if W * H <= 128 * 128:
            self.windows = [None]
        else:
            s = CPU_WINDOW
            pos = [(W // 2, H // 2), (W // 4, H // 4), (3 * W // 4, 3 * H // 4)]
            self.windows = [(min(x // 16 * 16, W - s) // 16 * 16, min(y // 16 * 16, H - s) // 16 * 16, s, s) for x, y in pos]
        self.n = 0
        self.tg = self.tp = self.wall = 0.0

    def step(self):
        from oracle import bench_ref
        for pc in (self.stat, self.dyn):
            for p in pc.parameters():
                p.grad = None
        t0 = time.perf_counter()
        r = bench_ref.blurry_view_step(self.stat, self.dyn, self.cams, self.deltas, self.bg, self.tgt,
                                       window=self.windows[self.n % len(self.windows)])
        self.wall += time.perf_counter() - t0
        self.n += 1
        self.tg += r["t_gaussian"]
        self.tp += r["t_pixel"]
        self.win_px, self.frame_px = r["window_pixels"], r["frame_pixels"]
        return r

    def reset(self):
        self.n, self.tg, self.tp, self.wall = 0, 0.0, 0.0, 0.0

    def result(self):
        """-> (Mpix/s of the full-frame step, full-frame step ms, measured sample step ms, description)"""
        from oracle import bench_ref
        tg, tp = self.tg / self.n, self.tp / self.n
        full = bench_ref.extrapolate_full_step(tg, tp, self.win_px, self.frame_px)
        if self.windows[0] is None:
            desc = (f"oracle port (pure PyTorch fp32) of the full step: {self.N} Gaussians, {self.W}x{self.H}, "
                    f"K={self.K}, fwd+L1+bwd ({self.n} steps, {self.wall / self.n * 1e3:.0f} ms each)")
        else:
            desc = (f"oracle port (pure PyTorch fp32), bounded sample of {self.workload}: all {self.N} Gaussians "
                    f"synthesised+projected for K={self.K} sub-frames (fwd+bwd {tg * 1e3:.0f} ms/step), one "
                    f"{CPU_WINDOW}x{CPU_WINDOW}-px window of the {self.W}x{self.H} frame rasterised+decoded+averaged, "
                    f"L1, bwd ({tp * 1e3:.0f} ms/step; 3 window positions cycled); full-frame step = t_gaussian + "
                    f"t_pixel*{self.frame_px // self.win_px} = {full:.1f} s ({self.n} steps measured, "
                    f"{self.wall / self.n * 1e3:.0f} ms each)")
        return self.K * self.frame_px / full / 1e6, full * 1e3, self.wall / self.n * 1e3, desc


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cs = CpuSample(args.workload)
    for _ in range(max(1, min(args.warmup, 2))):
        cs.step()
    cs.reset()
    for _ in range(args.steps):
        cs.step()
    val, full_ms, sample_ms, desc = cs.result()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sample_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args.workload, args.gpus),
        "full_frame_ms_per_step": full_ms,
        "note": "reference arm = CPU oracle port of the gsplat-1.4.0 + MoBGS renderer path (gsplat itself is not "
                "installable here) on a bounded sample of config.workload; `value` is the full-frame throughput the "
                "sample implies (cpu_baseline.sample), `ms_per_step` the measured sample step",
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": cs.cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
RAY_IMAGES = False      # --ray-images: materialise Camera.cam_ray [K,6,H,W] (round-1 path) instead of in-kernel rays


def build_rays(viewmats, intr, W, H):
    """Camera.cam_ray of the K sub-frame cameras (scene/cameras.py:132-146).  Default: a RayPose — 12 pose floats per
    camera from one batched inverse; the blend kernels generate every pixel's ray in registers and reduce the pose
    gradient themselves.  --ray-images: the [K,6,H,W] tensor from one
    mobgs_camera_rays_fwd launch, read back by the decoder epilogue / prologue."""
    from mobgs_b200.cameras import camera_rays_from_w2c, ray_pose_from_w2c
    # the synthetic poses are rigid transforms: closed-form inverse (torch.inverse synchronises the stream for its info check)
    if RAY_IMAGES:
        return camera_rays_from_w2c(viewmats, intr.fx, intr.fy, intr.cx, intr.cy, W, H, rigid=True)
    return ray_pose_from_w2c(viewmats, intr.fx, intr.fy, intr.cx, intr.cy, rigid=True)


class GpuJob:
    """Scene + host / device inputs of one workload on this rank, and the train step over them."""

    def __init__(self, workload, dev, rank, world, shard_sub):
        import torch
        from mobgs_b200.dist import shard_items
        from mobgs_b200.scene import subframe_w2c, synthetic_scene
        ns, nd, W, H, K, fp = WORKLOADS[workload]
        self.workload, self.dev, self.rank, self.world, self.shard_sub = workload, dev, rank, world, shard_sub
        self.W, self.H, self.K, self.N = W, H, K, ns + nd
        self.stat, self.dyn, self.intr = synthetic_scene(ns, nd, W, H, seed=1234, device=dev,
                                                         footprint_px=footprint_px(fp, W))   # same replica on every rank
        self.all_params = [p for pc in (self.stat, self.dyn) for p in pc.parameters() if p.requires_grad]
        self.my_k = list(shard_items(K, rank, world)) if shard_sub else list(range(K))
        gen = torch.Generator().manual_seed(100 + (0 if shard_sub else rank))   # views: each rank its own view
        self.tgt_host = torch.rand(3, H, W, generator=gen).pin_memory()
        base_time = 0.3 + 0.4 * float(torch.rand(1, generator=gen))
        yaw = 0.0 if shard_sub else 0.5 * (rank - (world - 1) / 2)   # degrees
        view_host = torch.stack([subframe_w2c(k, K) for k in range(K)])
        a = math.radians(yaw)
        R = torch.eye(4); R[0, 0], R[0, 2], R[2, 0], R[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
        self.view_host = (view_host @ R).contiguous().pin_memory()
        deltas = torch.linspace(-1, 1, K) * 0.4 if K > 1 else torch.zeros(1)
        self.tpoly_host = (base_time + deltas / 23).pin_memory()
        intr = self.intr
        self.Kmat = torch.tensor([[intr.fx, 0, intr.cx], [0, intr.fy, intr.cy], [0, 0, 1.0]], device=dev)
        self.bg = torch.zeros(3, device=dev)
        # device-resident copies for the kernel-side ("value") measurement
        self.tgt_d, self.view_d, self.tpoly_d = self.tgt_host.to(dev), self.view_host.to(dev), self.tpoly_host.to(dev)
        self.rays_d = build_rays(self.view_d, intr, W, H)
        # e2e input pipeline: like a data loader, the NEXT step's host buffers are copied on a side
        # stream while the current step computes; every step still issues exactly one H2D of its inputs
        # inside the timed region (h2d_bytes_per_step) and reads its loss back.
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.pending = {}
        self.stats = {}
        self.loss_ring = torch.zeros(256, dtype=torch.float32).pin_memory()
        self.loss_i = 0
        self.symmetric = None        # mobgs_b200.dist.SymmetricGradients when the NVLS gradient buffer is in use

    @property
    def h2d_bytes(self):
        return (self.tgt_host.numel() + self.view_host.numel() + self.tpoly_host.numel()) * 4

    def prefetch(self):
        import torch
        with torch.cuda.stream(self.copy_stream):
            bufs = (self.view_host.to(self.dev, non_blocking=True), self.tpoly_host.to(self.dev, non_blocking=True),
                    self.tgt_host.to(self.dev, non_blocking=True))
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.pending["next"] = (bufs, ev)

    def step(self, resident: bool):
        import torch
        from mobgs_b200.dist import FlatGradients, blur_from_partial_sums
        from mobgs_b200.losses import l1_loss
        from mobgs_b200.subframes import render_subframes
        stat, dyn, W, H, K, dev = self.stat, self.dyn, self.W, self.H, self.K, self.dev
        stats, my_k = self.stats, self.my_k
        for p in self.all_params:
            p.grad = None
        if self.symmetric is not None:
            self.symmetric.begin_step()
        if resident:
            view, tpoly, tgt, rays = self.view_d, self.tpoly_d, self.tgt_d, self.rays_d
        else:   # e2e: host buffers in, loss out
            if "next" not in self.pending:
                self.prefetch()
            (view, tpoly, tgt), ev = self.pending.pop("next")
            torch.cuda.current_stream().wait_event(ev)
            for t in (view, tpoly, tgt):
                t.record_stream(torch.cuda.current_stream())
            self.prefetch()                                  # inputs of the following step
            rays = build_rays(view, self.intr, W, H)
        view = view.requires_grad_(True) if resident else view.clone().requires_grad_(True)
        if self.shard_sub:
            # one view, its K sub-frames split across ranks: partial image sums are all-reduced in the
            # forward (collective 1 of SURVEY §8e), Gaussian gradients in the backward (collective 2)
            if my_k:
                ks = slice(my_k[0], my_k[-1] + 1)
                out = render_subframes(stat, dyn, view[ks], self.Kmat, tpoly[ks].clamp(0, 1), tpoly[ks],
                                       rays[ks] if rays.shape[0] > 1 else rays, self.bg, W, H)
                local = out["subframes"]
            else:
                out, local = None, torch.zeros(0, 3, H, W, device=dev)
            pred = blur_from_partial_sums(local, K)
        else:
            out = render_subframes(stat, dyn, view, self.Kmat, tpoly.clamp(0, 1), tpoly, rays, self.bg, W, H)
            pred = out["render"]
        loss = l1_loss(pred, tgt)        # utils/loss_utils.py:233 (train.py:621) as one fused launch
        if loss.requires_grad:
            loss.backward()
        if self.world > 1:
            if "fg" not in stats:      # parameter set that receives gradients (fixed across steps)
                stats["fg_params"] = [p for p in self.all_params if p.grad is not None or self.shard_sub]
                if self.shard_sub:          # idle / static-only ranks still contribute zeros
                    stats["fg_params"] = [p for p in stats["fg_params"] if p is not stat.control_xyz
                                          and p is not stat._omega and p is not stat._features_t
                                          and p is not stat._trbf_center and p is not dyn._trbf_center
                                          and p is not dyn._xyz and all(p is not q for q in stat.rgbdecoder.parameters())]
                stats["fg"] = FlatGradients(stats["fg_params"], inplace_shared=not self.shard_sub)
            stats["fg"].reduce()
            stats["allreduce_bytes"] = stats["fg"].last_collective_elems * 4
        view.grad = None
        stats["out"] = out
        if not resident:
            # D2H read of the step's result: an asynchronous copy into a pinned ring inside the step (complete when the
            # step's closing event is), consumed by the host afterwards (`losses()`), so the host keeps enqueueing
            slot = self.loss_ring[self.loss_i % self.loss_ring.shape[0]]
            slot.copy_(loss.detach(), non_blocking=True)
            self.loss_i += 1
            return slot
        return loss

    def losses(self):
        """host view of the e2e steps' results (after a synchronize)"""
        n = min(self.loss_i, self.loss_ring.shape[0])
        return [float(v) for v in self.loss_ring[:n]]


def timed(step_fn, flush, steps, warmup, dev_index, world, sample_clocks=False, reset_lib_timing=False, host_sync=False):
    """W untimed warm-up calls, then `steps` calls timed one by one with CUDA events (L2 flushed, untimed,
    between them), barrier + synchronize on both sides, max over ranks.  -> (ms per step, clocks)
    Default: the host only ENQUEUES (as a training loop does): every step is bracketed by its own event pair and the
    host synchronises once after the K steps, so the Python prologue of step i+1 overlaps the GPU work of step i.
    host_sync=True: the host waits for every step before starting the next (round 1's loop; exposes ~0.1-0.3 ms of
    host prologue per step as GPU idle time)."""
    import torch
    import torch.distributed as dist
    from mobgs_b200 import _lib
    sampler = ClockSampler(dev_index) if sample_clocks else None
    if sampler:
        sampler.start()            # nvidia-smi needs a few 100 ms to deliver its first row
    for _ in range(warmup):
        step_fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.mark_start()
    if reset_lib_timing:
        _lib.TIMING = {}
        _lib.LAUNCH_COUNT = 0
    pairs = []
    for _ in range(steps):
        flush.zero_()                      # evict L2 between timed iterations (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_fn()
        e1.record()
        if host_sync:
            e1.synchronize()
        pairs.append((e0, e1))
    torch.cuda.synchronize()
    tot = sum(a.elapsed_time(b) for a, b in pairs)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        dist.barrier()
        t = torch.tensor([tot], device=torch.device("cuda", dev_index))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot = float(t)
    return tot / steps, clocks


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mobgs_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    shard_sub = args.shard == "subframes" and world > 1
    # gradient all-reduce overlapped with the second half of the projection backward: on by default at N > 1 (measured
    # at N = 2: 5.879 -> 5.815 ms WITH high-priority collective streams, 5.926 ms without them — the collective's CTAs
    # must win SM slots against the queued projection CTAs)
    overlap = world > 1 and not shard_sub and (args.overlap or not args.no_overlap)
    if overlap:
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")      # read when the process group creates its streams
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    symmetric = None
    # measured (profiles/r2_allreduce_n*.json): 116 MB multimem all-reduce 0.285 ms vs NCCL 0.36 ms on 8 GPUs, but 1.03 ms vs
    # 0.245 ms on 2 — the in-switch reduction only pays with many peers, so it is used from 8 ranks up
    if world >= 8 and not shard_sub and not args.no_symmetric:
        # flat gradient buffer in NVLS symmetric memory, reduced by the multimem (in-switch) all-reduce
        from mobgs_b200.dist import SymmetricGradients
        symmetric = SymmetricGradients()
        try:
            if not symmetric.install():
                symmetric = None
        except Exception as e:  # noqa: BLE001
            print(f"[bench] symmetric-memory gradients unavailable ({e}); using NCCL", file=sys.stderr)
            symmetric = None
    if overlap:     # chunked projection backward, each chunk's gradient all-reduce on a side stream (mobgs_b200.fused.GradSink)
        from mobgs_b200.dist import overlap_gradient_allreduce
        overlap_gradient_allreduce(True, n_chunks=args.overlap_chunks)
    job = GpuJob(args.workload, dev, rank, world, shard_sub)
    job.symmetric = symmetric
    W, H, K, N = job.W, job.H, job.K, job.N
    flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)   # > 126 MB L2

    # ---- kernel-side throughput: inputs resident in HBM ----
    _lib.TIMING = {}
    _lib.LAUNCH_COUNT = 0
    for _ in range(args.warmup):
        job.step(True)
    torch.cuda.synchronize()
    ms, clocks = timed(lambda: job.step(True), flush, args.steps, 3, local, world, sample_clocks=True,
                       reset_lib_timing=True)
    launches = _lib.LAUNCH_COUNT
    kernel_ms = {n: sum(a.elapsed_time(b) for a, b in ev) / args.steps for n, ev in _lib.TIMING.items()}
    _lib.TIMING = None

    # ---- algorithmic bytes of the dominant kernels (I_eff measured from the forward outputs) ----
    my_k = job.my_k
    ks_loc = slice(my_k[0], my_k[-1] + 1) if my_k else slice(0, 1)
    ieff, itot = measure_intersections(job.stat, job.dyn, job.view_d[ks_loc], job.Kmat, job.tpoly_d[ks_loc], W, H)
    P = K * W * H
    Pjob = P if shard_sub else world * P       # pixels rendered by the whole job per step
    P_loc = len(my_k) * W * H                  # pixels this rank's kernels render per step
    bytes_fwd = 68.0 * ieff + 48.0 * P_loc
    bytes_bwd = 132.0 * ieff + 52.0 * P_loc
    peak, peak_kind = _peak()
    dom = max(("mobgs_blend_fwd", "mobgs_blend_bwd"), key=lambda n: kernel_ms.get(n, 0.0))
    dom_bytes = bytes_bwd if dom == "mobgs_blend_bwd" else bytes_fwd
    ach = dom_bytes / (kernel_ms[dom] * 1e-3) / 1e9
    traffic = traffic_src = None
    pipes = {}
    try:    # dram bytes of one launch from the committed `ncu --set full` capture of this workload
        for name in ("r2c_traffic.json", "r2b_traffic.json", "r2_traffic.json", "r1_traffic.json"):      # newest capture first
            pth = os.path.join(ROOT, "profiles", name)
            if os.path.exists(pth):
                with open(pth) as f:
                    tj = json.load(f)
                traffic = tj.get(args.workload, {}).get(dom)
                if traffic is not None:
                    traffic_src = "profiles/" + name
                    pipes = {"issue_slot_utilisation": tj.get("issue_slot_utilisation", {}).get(dom),
                             "lsu_pipe_utilisation": tj.get("lsu_pipe_utilisation", {}).get(dom)}
                    break
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "peak_kind": peak_kind,
                "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
                "ms_per_launch": kernel_ms[dom], "algorithmic_bytes": dom_bytes,
                "intersections_consumed": ieff, "intersections_listed": itot,
                "intersections_per_gaussian_subframe": itot / max(1, len(my_k) * N), "pixels": P_loc,
                # what actually bounds the kernel (same ncu capture as `traffic`): fraction of the issue slots / of the LSU
                # (shared-memory wavefront) pipe in use
                **{k: v for k, v in pipes.items() if v is not None},
                "note": "the blend kernels are bound by instruction issue and the shared-memory pipe, not by HBM "
                        "(ncu summaries under profiles/); the HBM fraction is reported because BASELINE.json's "
                        "north_star asks for it. algorithmic_bytes = 132*I_eff + 52*P (bwd) / 68*I_eff + 48*P (fwd)"}

    # ---- end to end through the public API with host buffers ----
    e2e_ms, _ = timed(lambda: job.step(False), flush, args.steps, max(1, args.warmup // 2), local, world)
    e2e_losses = job.losses()
    # the same two loops with the host waiting for every step (round 1's measurement loop), for comparison
    hs_steps = max(3, min(args.steps, 10))
    hs_ms, _ = timed(lambda: job.step(True), flush, hs_steps, 1, local, world, host_sync=True)
    hs_e2e_ms, _ = timed(lambda: float(job.step(False)), flush, hs_steps, 1, local, world, host_sync=True)

    extra = {}
    if world == 1 and not args.no_extras:
        # the same step on BASELINE configs[0] — the one config the CPU port runs in full — on GPU and CPU
        j1 = GpuJob(REFERENCE_CONFIG, dev, 0, 1, False)
        g1_ms, _ = timed(lambda: j1.step(True), flush, 20, 5, local, 1)
        g1_e2e, _ = timed(lambda: j1.step(False), flush, 20, 3, local, 1)
        c1 = CpuSample(REFERENCE_CONFIG)
        c1.step(); c1.reset()
        t0 = time.perf_counter()
        while c1.n < 3 or time.perf_counter() - t0 < 3.0:
            c1.step()
        c1_val, c1_ms, _, c1_desc = c1.result()
        px1 = j1.K * j1.W * j1.H
        extra["gpu_on_reference_config"] = {
            "workload": REFERENCE_CONFIG, "same_config": True, "gpu_ms_per_step": g1_ms, "gpu_e2e_ms_per_step": g1_e2e,
            "gpu_Mpix_per_s": px1 / (g1_ms * 1e-3) / 1e6, "cpu_ms_per_step": c1_ms, "cpu_Mpix_per_s": c1_val,
            "cpu_cores": c1.cores, "cpu_sample": c1_desc, "ratio": c1_ms / g1_ms, "e2e_ratio": c1_ms / g1_e2e,
            "note": "1 k Gaussians at 128x128 is launch-latency bound on the GPU (~20 launches); the ratio says how "
                    "the two arms compare on the only config the CPU port runs in full, nothing about kernel quality"}
        del j1, c1
        try:
            extra["full_step"] = measure_full_step(args, job, flush, local)
        except Exception as e:  # noqa: BLE001  (the headline must survive a failure of the extra measurement)
            extra["full_step"] = {"error": f"{type(e).__name__}: {e}"}
        try:    # host path: the same step as one CUDA graph, here and at the reference's real (host-bound) training shape
            _lib.TIMING = None
            extra["cuda_graph"] = measure_graphed(args, job, flush, local, ms, e2e_ms)
            torch.cuda.empty_cache()
            jsb = GpuJob(TRAINING_SHAPE, dev, 0, 1, False)
            sb_ms, _ = timed(lambda: jsb.step(True), flush, 30, 5, local, 1)
            sb_e2e, _ = timed(lambda: jsb.step(False), flush, 30, 5, local, 1)
            extra["cuda_graph"][TRAINING_SHAPE] = measure_graphed(args, jsb, flush, local, sb_ms, sb_e2e, steps=30)
            del jsb
        except Exception as e:  # noqa: BLE001
            extra.setdefault("cuda_graph", {})["error"] = f"{type(e).__name__}: {e}"

    if overlap:
        from mobgs_b200.dist import overlap_gradient_allreduce
        overlap_gradient_allreduce(False)
    if symmetric is not None:
        symmetric.uninstall()
    if not args.no_extras:
        try:    # BASELINE configs[3]: batch 2 x K = 9 sub-frames of the 1 M / 1080p scene, the 18 (view, sub-frame) items split over the ranks
            extra["strong_scaling"] = measure_strong_scaling(args, dev, rank, world, flush, local)
        except Exception as e:  # noqa: BLE001
            extra["strong_scaling"] = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cs = CpuSample(args.workload)
            cs.step(); cs.reset()
            t0 = time.perf_counter()
            while cs.n < 2 or time.perf_counter() - t0 < 12.0:
                cs.step()
            cval, _, _, cdesc = cs.result()
            cpu = {"value": cval, "unit": "Mpix/s", "cores": cs.cores, "kind": "port", "sample": cdesc}
        line = {
            "metric": METRIC, "value": Pjob / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if shard_sub else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(args.workload, world, shard_sub),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": Pjob / (e2e_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": job.h2d_bytes, "d2h_bytes_per_step": 4,
                    "last_loss_on_host": e2e_losses[-1] if e2e_losses else None,
                    "pipeline": "inputs: pinned host -> device on a side stream one step ahead; result: async D2H of the "
                                "loss into a pinned ring inside the step; the host enqueues ahead and synchronises after "
                                "the K timed steps"},
            "host_sync_per_step": {"ms_per_step": hs_ms, "e2e_ms_per_step": hs_e2e_ms, "steps": hs_steps,
                                   "note": "same steps with the host waiting for each one (float(loss) / event "
                                           "synchronize) before enqueueing the next"},
            "gpu_launches": launches, "kernel_ms_per_step": kernel_ms, "clocks": clocks,
        }
        line.update(extra)
        if world > 1:
            line["allreduce_bytes_per_step"] = job.stats.get("allreduce_bytes")
            line["gradient_allreduce"] = (f"overlapped: projection backward in {args.overlap_chunks} Gaussian ranges, each "
                                          "range's slice of the flat gradient buffer all-reduced on a high-priority side "
                                          "stream while the next range's kernel runs ("
                                          + ("multimem / NVLS symmetric memory" if symmetric is not None else "NCCL") + ")"
                                          if overlap else
                                          ("one in-place multimem (NVLS symmetric-memory) all-reduce of the flat gradient "
                                           "buffer after the backward" if symmetric is not None else
                                           "one in-place NCCL all-reduce of the flat gradient buffer after the backward"))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_graphed(args, job, flush, local, eager_ms, eager_e2e_ms, steps=None):
    """The headline step (K-sub-frame render + decode + blur mean + L1 + full backward) captured into ONE CUDA graph
    (mobgs_b200.graphs.GraphedStep) and replayed: resident inputs, and end to end (pinned host inputs copied into the
    graph's static buffers, loss copied back to the pinned ring, inside the timed step).  Every replay is validated
    (tile-list capacity) before the next one is enqueued, so the host waits once per step."""
    import torch
    from mobgs_b200.graphs import GraphedStep
    from mobgs_b200.losses import l1_loss
    from mobgs_b200.subframes import render_subframes
    steps = steps or args.steps
    W, H = job.W, job.H

    def fn(view, tpoly, tgt):
        rays = build_rays(view.detach(), job.intr, W, H)
        out = render_subframes(job.stat, job.dyn, view, job.Kmat, tpoly.clamp(0, 1), tpoly, rays, job.bg, W, H)
        loss = l1_loss(out["render"], tgt)
        loss.backward()
        return loss.detach()

    job.stats.pop("out", None)          # (the last eager step's autograd graph: its AccumulateGrad nodes would break the capture)
    for p in job.all_params:
        p.grad = None
    view = job.view_d.clone().requires_grad_(True)
    step = GraphedStep(fn, [view, job.tpoly_d, job.tgt_d], job.all_params)
    ins_res = step.static_inputs
    res_ms, _ = timed(lambda: step(*ins_res), flush, steps, 3, local, 1)
    ring = job.loss_ring

    def e2e():
        # as job.step(False): this step's pinned host inputs were copied on the side stream while the previous step ran
        # (one H2D of the step's inputs per step); they reach the graph's static buffers by device copies
        if "next" not in job.pending:
            job.prefetch()
        (v, tp, tg), ev = job.pending.pop("next")
        torch.cuda.current_stream().wait_event(ev)
        for t in (v, tp, tg):
            t.record_stream(torch.cuda.current_stream())
        loss = step(v, tp, tg)
        job.prefetch()
        ring[0:1].copy_(loss.reshape(1), non_blocking=True)
    e2e_ms, _ = timed(e2e, flush, steps, 3, local, 1)
    step.validate()
    torch.cuda.synchronize()
    n, cap = step.intersection_counts()[0]
    out = {"workload": job.workload, "ms_per_step": res_ms, "e2e_ms_per_step": e2e_ms, "eager_ms_per_step": eager_ms,
           "eager_e2e_ms_per_step": eager_e2e_ms, "replays": step.replays, "captures": step.captures,
           "overflows": step.overflows, "intersections": n, "list_capacity": cap, "loss_on_host": float(ring[0]),
           "note": "one cudaGraphLaunch per step; the host validates each replay (intersection count vs the capacity baked "
                   "into the graph) before enqueueing the next"}
    del step
    return out


def measure_strong_scaling(args, dev, rank, world, flush, local, workload="c4_1M_1080p_K9", views=2):
    """One optimiser step's renders at the reference's batch size (arguments/stereo/default.py:20: 2 views) and
    num_warp (arguments/__init__.py:214: K = 9) on the 1 M / 1080p scene — BASELINE configs[3] — with the
    views x K (view, sub-frame) work items split over the ranks in contiguous blocks (mobgs_b200.dist.shard_subframes).
    TOTAL work is fixed: collective 1 = one all-reduce of the per-view partial image sums [views,3,H,W] in the
    forward, collective 2 = the Gaussian-gradient all-reduce.  The same code runs at N = 1 (no collectives): the
    driver's per-N lines give the strong-scaling curve."""
    import torch
    import torch.distributed as dist
    from mobgs_b200.dist import FlatGradients, all_reduce_sum, shard_subframes
    from mobgs_b200.losses import l1_loss
    from mobgs_b200.subframes import render_subframes
    job = GpuJob(workload, dev, 0, 1, False)            # same scene / inputs on every rank
    W, H, K = job.W, job.H, job.K
    mine = shard_subframes(views, K, rank, world)
    gen = torch.Generator().manual_seed(77)
    tgts = torch.rand(views, 3, H, W, generator=gen).to(dev)
    view_v, tpoly_v, rays_v = [], [], []
    for v in range(views):
        a = math.radians(0.8 * (v - (views - 1) / 2))
        R = torch.eye(4); R[0, 0], R[0, 2], R[2, 0], R[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
        vm = (job.view_host @ R).to(dev)
        view_v.append(vm); tpoly_v.append(job.tpoly_d + 0.05 * v); rays_v.append(build_rays(vm, job.intr, W, H))
    fg = {}

    def step():
        for p in job.all_params:
            p.grad = None
        partial = []
        for v in range(views):
            ks = [r for (vv, r) in mine if vv == v]
            if ks:
                sl = slice(ks[0].start, ks[0].stop)
                out = render_subframes(job.stat, job.dyn, view_v[v][sl], job.Kmat, tpoly_v[v][sl].clamp(0, 1), tpoly_v[v][sl],
                                       rays_v[v][sl], job.bg, W, H)
                partial.append(out["subframes"].sum(dim=0))
            else:
                partial.append(torch.zeros(3, H, W, device=dev))
        pred = all_reduce_sum(torch.stack(partial)) / K + 1e-10          # train.py:540-541 for every view at once
        loss = l1_loss(pred, tgts)
        if loss.requires_grad:
            loss.backward()
        if world > 1:
            if "fg" not in fg:
                ps = [p for p in job.all_params if p is not job.stat.control_xyz and p is not job.stat._omega
                      and p is not job.stat._features_t and p is not job.stat._trbf_center and p is not job.dyn._trbf_center
                      and p is not job.dyn._xyz and all(p is not q for q in job.stat.rgbdecoder.parameters())]
                fg["fg"] = FlatGradients(ps, inplace_shared=False)
            fg["fg"].reduce()
        return loss

    ms, _ = timed(step, flush, max(3, min(args.steps, 10)), 3, local, world)
    return {"workload": workload, "views": views, "subframes": K, "items": views * K, "n_gpus": world,
            "items_on_rank0": sum(len(r) for _, r in shard_subframes(views, K, 0, world)),
            "ms_per_step": ms, "scaling": "strong",
            "Mpix_per_s": views * K * W * H / (ms * 1e-3) / 1e6,
            "collectives": "all-reduce of [views,3,H,W] partial image sums (forward) + Gaussian-gradient all-reduce"}


def measure_intersections(stat, dyn, view, Kmat, tpoly, W, H):
    """I (listed) and I_eff (list entries up to the last one any pixel of the tile blended)."""
    import torch
    from mobgs_b200 import fused
    from mobgs_b200.gaussian_renderer import _dynamic_params, _static_params
    from mobgs_b200.ops import build_tile_lists
    with torch.no_grad():
        K = view.shape[0]
        rec, radii, depths, _ = fused.synth_project(_static_params(stat), _dynamic_params(dyn),
                                                    dyn.current_control_num, view, Kmat[None].expand(K, -1, -1),
                                                    tpoly.clamp(0, 1), tpoly, W, H)
        lists = build_tile_lists(rec, radii, depths, W, H, True)
        _, _, last = fused._BlendRecords.apply(rec, radii, depths, None, None, 10, W, H, None, True, 0)
        tx, ty = math.ceil(W / 16), math.ceil(H / 16)
        pad = torch.full((K, ty * 16, tx * 16), -1, dtype=torch.int32, device=rec.device)
        pad[:, :H, :W] = last
        tile_last = pad.reshape(K, ty, 16, tx, 16).permute(0, 1, 3, 2, 4).reshape(K * ty * tx, 256).max(dim=1).values
        beg = lists.tile_offsets[:-1]
        ieff = torch.clamp(tile_last - beg + 1, min=0).sum().item()
    return float(ieff), float(lists.n_isect)


# ---------------------------------------------------------------------------------------------
# The optimiser step MoBGS actually runs (train.py:430-680, 796-800), composed from the fused pieces
# ---------------------------------------------------------------------------------------------
def measure_full_step(args, job, flush, local, views=2):
    """2 views x (render_blurry_view + get_flow_batched) + photo loss + backward + reg / flow-warp losses +
    backward + fused Adam, in the order train.py composes them (photo_loss.backward(retain_graph=True) at
    :629, loss.backward() at :680, optimizer steps at :796-800).  Inputs resident; CUDA events."""
    import torch
    from mobgs_b200 import _lib
    from mobgs_b200.cameras import camera_rays_from_w2c, ray_pose_from_w2c
    from mobgs_b200.gaussian_renderer import get_flow_batched
    from mobgs_b200.losses import flow_warp_loss, photo_loss, reg_loss
    from mobgs_b200.optim import FusedAdam, fused_step
    from mobgs_b200.subframes import render_blurry_view
    from mobgs_b200.scene import PinholeCamera
    stat, dyn, W, H, K, dev, intr = job.stat, job.dyn, job.W, job.H, job.K, job.dev, job.intr
    half = K // 2
    lam_dssim, lam_flow = 0.2, 1e-2            # arguments/__init__.py:142, :185
    gen = torch.Generator().manual_seed(5)
    yaw = [math.radians(0.7 * (v - (views - 1) / 2)) for v in range(views)]
    view_d = []
    for a in yaw:
        R = torch.eye(4); R[0, 0], R[0, 2], R[2, 0], R[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
        view_d.append((job.view_host @ R).to(dev))
    gt = torch.rand(views, 3, H, W, generator=gen).to(dev)
    gt_depth = (2 + 8 * torch.rand(views, 1, H, W, generator=gen)).to(dev)
    times = [0.35, 0.6][:views] + [0.5] * max(0, views - 2)
    exposure_time = (torch.linspace(-1, 1, K) * 0.4).to(dev)
    deltas = torch.tensor([(k - half) / max(half, 1) for k in range(K)], device=dev)

    def pixel_grid(w, h, use_center=None):
        return PinholeCamera.get_pixels(None, w, h, use_center)

    def groups(pc, names):
        return [{"params": [getattr(pc, n)], "lr": 1.6e-4, "name": n} for n in names]
    opt_s = FusedAdam(groups(stat, ("_xyz", "_features_dc", "_opacity", "_scaling", "_rotation")), lr=0.0, eps=1e-15)
    opt_d = FusedAdam(groups(dyn, ("control_xyz", "_features_dc", "_features_t", "_opacity", "_scaling", "_rotation",
                                   "_omega")) + [{"params": list(dyn.rgbdecoder.parameters()), "lr": 1e-4,
                                                  "name": "decoder"}], lr=0.0, eps=1e-15)
    saved = [p.detach().clone() for p in job.all_params]
    parts = {}

    def cams_of(v, w2c):
        """K camera stand-ins of view v + their rays: a pose-differentiable RayPose (12 floats per camera, rays generated
        inside the blend kernels) or, with --ray-images, the stacked [K,6,H,W] tensor of one camera_rays launch"""
        # the centre sub-frame is the dataset camera itself: numpy pose, not learnable (train.py:441, scene/cameras.py:121-124)
        centre = torch.arange(K, device=dev)[:, None, None] == half
        w2c_r = torch.where(centre, w2c.detach(), w2c)
        if RAY_IMAGES:
            rays = camera_rays_from_w2c(w2c_r, intr.fx, intr.fy, intr.cx, intr.cy, W, H, rigid=True)
            per_cam = rays.split(1)
        else:
            rays = ray_pose_from_w2c(w2c_r, intr.fx, intr.fy, intr.cx, intr.cy, rigid=True)
            per_cam = [rays[k] for k in range(K)]
        cams = [SimpleNamespace(world_view_transform=w2c_r[k].transpose(0, 1), K=job.Kmat, time=times[v], max_time=23,
                                image_width=W, image_height=H, cam_ray=per_cam[k], get_pixels=pixel_grid)
                for k in range(K)]
        return cams, rays

    def step():
        for p in job.all_params:
            p.grad = None
        preds, depths, d_alphas, oris, lat_img, lat_alpha, e2m, m2e, vsps = [], [], [], [], [], [], [], [], []
        for v in range(views):
            w2c = view_d[v].clone().requires_grad_(True)
            cams, rays = cams_of(v, w2c)
            pkg = render_blurry_view(cams[half], cams, exposure_time, stat, dyn, None, job.bg, rays=rays)
            preds.append(pkg["render"]); depths.append(pkg["depth"]); d_alphas.append(pkg["d_alpha"])
            oris.append(pkg["render_center"]); vsps.append(pkg["viewspace_points"])
            a, b, li, la = get_flow_batched(cams[half], stat, dyn, None, job.bg, deltas, rays=cams[half].cam_ray)
            e2m.append(a); m2e.append(b); lat_img.append(li); lat_alpha.append(la)
        image = torch.stack(preds)
        photo = photo_loss(image, gt, lam_dssim)
        photo.backward(retain_graph=True)                                               # train.py:629
        _ = [t.grad for t in vsps]                                                      # densification statistics, :634-648
        reg, _sums = reg_loss(torch.stack(depths), gt_depth, torch.stack(d_alphas), 0.2, 1e-7)
        flow = flow_warp_loss(torch.stack(oris), torch.stack(lat_img), torch.stack(e2m), torch.stack(m2e),
                              torch.stack(lat_alpha)[:, :, None], torch.stack(d_alphas))
        (reg + lam_flow * flow).backward()                                              # train.py:680
        fused_step([opt_s, opt_d])                                                      # train.py:796-800
        return photo

    _lib.TIMING = None
    ms, _ = timed(step, flush, max(3, min(args.steps, 10)), 3, local, 1)
    # component split (separately timed, same inputs): the two render families alone, forward + backward
    def only_blurry():
        for p in job.all_params:
            p.grad = None
        for v in range(views):
            cams, rays = cams_of(v, view_d[v].clone().requires_grad_(True))
            pkg = render_blurry_view(cams[half], cams, exposure_time, stat, dyn, None, job.bg, rays=rays)
            (pkg["render"].mean() + pkg["depth"].mean() + pkg["d_alpha"].mean()).backward()

    def only_flow():
        for p in job.all_params:
            p.grad = None
        for v in range(views):
            cams, _ = cams_of(v, view_d[v])
            a, b, li, la = get_flow_batched(cams[half], stat, dyn, None, job.bg, deltas, rays=cams[half].cam_ray)
            (a.mean() + b.mean() + li.mean() + la.mean()).backward()
    parts["render_blurry_view_fwd_bwd_ms"], _ = timed(only_blurry, flush, 5, 2, local, 1)
    parts["get_flow_batched_fwd_bwd_ms"], _ = timed(only_flow, flush, 5, 2, local, 1)
    # per-entry-point device time (CUDA events around every C-ABI call) of one full step / of the two render families
    for name, fn in (("full_step", step), ("render_blurry_view", only_blurry), ("get_flow_batched", only_flow)):
        torch.cuda.synchronize()
        _lib.TIMING = {}
        fn()
        torch.cuda.synchronize()
        parts[name + "_kernel_ms"] = {n: round(sum(a.elapsed_time(b) for a, b in ev), 4) for n, ev in _lib.TIMING.items()}
        parts[name + "_kernel_calls"] = {n: len(ev) for n, ev in _lib.TIMING.items()}
        _lib.TIMING = None
    with torch.no_grad():                       # the optimiser moved the scene: restore it for whatever follows
        for p, s in zip(job.all_params, saved):
            p.copy_(s)
    return {"full_step_ms": ms, "views": views, "subframes": K,
            "composition": "per view: render_blurry_view (K sub-frames + centre dynamic-only / static-only lists) + "
                           "get_flow_batched (K exposures); photo_loss (L1 + 0.2 DSSIM) backward; reg_loss + 1e-2 * "
                           "flow_warp_loss backward; fused Adam over both models (train.py:430-680, 796-800)",
            **parts}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ray-images", action="store_true",
                    help="materialise Camera.cam_ray [K,6,H,W] (mobgs_camera_rays_fwd/bwd) instead of generating rays inside "
                         "the blend kernels from 12 pose floats per camera (the default)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra keys of the N=1 line (gpu_on_reference_config, full_step)")
    ap.add_argument("--overlap", action="store_true",
                    help="N>1: projection backward in Gaussian ranges with each range's gradient all-reduce on a "
                         "high-priority side stream (the default at N > 1)")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: one all-reduce after the whole backward")
    ap.add_argument("--no-symmetric", action="store_true", help="N>1: NCCL all-reduce instead of the NVLS multimem one")
    ap.add_argument("--overlap-chunks", type=int, default=2)
    ap.add_argument("--shard", default="views", choices=["views", "subframes"],
                    help="N>1: 'views' = one view per rank (weak scaling, default); 'subframes' = the K "
                         "sub-frames of one view split across ranks (strong scaling, BASELINE configs[3])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        global RAY_IMAGES
        RAY_IMAGES = bool(args.ray_images)
        run_ours(args)


if __name__ == "__main__":
    main()
