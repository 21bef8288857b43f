#!/usr/bin/env python
"""bench.py — the MoBGS render + deblur hot path on B200.

One "step" = one blurry training view: K latent sub-frames rendered in one launch chain
(synth+project -> tile bin/sort -> blend -> decode+mean), L1 loss against a target image, full
backward to every Gaussian / decoder / pose gradient.  At N>1 every rank renders its own view
(data parallel over views; parameters replicated) and the flat Gaussian-gradient buffer is
all-reduced over NCCL — weak scaling.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every key.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_static, n_dynamic, width, height, K)
    "c1_1k_128_K1": (700, 300, 128, 128, 1),                 # BASELINE configs[0] (CPU-runnable)
    "c2_200k_960x540_K7": (140_000, 60_000, 960, 540, 7),     # configs[1]
    "c3_500k_960x540_K7": (350_000, 150_000, 960, 540, 7),    # configs[2]
    "c4_1M_1080p_K7": (700_000, 300_000, 1920, 1080, 7),      # the metric's "1M Gaussians K=7"
    "c4_1M_1080p_K9": (700_000, 300_000, 1920, 1080, 9),      # configs[3]
    "sb_150k_512x288_K9": (100_000, 50_000, 512, 288, 9),     # the reference's real training shape (Stereo-Blur loader size, num_warp=9)
}
DEFAULT_WORKLOAD = "c4_1M_1080p_K7"
METRIC = "rendered_Mpix_per_s_train_step"   # K*H*W / (fwd+loss+bwd time); ms_per_step = train-step ms
HBM_PEAK_FALLBACK = 6650.0                  # GB/s, B200_PROFILING.md fallback


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
# CPU arm: the oracle port of the reference path (oracle/), bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_reference_step_factory(sample):
    """Returns (step_fn, pixels_per_step, description).  One step = K sub-frame oracle renders of
    the sample scene + blur mean + L1 + backward, on all host cores."""
    import torch
    from oracle import mobgs_ref as M
    from mobgs_b200.scene import make_camera, subframe_w2c, synthetic_scene
    ns, nd, W, H, K = sample
    torch.set_num_threads(os.cpu_count() or 1)
    stat, dyn, intr = synthetic_scene(ns, nd, W, H, seed=1234)
    cams = [make_camera(intr, subframe_w2c(k, K)) for k in range(K)]
    deltas = (torch.linspace(-1, 1, K) * 0.4).tolist() if K > 1 else [0.0]
    bg = torch.zeros(3)
    tgt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(7))

    def step():
        for pc in (stat, dyn):
            for p in pc.parameters():
                p.grad = None
        imgs = [M.render_ref(cams[k], stat, dyn, None, bg, delta_exposure=deltas[k])["render"] for k in range(K)]
        loss = (M.blur_mean(imgs) - tgt).abs().mean()
        loss.backward()
        return float(loss.detach())

    desc = f"oracle port (pure PyTorch fp32), {ns + nd} Gaussians, {W}x{H}, K={K}, fwd+L1+bwd"
    return step, K * W * H, desc


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = WORKLOADS["c1_1k_128_K1"]
    step, pix, desc = cpu_reference_step_factory(sample)
    for _ in range(max(1, min(args.warmup, 3))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = pix / dt / 1e6
    cores = os.cpu_count() or 1
    ns, nd, W, H, K = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "gaussians": ns + nd, "width": W, "height": H, "subframes": K,
                   "note": "reference arm = CPU oracle port of the gsplat-1.4.0 + MoBGS renderer path on a "
                           "bounded sample (gsplat itself is not installable here); throughput in Mpix/s "
                           "is size-normalised"},
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def build_rays(viewmats, intr, W, H):
    """Camera.cam_ray for the K sub-frame cameras on the device (scene/cameras.py:132-146): [K,6,H,W],
    one mobgs_camera_rays_fwd launch instead of K meshgrid + matmul chains."""
    from mobgs_b200.cameras import camera_rays_from_w2c
    return camera_rays_from_w2c(viewmats, intr.fx, intr.fy, intr.cx, intr.cy, W, H)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mobgs_b200 import _lib
    from mobgs_b200.scene import subframe_w2c, synthetic_scene
    from mobgs_b200.subframes import render_subframes
    from mobgs_b200.dist import FlatGradients, blur_from_partial_sums, shard_items

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    ns, nd, W, H, K = WORKLOADS[args.workload]
    N = ns + nd
    stat, dyn, intr = synthetic_scene(ns, nd, W, H, seed=1234, device=dev)   # same replica on every rank
    all_params = [p for pc in (stat, dyn) for p in pc.parameters() if p.requires_grad]
    shard_sub = args.shard == "subframes" and world > 1
    my_k = list(shard_items(K, rank, world)) if shard_sub else list(range(K))
    gen = torch.Generator().manual_seed(100 + (0 if shard_sub else rank))   # views: each rank its own view
    tgt_host = torch.rand(3, H, W, generator=gen).pin_memory()
    base_time = 0.3 + 0.4 * float(torch.rand(1, generator=gen))
    yaw = 0.0 if shard_sub else 0.5 * (rank - (world - 1) / 2)   # degrees
    view_host = torch.stack([subframe_w2c(k, K) for k in range(K)])
    a = math.radians(yaw)
    R = torch.eye(4); R[0, 0], R[0, 2], R[2, 0], R[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
    view_host = (view_host @ R).contiguous().pin_memory()
    deltas = torch.linspace(-1, 1, K) * 0.4 if K > 1 else torch.zeros(1)
    tpoly_host = (base_time + deltas / 23).pin_memory()
    Kmat = torch.tensor([[intr.fx, 0, intr.cx], [0, intr.fy, intr.cy], [0, 0, 1.0]], device=dev)
    bg = torch.zeros(3, device=dev)

    # device-resident copies for the kernel-side ("value") measurement
    tgt_d, view_d, tpoly_d = tgt_host.to(dev), view_host.to(dev), tpoly_host.to(dev)
    rays_d = build_rays(view_d, intr, W, H)
    flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)   # > 126 MB L2

    stats = {}

    # e2e input pipeline: like a data loader, the NEXT step's host buffers are copied on a side
    # stream while the current step computes; every step still issues exactly one H2D of its inputs
    # inside the timed region (h2d_bytes_per_step) and reads its loss back.
    copy_stream = torch.cuda.Stream(device=dev)
    pending = {}

    def prefetch():
        with torch.cuda.stream(copy_stream):
            bufs = (view_host.to(dev, non_blocking=True), tpoly_host.to(dev, non_blocking=True),
                    tgt_host.to(dev, non_blocking=True))
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        pending["next"] = (bufs, ev)

    def step(resident: bool):
        for p in all_params:
            p.grad = None
        if resident:
            view, tpoly, tgt, rays = view_d, tpoly_d, tgt_d, rays_d
        else:   # e2e: host buffers in, loss out
            if "next" not in pending:
                prefetch()
            (view, tpoly, tgt), ev = pending.pop("next")
            torch.cuda.current_stream().wait_event(ev)
            for t in (view, tpoly, tgt):
                t.record_stream(torch.cuda.current_stream())
            prefetch()                                  # inputs of the following step
            rays = build_rays(view, intr, W, H)
        view = view.requires_grad_(True) if resident else view.clone().requires_grad_(True)
        if shard_sub:
            # one view, its K sub-frames split across ranks: partial image sums are all-reduced in the
            # forward (collective 1 of SURVEY §8e), Gaussian gradients in the backward (collective 2)
            if my_k:
                ks = slice(my_k[0], my_k[-1] + 1)
                out = render_subframes(stat, dyn, view[ks], Kmat, tpoly[ks].clamp(0, 1), tpoly[ks],
                                       rays[ks] if rays.shape[0] > 1 else rays, bg, W, H)
                local = out["subframes"]
            else:
                out, local = None, torch.zeros(0, 3, H, W, device=dev)
            pred = blur_from_partial_sums(local, K)
        else:
            out = render_subframes(stat, dyn, view, Kmat, tpoly.clamp(0, 1), tpoly, rays, bg, W, H)
            pred = out["render"]
        loss = (pred - tgt).abs().mean()
        if loss.requires_grad:
            loss.backward()
        if world > 1:
            if "fg" not in stats:      # parameter set that receives gradients (fixed across steps)
                stats["fg_params"] = [p for p in all_params if p.grad is not None or shard_sub]
                if shard_sub:          # idle / static-only ranks still contribute zeros
                    stats["fg_params"] = [p for p in stats["fg_params"] if p is not stat.control_xyz
                                          and p is not stat._omega and p is not stat._features_t
                                          and p is not stat._trbf_center and p is not dyn._trbf_center
                                          and p is not dyn._xyz and all(p is not q for q in stat.rgbdecoder.parameters())]
                stats["fg"] = FlatGradients(stats["fg_params"], inplace_shared=not shard_sub)
            stats["fg"].reduce()
            stats["allreduce_bytes"] = stats["fg"].last_collective_elems * 4
        view.grad = None
        stats["out"] = out
        if not resident:
            return float(loss.detach())          # D2H read of the step's result
        return loss

    def timed(resident, steps, warmup, sample_clocks=False):
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()            # nvidia-smi needs a few 100 ms to deliver its first row
        for _ in range(warmup):
            step(resident)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.mark_start()
        if resident:
            _lib.TIMING = {}
            _lib.LAUNCH_COUNT = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(steps):
            flush.zero_()                      # evict L2 between timed iterations (untimed)
            e0.record()
            step(resident)
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        if world > 1:
            dist.barrier()
            t = torch.tensor([tot], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tot = float(t)
        return tot / steps, clocks

    # ---- kernel-side throughput: inputs resident in HBM ----
    _lib.TIMING = {}
    _lib.LAUNCH_COUNT = 0
    for _ in range(args.warmup):
        step(True)
    torch.cuda.synchronize()
    ms, clocks = timed(True, args.steps, 3, sample_clocks=True)
    launches = _lib.LAUNCH_COUNT
    kernel_ms = {n: sum(a.elapsed_time(b) for a, b in ev) / args.steps for n, ev in _lib.TIMING.items()}
    _lib.TIMING = None

    # ---- algorithmic bytes of the dominant kernels (I_eff measured from the forward outputs) ----
    ks_loc = slice(my_k[0], my_k[-1] + 1) if my_k else slice(0, 1)
    ieff, itot = measure_intersections(stat, dyn, view_d[ks_loc], Kmat, tpoly_d[ks_loc], W, H)
    P = K * W * H
    Pjob = P if shard_sub else world * P       # pixels rendered by the whole job per step
    P_loc = len(my_k) * W * H                  # pixels this rank's kernels render per step
    bytes_fwd = 68.0 * ieff + 48.0 * P_loc
    bytes_bwd = 132.0 * ieff + 52.0 * P_loc
    peak, peak_kind = _peak()
    dom = max(("mobgs_blend_fwd", "mobgs_blend_bwd"), key=lambda n: kernel_ms.get(n, 0.0))
    dom_bytes = bytes_bwd if dom == "mobgs_blend_bwd" else bytes_fwd
    ach = dom_bytes / (kernel_ms[dom] * 1e-3) / 1e9
    traffic = issue = lsu = None
    try:    # dram bytes / pipe utilisation of one launch from the committed ncu --set full captures
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            prof = json.load(f)
        traffic = prof.get(args.workload, {}).get(dom)
        issue = prof.get("issue_slot_utilisation", {}).get(dom)
        lsu = prof.get("lsu_pipe_utilisation", {}).get(dom)
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "peak_kind": peak_kind,
                "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "ms_per_launch": kernel_ms[dom], "algorithmic_bytes": dom_bytes,
                "intersections_consumed": ieff, "intersections_listed": itot, "pixels": P_loc,
                "issue_slot_utilisation": issue, "lsu_pipe_utilisation": lsu,
                "note": "the blend kernels are bound by instruction issue and the shared-memory (LSU) pipe, not by HBM "
                        "(ncu figures above from profiles/r1_blend_*_ncu.txt); the HBM fraction is reported because "
                        "BASELINE.json's north_star asks for it. algorithmic_bytes = 132*I_eff + 52*P; since the "
                        "decoder VJP is fused into this kernel it also reads img10/rays/gradients (~100 B/pixel) "
                        "that the byte model does not count"}

    # ---- end to end through the public API with host buffers ----
    e2e_ms, _ = timed(False, args.steps, max(1, args.warmup // 2))
    h2d = tgt_host.numel() * 4 + view_host.numel() * 4 + tpoly_host.numel() * 4

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cstep, cpix, cdesc = cpu_reference_step_factory(WORKLOADS["c1_1k_128_K1"])
            cstep()
            t0 = time.perf_counter()
            n = 0
            while n < 2 or time.perf_counter() - t0 < 12.0:
                cstep(); n += 1
            cdt = (time.perf_counter() - t0) / n
            cpu = {"value": cpix / cdt / 1e6, "unit": "Mpix/s", "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": cdesc + f" ({n} steps, {cdt * 1e3:.0f} ms each)"}
        line = {
            "metric": METRIC, "value": Pjob / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if shard_sub else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "gaussians": N, "static": ns, "dynamic": nd, "width": W,
                       "height": H, "subframes": K, "views_per_step": 1 if shard_sub else world,
                       "parallelism": f"subframes_over_{world}" if shard_sub else f"dp{world}_views",
                       "l2": "flushed between timed iterations (192 MB memset, untimed)",
                       "step": "K-sub-frame render + decode + blur mean + L1 + full backward"
                               + (" + NCCL all-reduce of Gaussian gradients" if world > 1 else "")},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": Pjob / (e2e_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "kernel_ms_per_step": kernel_ms, "clocks": clocks,
        }
        if world > 1:
            line["allreduce_bytes_per_step"] = stats.get("allreduce_bytes")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_intersections(stat, dyn, view, Kmat, tpoly, W, H):
    """I (listed) and I_eff (list entries up to the last one any pixel of the tile blended)."""
    import torch
    from mobgs_b200 import fused
    from mobgs_b200.gaussian_renderer import _dynamic_params, _static_params
    from mobgs_b200.ops import build_tile_lists
    from mobgs_b200 import _lib as L
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard", default="views", choices=["views", "subframes"],
                    help="N>1: 'views' = one view per rank (weak scaling, default); 'subframes' = the K "
                         "sub-frames of one view split across ranks (strong scaling, BASELINE configs[3])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
