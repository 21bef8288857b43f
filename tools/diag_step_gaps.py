"""(GPU box) where does the headline step's time outside the C-ABI kernels go?  Times the resident step three ways:
per-step host sync (bench.py's loop), events only (host runs ahead), and the host's enqueue wall time."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mobgs_b200 import _lib

def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c4_1M_1080p_K7"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _lib.load()
    job = bench.GpuJob(wl, dev, 0, 1, False)
    flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)
    for _ in range(5):
        job.step(True)
    torch.cuda.synchronize()
    out = {"workload": wl}
    # (a) per-step sync
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(20):
        flush.zero_(); e0.record(); job.step(True); e1.record(); e1.synchronize(); tot += e0.elapsed_time(e1)
    out["sync_per_step_ms"] = tot / 20
    # (b) events only
    evs = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); job.step(True); b.record(); evs.append((a, b))
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    out["events_only_ms"] = sum(a.elapsed_time(b) for a, b in evs) / 20
    out["host_enqueue_ms"] = (t1 - t0) * 1e3 / 20
    # (c) kernel sum
    _lib.TIMING = {}
    for _ in range(10):
        flush.zero_(); job.step(True)
    torch.cuda.synchronize()
    km = {n: sum(a.elapsed_time(b) for a, b in ev) / 10 for n, ev in _lib.TIMING.items()}
    _lib.TIMING = None
    out["kernel_ms"] = km
    out["kernel_sum_ms"] = sum(km.values())
    # (d) torch profiler: every device activity of one step
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(3):
                job.step(True)
            torch.cuda.synchronize()
        rows = []
        for ev in prof.key_averages():
            dt = getattr(ev, "device_time_total", 0) or getattr(ev, "cuda_time_total", 0)
            if dt > 0 and ev.device_type.name == "CUDA":
                rows.append((ev.key[:90], dt / 3 / 1e3, ev.count / 3))
        rows.sort(key=lambda r: -r[1])
        out["device_activities_ms_per_step"] = [{"name": n, "ms": round(t, 4), "calls": c} for n, t, c in rows[:40]]
    except Exception as e:  # noqa: BLE001
        out["profiler_error"] = repr(e)
    print(json.dumps(out, indent=1))

main()
