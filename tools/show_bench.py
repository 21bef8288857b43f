"""Pretty-print the JSON line(s) bench.py emits (reads stdin)."""
import json
import sys

for l in sys.stdin:
    l = l.strip()
    if not l.startswith("{"):
        if l:
            print(l)
        continue
    d = json.loads(l)
    print(f"{d.get('impl','ours')} {d['config'].get('workload')} n_gpus={d['n_gpus']} ms/step={d['ms_per_step']:.3f} "
          f"value={d['value']:.1f} {d['unit']} e2e={d['e2e'].get('ms_per_step')} ms ({d['e2e']['value']:.1f})")
    if "kernel_ms_per_step" in d:
        print("  kernels:", {k.replace('mobgs_', ''): round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
        r = d["roofline"]
        print(f"  roofline: {r['kernel']} {r['achieved']:.1f}/{r['peak']:.0f} GB/s frac={r['frac']:.4f} I_eff={r['intersections_consumed']:.0f}")
        print("  clocks:", d.get("clocks"), "launches:", d.get("gpu_launches"))
    if d.get("cpu_baseline"):
        print("  cpu:", d["cpu_baseline"])
