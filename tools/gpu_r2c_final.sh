#!/bin/bash
# (GPU box) final evidence of round 2, session 3: GPU suite, smoke, the default bench line, the reference arm, the other workloads
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/r2c_tests.txt 2>&1
tail -4 gpurun_out/r2c_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py 2>gpurun_out/r2c_bench.err > gpurun_out/r2c_bench_c4.json
grep -v Warning gpurun_out/r2c_bench.err | tail -3
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_bench_ref.json 2>gpurun_out/r2c_bench_ref.err
for w in c2_200k_960x540_K7 c3_500k_960x540_K7 c4_1M_1080p_K9 sb_150k_512x288_K9 c4L_1M_1080p_K7; do
  timeout 600 python bench.py --workload $w --no-extras --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2c_bench_$w.json 2>gpurun_out/r2c_bench_$w.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2c_bench_")[1], "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"].get("ms_per_step", 0), 4) if isinstance(d.get("e2e"), dict) else None,
              "value", round(d["value"], 3))
    except Exception as e:
        print(f, "parse failed", e)
PY
