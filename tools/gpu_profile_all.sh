#!/bin/bash
# (GPU box) everything profiles/ is refreshed from: bench line, reference arm, launch list, ncu --set full captures.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/launches.csv \
  python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
for kn in blend_bwd blend_fwd tile_rank_sort synth_project_bwd synth_project_fwd; do
  out=$kn; [ $kn = tile_rank_sort ] && out=tile_sort
  ncu --set full --clock-control none --import-source on -k regex:$kn -s 2 -c 1 -f -o gpurun_out/$out \
    python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/ncu_$out.log 2>&1
done
ls -la gpurun_out | head -30
