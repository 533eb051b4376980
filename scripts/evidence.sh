#!/bin/bash
# One gpurun call that produces the round's evidence under gpurun_out/ (copied into profiles/ afterwards).
# Bench numbers first (never under ncu), then the launch list, the three --set full captures, the sweep.
set -x
R=${1:-r02}
O=gpurun_out
python bench.py --impl reference --steps 20 --warmup 3 > $O/${R}_bench_reference_arm.json 2> $O/${R}_bench_ref.err
python bench.py --steps 20 --warmup 3 > $O/${R}_bench_final.json 2> $O/${R}_bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_final_launches.csv python bench.py --steps 2 --warmup 1 > $O/${R}_ncu_launches.log 2>&1
for k in parse_v7 decompress_v7 compress_c3; do
  ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o $O/prof_${R}_final_$k python scripts/gpu_probe.py 1024 65536 mixed c3 > $O/${R}_ncu_$k.log 2>&1
done
python scripts/sweep.py 1024 > $O/${R}_chunk_sweep_1gib.jsonl 2> $O/${R}_sweep.err
tail -c 600 $O/${R}_bench_final.json
