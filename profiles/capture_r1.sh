#!/bin/bash
# Round-1 evidence capture (run under gpurun, one GPU): bench lines, ncu launch list, ncu --set full per kernel.
# Numbers printed by a run under ncu are never bench values; the bench lines come from the two plain runs below.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1.json 2> gpurun_out/bench_ref_r1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu > /dev/null 2>&1
for k in assemble_kernel points_kernel schur_splitk_kernel match_kernel cull_tiles_kernel plan_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/full_${k}_r1 \
      python bench.py --steps 1 --warmup 3 --skip-cpu > /dev/null 2>&1
done
ls -la gpurun_out/
