#!/bin/bash
# Round-2 evidence capture (run under gpurun, one GPU): bench lines, ncu launch list, ncu --set full per kernel.
# Numbers printed by a run under ncu are never bench values; the bench lines come from the two plain runs below.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-assoc > /dev/null 2>&1
for k in assemble_kernel plan_kernel schur_dmma_kernel solve_tiled_kernel points_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/full_${k}_r2 \
      python bench.py --steps 1 --warmup 3 --skip-cpu --skip-assoc > /dev/null 2>&1
done
# the sliding-window Schur kernel first runs on the 512-window chunks of the e2e legs (40 launches with these flags: 2 warm-up + 3 timed calls x 4 chunks, twice): skip them
ncu --set full --clock-control none --import-source on -k regex:schur_tma_kernel --launch-skip 40 -c 1 -f -o gpurun_out/full_schur_tma_kernel_r2 \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-assoc > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:match_kernel -c 1 -f -o gpurun_out/full_match_kernel_r2 \
    python bench.py --steps 1 --warmup 3 --skip-cpu --only-assoc --windows 64 > /dev/null 2>&1
ls -la gpurun_out/*_r2*
