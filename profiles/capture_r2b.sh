#!/bin/bash
# Late round-2 capture (run under gpurun, one GPU) after the assemble_kernel / plan_kernel changes: bench lines of both arms, ncu
# launch list, ncu --set full of the two changed kernels.  The other kernels' captures (profiles/capture_r2.sh) still describe the
# code in the tree.  Numbers printed by a run under ncu are never bench values.
mkdir -p gpurun_out
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-assoc > /dev/null 2>&1
for k in assemble_kernel plan_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/full_${k}_r2 \
      python bench.py --steps 1 --warmup 3 --skip-cpu --skip-extras > /dev/null 2>&1
done
ls -la gpurun_out/*_r2*
