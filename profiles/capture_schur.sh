#!/bin/bash
# One ncu --set full launch of the sliding-window Schur kernel on the 4096-window batch (the e2e legs launch it on 512-window
# chunks first: skipped).  usage (under gpurun): bash profiles/capture_schur.sh <tag>
tag=$1
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:schur_tma_kernel --launch-skip 80 -c 6 -f -o gpurun_out/full_schur_tma_${tag} \
    python bench.py --steps 2 --warmup 1 --skip-cpu --skip-assoc > /dev/null 2>&1
ls -la gpurun_out/*_${tag}.ncu-rep
