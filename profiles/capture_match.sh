#!/bin/bash
# One ncu --set full launch of the association match kernel (cfg 3 sweep).  usage (under gpurun): bash profiles/capture_match.sh <tag>
tag=$1
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:match_kernel -c 1 -f -o gpurun_out/full_match_kernel_${tag} \
    python bench.py --steps 1 --warmup 3 --skip-cpu --only-assoc --windows 64 > /dev/null 2>&1
ls -la gpurun_out/*_${tag}.ncu-rep
