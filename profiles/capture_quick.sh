#!/bin/bash
# Quick per-kernel capture while iterating: one ncu --set full launch of the kernels named on the command line.
# usage (under gpurun): bash profiles/capture_quick.sh <tag> <kernel regex> [...]
tag=$1; shift
mkdir -p gpurun_out
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/full_${k}_${tag} \
      python bench.py --steps 1 --warmup 3 --skip-cpu --skip-extras > /dev/null 2>&1
done
ls -la gpurun_out/*_${tag}.ncu-rep
