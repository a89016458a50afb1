#!/bin/bash
# A/B of kernel variants on one box: every tc-viml_b200/build/variants/lib_*.so through the headline bench (device-resident leg).
# usage (under gpurun): bash profiles/ab.sh [extra bench flags]
mkdir -p gpurun_out
for so in tc-viml_b200/build/variants/lib_*.so; do
  tag=$(basename $so .so)
  for rep in 1; do
    VIML_LIB_PATH=$PWD/$so python bench.py --steps 20 --warmup 3 --skip-cpu --skip-extras "$@" 2> gpurun_out/ab_$tag.err > gpurun_out/ab_$tag.json
    python - "$tag" gpurun_out/ab_$tag.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
print(sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "assemble ms %.4f" % (d["roofline"]["achieved"] and 679.3 / d["roofline"]["achieved"]), "frac %.4f" % d["roofline"]["frac"], "e2e ms %.3f" % d["e2e"]["ms_per_step"])
PY
  done
done
