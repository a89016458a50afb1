#!/bin/bash
# A/B of kernel variants on one box: every tc-viml_b200/build/variants/lib_*.so through the parity tests of the linearisation path
# and the headline bench (device-resident leg).  usage (under gpurun): bash profiles/ab.sh [extra bench flags]
mkdir -p gpurun_out
for so in tc-viml_b200/build/variants/lib_*.so; do
  tag=$(basename $so .so)
  if [ -n "$AB_TESTS" ]; then
    VIML_LIB_PATH=$PWD/$so timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$AB_TESTS" 2>&1 | tail -1
  fi
  VIML_LIB_PATH=$PWD/$so python bench.py --steps 20 --warmup 3 --skip-cpu --skip-extras "$@" 2> gpurun_out/ab_$tag.err > gpurun_out/ab_$tag.json || { echo "$tag FAILED"; tail -2 gpurun_out/ab_$tag.err; continue; }
  python - "$tag" gpurun_out/ab_$tag.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
print(sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "assemble ms %.4f" % d["roofline"]["avg_launch_ms"], "frac %.4f" % d["roofline"]["frac"], "e2e ms %.3f" % d["e2e"]["ms_per_step"])
PY
done
