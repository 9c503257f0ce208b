#!/bin/bash
# A/B of library builds under build/: tools/exp_libs.sh "<configs>" lib1.so lib2.so ...  (kernel-only bench, per-kernel us)
cfgs=$1; shift
for lib in "$@"; do
  for c in $cfgs; do
    TOPSICLE_B200_LIB=$PWD/build/$lib python bench.py --config $c --no-cpu-baseline --no-e2e --no-parse --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$lib config $c', round(d['value']), {k:round(v*1e3,1) for k,v in d['device_ms_per_step'].items()}, round(d['roofline']['pipelined_scan_frac'],3))"
  done
done
