#!/bin/bash
# A/B of K3 builds: tools/exp_k3_ab.sh <lib1> <lib2> ... (paths under build/), kernel-only bench of configs 5, 3, 2; twice
for rep in 1; do
for lib in "$@"; do
  for c in 5 3 2; do
    TOPSICLE_B200_LIB=$PWD/build/$lib python bench.py --config $c --no-cpu-baseline --no-e2e --no-parse --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$lib config $c', round(d['value']), {k:round(v*1e3,1) for k,v in d['device_ms_per_step'].items()}, round(d['roofline']['pipelined_scan_frac'],3))"
  done
done
done
