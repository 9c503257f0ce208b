#!/bin/bash
# kernel-only bench of configs 2, 5, 3, 4 with the current build (per-kernel us, pipelined fraction)
for c in ${@:-2 5 3 4}; do
  python bench.py --config $c --no-cpu-baseline --no-e2e --no-parse --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('config $c', round(d['value']), {k:round(v*1e3,1) for k,v in d['device_ms_per_step'].items()}, round(d['roofline']['pipelined_scan_frac'],3))"
done
