#!/bin/bash
# A/B of K3's split mode (TPS_K3_SPLIT=0: a read is always one CTA's work): kernel-only bench of configs 2, 4, 5, 3
for sp in 1 0; do
  for c in 2 4 5 3; do
    TPS_K3_SPLIT=$sp python bench.py --config $c --no-cpu-baseline --no-e2e --no-parse --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('split=$sp config $c', round(d['value']), {k:round(v*1e3,1) for k,v in d['device_ms_per_step'].items()}, round(d['roofline']['pipelined_scan_frac'],3))"
  done
done
