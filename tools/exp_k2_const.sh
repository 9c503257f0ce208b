#!/bin/bash
# A/B of the K2 with the literal set in its instructions (TPS_K2_CONST=0: table-driven K2): kernel-only bench
for kc in 1 0; do
  for c in 2 5 3; do
    TPS_K2_CONST=$kc python bench.py --config $c --no-cpu-baseline --no-e2e --no-parse --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('k2const=$kc config $c', round(d['value']), {k:round(v*1e3,1) for k,v in d['device_ms_per_step'].items()}, round(d['roofline']['pipelined_scan_frac'],3))"
  done
done
