#!/bin/bash
# K1 residency / ring depth under co-run with the round-2 tail kernels: pipelined value of configs 5, 3, 2
for ctas in 2 3 4; do for st in 2 3 4; do
  for c in 5 3 2; do
    TPS_K1_CTAS_PER_SM=$ctas TPS_K1_STAGES=$st python bench.py --config $c --no-cpu-baseline --no-e2e --no-parse --steps 40 --warmup 5 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('ctas $ctas stages $st config $c', round(d['value']), round(d['roofline']['pipelined_scan_frac'],3), 'k1', round(d['device_ms_per_step']['k1_pack']*1e3,1))"
  done
done; done
