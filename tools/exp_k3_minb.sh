#!/bin/bash
# K3 register cap / occupancy A/B: build/lib_minb{9,10,12}.so (nvcc -DTPS_K3N_MINB=N), kernel-only bench of configs 5, 3, 2
for m in 9 10 12; do
  for c in 5 3 2; do
    TOPSICLE_B200_LIB=$PWD/build/lib_minb$m.so python bench.py --config $c --no-cpu-baseline --no-e2e --no-parse --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('minb $m config $c', round(d['value']), {k:round(v*1e3,1) for k,v in d['device_ms_per_step'].items()}, round(d['roofline']['pipelined_scan_frac'],3))"
  done
done
timeout 600 python -m pytest tests/test_gpu_window_bp.py -x -q 2>&1 | tail -2
TPS_K1_TMA=0 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_window_bp.py -x -q -k "cfg0 or cfg2 or cfg4 or cfg12 or many_reads" 2>&1 | tail -4
