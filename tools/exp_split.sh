#!/bin/bash
# GPU experiment 2: context-wide pack/tail streams (split mode) with the bulk-copy K1.
run() { python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-e2e --no-parse "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['device_ms_per_step'].items()}, 'frac', round(d['roofline']['frac'],4))"; }
echo "== parity default (TMA + split)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
echo "== parity no-split, reg K1"; TPS_SPLIT_STREAMS=0 TPS_K1_TMA=0 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
echo "== default: streams 1 2 3"; run --streams 1; run --streams 2; run --streams 3
echo "== no split: streams 1 2 3"; for s in 1 2 3; do TPS_SPLIT_STREAMS=0 run --streams $s; done
for cfg in "3 2" "4 3" "3 3" "6 1" "8 1" "2 4"; do set -- $cfg
  echo "== split, TMA stages=$1 ctas=$2: streams 2, 3"
  TPS_K1_STAGES=$1 TPS_K1_CTAS_PER_SM=$2 run --streams 2
  TPS_K1_STAGES=$1 TPS_K1_CTAS_PER_SM=$2 run --streams 3
done
echo "== other configs, default, streams 2"; for c in 3 4 5; do run --config $c --streams 2; done
