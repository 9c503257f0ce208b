#!/bin/bash
# GPU experiment: bulk-copy (TMA) staged K1 and split-stream tails against the register-staged default.
# Parity first (K1 + full parity file under each mode), then bench lines (kernel-only, no CPU legs).
run() { python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-e2e --no-parse "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['device_ms_per_step'].items()}, 'frac', round(d['roofline']['frac'],4))"; }
echo "== parity TMA"; TPS_K1_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random_sweep.py -x -q 2>&1 | tail -2
echo "== parity TMA stages 2 ctas 1"; TPS_K1_TMA=1 TPS_K1_STAGES=2 TPS_K1_CTAS_PER_SM=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "k1_pack or demo_csv or random" 2>&1 | tail -2
echo "== parity SPLIT+TMA"; TPS_SPLIT_STREAMS=1 TPS_K1_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline_scale.py -x -q 2>&1 | tail -2
echo "== base reg K1: streams 1, 2"; run --streams 1; run --streams 2
for cfg in "3 3" "4 3" "3 4" "4 2" "6 2" "5 2" "8 1" "12 1"; do set -- $cfg
  echo "== TMA stages=$1 ctas=$2: streams 1, 2"
  TPS_K1_TMA=1 TPS_K1_STAGES=$1 TPS_K1_CTAS_PER_SM=$2 run --streams 1
  TPS_K1_TMA=1 TPS_K1_STAGES=$1 TPS_K1_CTAS_PER_SM=$2 run --streams 2
done
echo "== SPLIT, reg K1 ctas 4/3/2"
for c in 4 3 2; do TPS_SPLIT_STREAMS=1 TPS_K1_CTAS_PER_SM=$c run --streams 2; done
echo "== SPLIT + TMA"
for cfg in "4 3" "4 2" "6 2" "8 1" "12 1"; do set -- $cfg
  echo "stages=$1 ctas=$2"; TPS_SPLIT_STREAMS=1 TPS_K1_TMA=1 TPS_K1_STAGES=$1 TPS_K1_CTAS_PER_SM=$2 run --streams 2
  TPS_SPLIT_STREAMS=1 TPS_K1_TMA=1 TPS_K1_STAGES=$1 TPS_K1_CTAS_PER_SM=$2 run --streams 3
done
