#!/bin/bash
# GPU experiment 3: how much K1 keeps in flight (stages x stage size x CTAs/SM) against the co-run with the tails
export TL_BRIEF=1
for kb in 16 8; do for cfg in "2 2" "3 2" "4 2" "2 3" "3 3" "2 4" "3 4" "4 4" "6 2"; do set -- $cfg
  echo -n "stageKB=$kb stages=$1 ctas=$2 | "
  TPS_K1_STAGE_KB=$kb TPS_K1_STAGES=$1 TPS_K1_CTAS_PER_SM=$2 python tools/timeline.py 1 8 | tr '\n' ' '
  TPS_K1_STAGE_KB=$kb TPS_K1_STAGES=$1 TPS_K1_CTAS_PER_SM=$2 python tools/timeline.py 3 24
done; done
