python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-e2e --no-parse"
for cfg in "1 4" "2 4" "2 3" "3 3" "1 3" "2 2"; do set -- $cfg; echo "streams=$1 k1ctas=$2"; TPS_K1_CTAS_PER_SM=$2 $B --streams $1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['device_ms_per_step'].items()})"; done
