for lib in "" _k3_128 _k3_512; do
  for c in 2 5; do
    echo "lib=$lib config=$c"
    TOPSICLE_B200_LIB=$PWD/topsicle_b200/libtopsicle_b200$lib.so python bench.py --config $c --steps 16 --warmup 3 --no-cpu-baseline --no-e2e --no-parse --streams 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), {k:round(v,4) for k,v in d['device_ms_per_step'].items()})"
  done
done
