#!/bin/bash
# device-span vs host-wall timing of the K timed steps at N ranks: tools/exp_scale_timing.sh <N> [config]
n=${1:-2}; c=${2:-2}
for steps in 20 200; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$((steps % 7)) bench.py --gpus $n --config $c --steps $steps --warmup 5 --no-e2e --no-parse --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l);print('N=$n config $c steps $steps', round(d['value']), round(d['value']/d['n_gpus']), 'ms/step', round(d['ms_per_step'],4), d['timing']['device_span_ms'], d['timing']['host_wall_between_barriers_ms'], round(d['roofline']['pipelined_scan_frac'],3))"
done
python bench.py --config $c --steps 20 --warmup 5 --no-e2e --no-parse --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l);print('N=1 config $c steps 20', round(d['value']), 'ms/step', round(d['ms_per_step'],4), d['timing']['device_span_ms'], d['timing']['host_wall_between_barriers_ms'], round(d['roofline']['pipelined_scan_frac'],3))"
