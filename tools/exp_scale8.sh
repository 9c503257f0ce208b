#!/bin/bash
# the driver's scaling bench at N = 8 (and 4) with device-span and host-wall timing side by side
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 --no-parse --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        open('gpurun_out/bench_r2_config2_${n}gpu.json','w').write(l)
        d=json.loads(l);print('N=$n config 2 steps 20', round(d['value']), round(d['value']/d['n_gpus']), 'ms/step', round(d['ms_per_step'],4), d['timing']['device_span_ms'], d['timing']['host_wall_between_barriers_ms'], round(d['roofline']['pipelined_scan_frac'],3), 'e2e', round(d['e2e']['value'],1))"
done
