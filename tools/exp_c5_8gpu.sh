#!/bin/bash
# BASELINE config 5 (three pattern sets) on the 8 GPUs of one box, final kernels, device-event timing
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --config 5 --steps 21 --warmup 6 --no-parse --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        open('gpurun_out/bench_r2_config5_8gpu.json','w').write(l)
        d=json.loads(l);print('N=8 config 5', round(d['value']), round(d['value']/8), 'ms/step', round(d['ms_per_step'],4), d['timing']['device_span_ms'], d['timing']['host_wall_between_barriers_ms'], round(d['roofline']['pipelined_scan_frac'],3), 'e2e', round(d['e2e']['value'],1))"
