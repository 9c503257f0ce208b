#!/bin/bash
# last pass of the round: GPU test suite + the bench lines of configs 2-5 (device-event timing) + the driver's arguments
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
python bench.py > gpurun_out/bench_r2_config2.json 2> gpurun_out/bench_r2_config2.err || tail -5 gpurun_out/bench_r2_config2.err
for c in 3 4 5; do python bench.py --config $c > gpurun_out/bench_r2_config$c.json 2> gpurun_out/bench_r2_config$c.err || tail -5 gpurun_out/bench_r2_config$c.err; done
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_config2_driver_args.json 2> /dev/null
for c in 2 3 4 5; do python -c "
import json;d=json.load(open('gpurun_out/bench_r2_config$c.json'));f=d['e2e_from_fastq'];print($c, round(d['value']), {k:round(v*1e3,1) for k,v in d['device_ms_per_step'].items()}, round(d['roofline']['frac'],3), round(d['roofline']['pipelined_scan_frac'],3), 'e2e', round(d['e2e']['value'],1), 'fastq', round(f['value'],1), round(f['ends_first']['value'],1), 'cpu', round(d['cpu_baseline']['value'],3), d['timing']['device_span_ms'], d['timing']['host_wall_between_barriers_ms']); print('   parity', [(p['pattern'],p['telophrase'],p['trc_pass_cpu'],p['trc_pass_gpu'],p['pass_sets_identical'],p['telo_length_exact'],p['telo_length_max_abs_diff'],p['rawcount_tables_identical']) for p in (d['parity_all'] or [])])"; done
python -c "
import json;d=json.load(open('gpurun_out/bench_r2_config2_driver_args.json'));print('driver args', round(d['value']), round(d['roofline']['pipelined_scan_frac'],3), round(d['e2e']['value'],1), round(d['e2e_from_fastq']['value'],1), d['timing'])"
