#!/bin/bash
# Round-2 single-GPU evidence: GPU test suite, bench lines of configs 2-5 with CPU baseline + parity, parity rate at 10 k reads.
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
python bench.py > gpurun_out/bench_r2_config2.json 2> gpurun_out/bench_r2_config2.err || tail -5 gpurun_out/bench_r2_config2.err
for c in 3 4 5; do python bench.py --config $c > gpurun_out/bench_r2_config$c.json 2> gpurun_out/bench_r2_config$c.err || tail -5 gpurun_out/bench_r2_config$c.err; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err
TPS_PARITY_READS=10000 timeout 1200 python -m pytest tests/test_gpu_parity_rate.py -x -q > gpurun_out/r2_parity_rate.log 2>&1; tail -3 gpurun_out/r2_parity_rate.log
for c in 2 3 4 5; do python -c "
import json;d=json.load(open('gpurun_out/bench_r2_config$c.json'));print($c, round(d['value']), {k:round(v*1e3,1) for k,v in d['device_ms_per_step'].items()}, round(d['roofline']['pipelined_scan_frac'],3), 'e2e', round(d['e2e']['value'],1), 'fastq', round(d['e2e_from_fastq']['value'],1), round(d['e2e_from_fastq']['ends_first']['value'],1)); print('   parity', [(p['pattern'],p['telophrase'],p['trc_pass_cpu'],p['trc_pass_gpu'],p['pass_sets_identical'],p['telo_length_exact'],p['telo_length_max_abs_diff'],p['rawcount_tables_identical']) for p in (d['parity_all'] or [])])"; done
