#!/bin/bash
# Round-2 final single-GPU evidence: GPU test suite, sanitizers, ncu launch list + full captures, bench lines.
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
TPS_K1_TMA=0 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_window_bp.py tests/test_gpu_parity.py -x -q -k "cfg0 or cfg2 or cfg4 or cfg12 or many_reads or step2_counts or generic" > gpurun_out/r2_racecheck.log 2>&1; tail -3 gpurun_out/r2_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_window_bp.py tests/test_gpu_parity.py -x -q -k "cfg1 or cfg3 or cfg9 or cfg11 or seed or k1_pack or demo_csv or empty or ends_first" > gpurun_out/r2_memcheck.log 2>&1; tail -3 gpurun_out/r2_memcheck.log
B="python bench.py --no-cpu-baseline --no-e2e --no-parse --distinct-batches 1 --streams 1"
ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 12 --csv --log-file gpurun_out/r2_launches.csv $B --steps 4 --warmup 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tps_ -s 6 -c 3 -f -o gpurun_out/r2_full $B --steps 2 --warmup 2 > /dev/null 2>&1
B5="python bench.py --config 5 --no-cpu-baseline --no-e2e --no-parse --distinct-batches 1 --streams 1"
ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 18 --csv --log-file gpurun_out/r2_launches_c5.csv $B5 --steps 6 --warmup 6 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tps_window -s 3 -c 3 -f -o gpurun_out/r2_full_k3_c5 $B5 --steps 3 --warmup 3 > /dev/null 2>&1
python bench.py > gpurun_out/bench_r2_config2.json 2> gpurun_out/bench_r2_config2.err || tail -5 gpurun_out/bench_r2_config2.err
for c in 3 4 5; do python bench.py --config $c > gpurun_out/bench_r2_config$c.json 2> gpurun_out/bench_r2_config$c.err || tail -5 gpurun_out/bench_r2_config$c.err; done
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_config2_driver_args.json 2> /dev/null
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err
for c in 2 3 4 5; do python -c "
import json;d=json.load(open('gpurun_out/bench_r2_config$c.json'));f=d['e2e_from_fastq'];print($c, round(d['value']), {k:round(v*1e3,1) for k,v in d['device_ms_per_step'].items()}, round(d['roofline']['frac'],3), round(d['roofline']['pipelined_scan_frac'],3), 'e2e', round(d['e2e']['value'],1), 'fastq', round(f['value'],1), round(f['ends_first']['value'],1), 'cpu', round(d['cpu_baseline']['value'],3)); print('   parity', [(p['pattern'],p['telophrase'],p['trc_pass_cpu'],p['trc_pass_gpu'],p['pass_sets_identical'],p['telo_length_exact'],p['telo_length_max_abs_diff'],p['rawcount_tables_identical']) for p in (d['parity_all'] or [])])"; done
python -c "
import json;d=json.load(open('gpurun_out/bench_r2_reference.json'));print('reference arm', d['value'], d['config'], d['cpu_baseline']['cores'])"
