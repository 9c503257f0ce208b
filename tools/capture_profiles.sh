#!/bin/bash
# Round evidence: racecheck, ncu launch list (4 steps), ncu full capture (1 step), un-profiled bench lines.
set -x
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "k1_pack or step2_counts or generic or demo_csv" 2>&1 | tail -4
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "k1_pack or demo_csv or empty or ends_first" 2>&1 | tail -4
B="python bench.py --no-cpu-baseline --no-e2e --no-parse --distinct-batches 1 --streams 1"
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 16 --csv --log-file gpurun_out/r1_launches.csv $B --steps 4 --warmup 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tps_ -s 8 -c 4 -o gpurun_out/r1_full $B --steps 2 --warmup 1 > /dev/null 2>&1
python bench.py > gpurun_out/bench_r1_config2.json 2> gpurun_out/bench_r1_config2.err
for c in 3 4 5; do python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_r1_config$c.json 2> gpurun_out/bench_r1_config$c.err; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.json 2> gpurun_out/bench_r1_reference.err
ls -la gpurun_out
