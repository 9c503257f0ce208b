#!/bin/bash
tag=${1:-r2x}
timeout 900 python -m pytest tests/test_gpu_window_bp.py -x -q > gpurun_out/${tag}_pytest_bp.log 2>&1; tail -5 gpurun_out/${tag}_pytest_bp.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random_sweep.py -x -q > gpurun_out/${tag}_pytest_par.log 2>&1; tail -4 gpurun_out/${tag}_pytest_par.log
for c in 5 3 2; do
  python bench.py --config $c --no-cpu-baseline --no-e2e --no-parse --steps 20 --warmup 3 > gpurun_out/${tag}_c$c.json 2> gpurun_out/${tag}_c$c.err || tail -5 gpurun_out/${tag}_c$c.err
  python -c "
import json;d=json.load(open('gpurun_out/${tag}_c$c.json'));print($c, d['value'], d['device_ms_per_step'], d['roofline']['pipelined_scan_frac'], d['trc_pass_reads_per_step'])"
done
bash tools/r2_prof_k3.sh 5 ${tag}
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_window_bp.py -x -q -k "cfg0 or cfg2 or cfg4 or cfg12 or many_reads" > gpurun_out/${tag}_racecheck.log 2>&1; tail -3 gpurun_out/${tag}_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_window_bp.py -x -q -k "cfg1 or cfg3 or cfg9 or cfg11 or seed" > gpurun_out/${tag}_memcheck.log 2>&1; tail -3 gpurun_out/${tag}_memcheck.log
