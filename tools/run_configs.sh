#!/bin/bash
# Kernel-only + end-to-end numbers for BASELINE.json configs 3, 4, 5 (config 2 is the default bench line).
for c in 3 4 5; do
  python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config$c.json 2> gpurun_out/bench_config$c.err || tail -5 gpurun_out/bench_config$c.err
done
