#!/bin/bash
# full capture (with source) of K2 on one config: tools/r2_prof_k2.sh <config> <tag>
c=${1:-2}; tag=${2:-r2k2}
B="python bench.py --config $c --no-cpu-baseline --no-e2e --no-parse --distinct-batches 1 --streams 1"
ncu --set full --clock-control none --import-source on -k regex:"tps_trc" -s 1 -c 1 -f -o gpurun_out/${tag}_full_c$c $B --steps 2 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out/${tag}_full_c$c.ncu-rep
