#!/bin/bash
# launch list + full capture of the window kernel(s) on one config: tools/r2_prof_k3.sh <config> <tag>
c=${1:-5}; tag=${2:-r2}
B="python bench.py --config $c --no-cpu-baseline --no-e2e --no-parse --distinct-batches 1 --streams 1"
ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 12 --csv --log-file gpurun_out/${tag}_launches_c$c.csv $B --steps 4 --warmup 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tps_window|tps_changepoint" -s 1 -c 1 -f -o gpurun_out/${tag}_full_c$c $B --steps 2 --warmup 1 > /dev/null 2>&1
