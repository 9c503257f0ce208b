#!/bin/bash
# Round-2 starting point: per-kernel launch list and a full capture of K3/K4 on configs 3 and 5.
for c in 5 3; do
  B="python bench.py --config $c --no-cpu-baseline --no-e2e --no-parse --distinct-batches 1 --streams 1"
  ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 16 --csv --log-file gpurun_out/r2a_launches_c$c.csv $B --steps 4 --warmup 2 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"tps_window|tps_changepoint" -s 2 -c 2 -f -o gpurun_out/r2a_full_c$c $B --steps 2 --warmup 1 > /dev/null 2>&1
done
ls -la gpurun_out | tail -8
