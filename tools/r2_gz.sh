#!/bin/bash
nproc
python tools/gz_throughput.py --gbases 1.1 --level 6 > gpurun_out/r2_gz_throughput.jsonl 2> gpurun_out/r2_gz_throughput.err; cat gpurun_out/r2_gz_throughput.jsonl | cut -c1-1500; tail -3 gpurun_out/r2_gz_throughput.err
TPS_PGZ_DEBUG=1 python tools/gz_throughput.py --gbases 0.3 --level 6 --no-zlib 2>&1 | grep "\[pgz\]" | head -12
