#!/bin/bash
# plain-gzip input: reader throughput (parallel inflate vs zlib vs uncompressed) and the scan from .gz on one GPU
nproc
python tools/gz_throughput.py --gbases ${1:-0.8} --level 6 > gpurun_out/r2_gz_throughput.jsonl 2> gpurun_out/r2_gz_throughput.err; cat gpurun_out/r2_gz_throughput.jsonl | cut -c1-1500; tail -3 gpurun_out/r2_gz_throughput.err
