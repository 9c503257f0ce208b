#!/bin/bash
# Round-2 multi-GPU evidence on ONE box with 8 GPUs (subsets of its GPUs for N = 1, 2, 4).
set -x
nvidia-smi -L | head -8; nproc; free -g | head -2; df -h /dev/shm | tail -1
python tools/h2d_ceiling.py --gpus 1 2 4 8 --seconds 2 > gpurun_out/r2_h2d_ceiling.jsonl 2> gpurun_out/r2_h2d_ceiling.err; cat gpurun_out/r2_h2d_ceiling.jsonl
timeout 600 python -m pytest tests/test_gpu_multi_device.py -x -q > gpurun_out/r2_pytest_multi_device.log 2>&1; tail -3 gpurun_out/r2_pytest_multi_device.log
python tools/scanner_multi_gpu.py --config 4 --gpus 1 2 4 8 --files-per-gpu 2 --gbases-per-file 0.75 > gpurun_out/r2_scanner_config4.jsonl 2> gpurun_out/r2_scanner_config4.err; cat gpurun_out/r2_scanner_config4.jsonl
python tools/scanner_multi_gpu.py --config 5 --gpus 1 8 --files-per-gpu 3 --gbases-per-file 0.5 > gpurun_out/r2_scanner_config5.jsonl 2> gpurun_out/r2_scanner_config5.err; cat gpurun_out/r2_scanner_config5.jsonl
python tools/cli_e2e.py 32768 --files 8 --reps 2 --both-modes > gpurun_out/r2_cli_8files_8gpu.log 2>&1; grep "^run" gpurun_out/r2_cli_8files_8gpu.log
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --config 4 --steps 20 --warmup 5 --parse-passes 1 > gpurun_out/bench_r2_config4_${n}gpu.json 2> gpurun_out/bench_r2_config4_${n}gpu.err || tail -5 gpurun_out/bench_r2_config4_${n}gpu.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config 5 --steps 21 --warmup 6 --parse-passes 1 > gpurun_out/bench_r2_config5_8gpu.json 2> gpurun_out/bench_r2_config5_8gpu.err || tail -5 gpurun_out/bench_r2_config5_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --parse-passes 1 > gpurun_out/bench_r2_config2_8gpu.json 2> gpurun_out/bench_r2_config2_8gpu.err || tail -5 gpurun_out/bench_r2_config2_8gpu.err
for f in gpurun_out/bench_r2_config4_{2,4,8}gpu.json gpurun_out/bench_r2_config5_8gpu.json gpurun_out/bench_r2_config2_8gpu.json; do python -c "
import json,sys;d=json.load(open('$f'));f=d['e2e_from_fastq'];print('$f', d['n_gpus'], round(d['value']), round(d['roofline']['pipelined_scan_frac'],3), 'e2e', round(d['e2e']['value'],1), 'fastq', round(f['value'],1), 'ends', round(f['ends_first']['value'],1))"; done
