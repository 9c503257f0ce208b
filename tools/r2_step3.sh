#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_gpu_pipeline_scale.py tests/test_gpu_heatmap.py -x -q > gpurun_out/r2n_pytest.log 2>&1; tail -3 gpurun_out/r2n_pytest.log
for c in 3; do
  python bench.py --config $c --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/r2m_c$c.json 2> gpurun_out/r2m_c$c.err || tail -5 gpurun_out/r2m_c$c.err
  python -c "
import json;d=json.load(open('gpurun_out/r2m_c$c.json'));f=d['e2e_from_fastq'];print($c, round(d['value']), {k:round(v*1e3,1) for k,v in d['device_ms_per_step'].items()}, round(d['roofline']['pipelined_scan_frac'],3), 'e2e', round(d['e2e']['value'],1), 'fastq', round(f['value'],1), f['host_seconds_last_pass'], 'ends', round(f['ends_first']['value'],1), f['ends_first']['host_seconds_last_pass'], f['ends_first']['rows_identical_to_whole_read_scan'])"
done
