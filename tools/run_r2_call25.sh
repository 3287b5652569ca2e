#!/bin/bash
# round 2, call 25: gather v3 (non-inlined flush)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_blocks_gpu.py tests/test_net_gpu.py tests/test_encoder_gpu.py tests/test_fullsize_gpu.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_call25.log
timeout 300 python tools/timeline_step.py > gpurun_out/timeline_call25.txt 2> gpurun_out/timeline.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off -k "regex:k_sparse_conv_gather|k_scatter_compact|k_ygemm" --csv --log-file gpurun_out/r02_sparse_conv_metrics.csv python tools/profile_step.py > gpurun_out/r02_sparse_conv_metrics.log 2>&1
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
