#!/bin/bash
# round 2, call 19: pooled min/max epilogue for the SA MLP tails, split-N attention context
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_blocks_gpu.py tests/test_net_gpu.py tests/test_encoder_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_call19.log
timeout 300 python tools/timeline_step.py > gpurun_out/timeline_call19.txt 2> gpurun_out/timeline.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
