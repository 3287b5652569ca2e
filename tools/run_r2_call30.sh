#!/bin/bash
# round 2, call 30: producer tile decode by shuffle, voxel prep of later levels on the side stream
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_blocks_gpu.py tests/test_net_gpu.py tests/test_encoder_gpu.py tests/test_fullsize_gpu.py tests/test_trainer_gpu.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_call30.log
TAPS=27 timeout 300 python tools/bench_convs.py > gpurun_out/convs_call30.txt 2>&1
timeout 300 python tools/timeline_step.py > gpurun_out/timeline_call30.txt 2> gpurun_out/timeline.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/bench_r2o.json 2> gpurun_out/bench_r2o.err
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
