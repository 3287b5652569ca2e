#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_blocks_gpu.py tests/test_net_gpu.py tests/test_encoder_gpu.py tests/test_fullsize_gpu.py tests/test_trainer_gpu.py -m gpu -q 2>&1 | tail -30 | tee gpurun_out/pytest_call21.log
LION_SA_FUSED=0 timeout 600 python -m pytest tests/test_net_gpu.py -m gpu -q -k "graph_determinism" 2>&1 | tail -5 | tee gpurun_out/pytest_call21_unfused.log
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
