#!/bin/bash
# round 2, call 18: SM sharing between the side stream and the convolutions (carve-out + 196 KB cap), temb on the side stream
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_net_gpu.py tests/test_encoder_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_call18.log
for kb in 0 196; do
  LION_TC_SHARE_KB=$kb timeout 300 python tools/timeline_step.py > gpurun_out/timeline_share$kb.txt 2> gpurun_out/timeline.err
done
for kb in 0 196; do
  LION_TC_SHARE_KB=$kb timeout 600 python bench.py --allow-knobs --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/bench_r2h_share$kb.json 2> gpurun_out/bench_r2h.err
done
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
