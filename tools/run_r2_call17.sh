#!/bin/bash
# round 2, call 17: step timeline (stamps) + side-stream neighbour search A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_net_gpu.py tests/test_encoder_gpu.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_call17.log
for nn in 0 1; do
  LION_AUX_NN=$nn timeout 300 python tools/timeline_step.py > gpurun_out/timeline_auxnn$nn.txt 2> gpurun_out/timeline.err
done
for nn in 0 1; do
  LION_AUX_NN=$nn timeout 600 python bench.py --allow-knobs --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/bench_r2g_auxnn$nn.json 2> gpurun_out/bench_r2g.err
done
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
