#!/bin/bash
# round 2, call 20: fused SA level-0 kernel (tcgen05 two-pass) -- parity, A/B, ncu traffic
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_blocks_gpu.py -m gpu -q -x -k "sa_module" 2>&1 | tail -5 | tee gpurun_out/pytest_call20a.log
timeout 1200 python -m pytest tests/test_blocks_gpu.py tests/test_net_gpu.py tests/test_encoder_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_call20.log
for fz in 0 1; do
  LION_SA_FUSED=$fz timeout 300 python tools/timeline_step.py > gpurun_out/timeline_safused$fz.txt 2> gpurun_out/timeline.err
done
for fz in 0 1; do
  LION_SA_FUSED=$fz timeout 600 python bench.py --allow-knobs --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/bench_r2j_safused$fz.json 2> gpurun_out/bench_r2j.err
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off -k "regex:k_sa_fused|k_act_pool_minmax|k_ball_query" --csv --log-file gpurun_out/r02_sa_fused_metrics.csv python tools/profile_step.py > gpurun_out/r02_sa_fused_metrics.log 2>&1
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
