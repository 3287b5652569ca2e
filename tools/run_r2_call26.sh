#!/bin/bash
# round 2, call 27: ygemm with the lean 8-warp epilogue; sparse conv1 for C=32 on/off
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_blocks_gpu.py tests/test_net_gpu.py tests/test_encoder_gpu.py tests/test_fullsize_gpu.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_call26.log
for mc in 32; do
  LION_SPARSE_MINC=$mc timeout 300 python tools/timeline_step.py > gpurun_out/timeline_minc$mc.txt 2> gpurun_out/timeline.err
  LION_SPARSE_MINC=$mc timeout 600 python bench.py --allow-knobs --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/bench_r2m_minc$mc.json 2> gpurun_out/bench_r2m.err
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off -k "regex:k_sparse_conv_gather|k_scatter_compact|k_ygemm" --csv --log-file gpurun_out/r02_sparse_conv_metrics.csv python tools/profile_step.py > gpurun_out/r02_sparse_conv_metrics.log 2>&1
for pdl in 0 1; do LION_PDL=$pdl timeout 120 python tools/bench_global.py >> gpurun_out/global_pdl.txt 2>/dev/null; done
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
