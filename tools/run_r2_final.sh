#!/bin/bash
# round 2, final evidence on one B200: all GPU tests, smoke, default bench line (every leg), convolution vs cuDNN
# per shape, step timeline, per-launch ncu metrics of one step, ncu --set full of the dominant convolution.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo bench rc=$?
timeout 300 python tools/bench_conv_vs_cudnn.py > gpurun_out/r02_conv_vs_cudnn.jsonl 2> gpurun_out/conv_vs_cudnn.err
timeout 300 python tools/timeline_step.py > gpurun_out/r02_timeline_final.txt 2> gpurun_out/timeline.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_step_metrics_final.csv python tools/profile_step.py > gpurun_out/step_metrics.log 2>&1; tail -2 gpurun_out/step_metrics.log
ONLY="fp3 r=32" timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 4 -c 1 -f -o gpurun_out/r02_conv_fp3_final python tools/bench_convs.py > gpurun_out/ncu_conv.log 2>&1; tail -2 gpurun_out/ncu_conv.log
cat gpurun_out/r02_bench_n1_final.json | cut -c1-600
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
