#!/bin/bash
# round 2, call 32: persistent cluster global prior (A/B), interleaved conv schedule (A/B + DRAM bytes), fast swish / index math in k_act_grid
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_net_gpu.py -m gpu -q -x -k "global" 2>&1 | tail -5 | tee gpurun_out/c32_pytest_global.log
for v in 1 0; do LION_GP_PERSIST=$v timeout 120 python tools/bench_global.py 2>&1 | tail -2; done | tee gpurun_out/c32_bench_global.txt
B=7 LION_GP_PERSIST=1 timeout 120 python tools/bench_global.py 2>&1 | tail -1 | tee -a gpurun_out/c32_bench_global.txt
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_net_gpu.py tests/test_fullsize_gpu.py tests/test_encoder_gpu.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/c32_pytest.log
for v in 1 0; do echo "LION_CONV_SCHED=$v"; LION_CONV_SCHED=$v TAPS=27 timeout 300 python tools/bench_convs.py 2>&1; done > gpurun_out/c32_convs_sched.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
for v in 1 0; do ONLY="fp3 r=32" LION_CONV_SCHED=$v timeout 300 ncu --metrics $M --clock-control none -k regex:k_conv_tc -s 4 -c 2 --csv --log-file gpurun_out/c32_fp3_sched$v.csv python tools/bench_convs.py > /dev/null 2>&1; done
ONLY="sa0.x conv" LION_CONV_SCHED=1 timeout 300 ncu --metrics $M --clock-control none -k regex:k_conv_tc -s 4 -c 2 --csv --log-file gpurun_out/c32_sa0_sched1.csv python tools/bench_convs.py > /dev/null 2>&1
timeout 300 python tools/timeline_step.py > gpurun_out/c32_timeline.txt 2> gpurun_out/c32_timeline.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/c32_bench.json 2> gpurun_out/c32_bench.err; tail -c 300 gpurun_out/c32_bench.err
LION_CONV_SCHED=0 timeout 600 python bench.py --allow-knobs --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/c32_bench_sched0.json 2> gpurun_out/c32_bench_sched0.err
cut -c1-200 gpurun_out/c32_bench.json gpurun_out/c32_bench_sched0.json
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
