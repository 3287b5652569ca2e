#!/bin/bash
# round 2, call 33: persistent global prior with clusters of 8 / 4 / 2 (co-residency decides), round-based conv schedule at r = 32
set -x
mkdir -p gpurun_out
LION_VERBOSE=1 timeout 300 python -m pytest tests/test_net_gpu.py -m gpu -q -x -k "global" 2>&1 | tail -6 | tee gpurun_out/c33_pytest_global.log
for v in 1 0 2; do LION_VERBOSE=1 LION_GP_PERSIST=$v timeout 120 python tools/bench_global.py 2>&1 | tail -2; done | tee gpurun_out/c33_bench_global.txt
B=7 timeout 120 python tools/bench_global.py 2>&1 | tail -1 | tee -a gpurun_out/c33_bench_global.txt
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_net_gpu.py tests/test_fullsize_gpu.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/c33_pytest.log
TAPS=27 timeout 300 python tools/bench_convs.py > gpurun_out/c33_convs.txt 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
ONLY="fp3 r=32" timeout 300 ncu --metrics $M --clock-control none -k regex:k_conv_tc -s 4 -c 2 --csv --log-file gpurun_out/c33_fp3.csv python tools/bench_convs.py > /dev/null 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/c33_bench.json 2> gpurun_out/c33_bench.err; tail -c 300 gpurun_out/c33_bench.err
LION_GP_PERSIST=0 timeout 600 python bench.py --allow-knobs --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/c33_bench_gp0.json 2> gpurun_out/c33_bench_gp0.err
cut -c1-200 gpurun_out/c33_bench.json gpurun_out/c33_bench_gp0.json
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
