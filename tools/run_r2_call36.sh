#!/bin/bash
# round 2, call 36: clock-sampler sensitivity (200 ms vs 1000 ms period) and the phase breakdown of the end-to-end pass
set -x
mkdir -p gpurun_out
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs > gpurun_out/c36_bench_1000.json 2> gpurun_out/c36_bench_1000.err
timeout 400 python bench.py --steps 3 --warmup 3 --clock-period-ms 200 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs > gpurun_out/c36_bench_200.json 2> gpurun_out/c36_bench_200.err
timeout 400 python bench.py --steps 3 --warmup 3 --clock-period-ms 5000 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs > gpurun_out/c36_bench_5000.json 2> gpurun_out/c36_bench_5000.err
cut -c1-220 gpurun_out/c36_bench_1000.json gpurun_out/c36_bench_200.json gpurun_out/c36_bench_5000.json
