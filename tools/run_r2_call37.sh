#!/bin/bash
# round 2, call 37: in-graph noise fetch for given_noise blocks (parity test + the end-to-end leg of the bench)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_net_gpu.py -m gpu -q -x -k "given_noise or ddpm10 or ddim" 2>&1 | tail -5 | tee gpurun_out/c37_pytest.log
timeout 500 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs > gpurun_out/c37_bench.json 2> gpurun_out/c37_bench.err
tail -c 400 gpurun_out/c37_bench.err
cut -c1-200 gpurun_out/c37_bench.json
