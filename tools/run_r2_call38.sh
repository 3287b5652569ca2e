#!/bin/bash
# round 2, call 38: given_noise block test; end-to-end leg with the noise upload streamed under the loops vs finished first
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_net_gpu.py -m gpu -q -x -k "given_noise" 2>&1 | grep -E "passed|failed|Error|assert" | tail -8 | tee gpurun_out/c38_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --e2e-compare --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs > gpurun_out/c38_bench.json 2> gpurun_out/c38_bench.err
tail -c 300 gpurun_out/c38_bench.err
cut -c1-200 gpurun_out/c38_bench.json
