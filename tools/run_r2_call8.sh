set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_net_gpu.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_call8.log
for sep in 0 1; do
LION_AFFINE_SEPARATE=$sep python bench.py --allow-knobs --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/bench_r2c_sep$sep.json 2> gpurun_out/bench_r2c.err; tail -c 600 gpurun_out/bench_r2c.err; cut -c1-700 gpurun_out/bench_r2c_sep$sep.json
done
