set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_call7.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 1500 gpurun_out/bench_r2b.err; cut -c1-1500 gpurun_out/bench_r2b.json
