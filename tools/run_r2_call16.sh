set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_net_gpu.py tests/test_ode_gpu.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_call16.log
for full in 1 0; do
LION_TC_FULL_SMEM=$full python bench.py --allow-knobs --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/bench_r2f_full$full.json 2> gpurun_out/bench_r2f.err; tail -c 200 gpurun_out/bench_r2f.err; cut -c1-330 gpurun_out/bench_r2f_full$full.json
done
TAPS=27 timeout 300 python tools/bench_convs.py 2>&1 | cut -c1-180 | tee gpurun_out/convs_call16.txt
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
