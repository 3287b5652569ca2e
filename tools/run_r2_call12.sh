set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_net_gpu.py tests/test_fullsize_gpu.py tests/test_encoder_gpu.py -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_xf_call12.log
for xf in 0 1; do LION_BENCH_XF=$xf TAPS=27 timeout 300 python tools/bench_convs.py 2>&1 | cut -c1-200; done | tee gpurun_out/convs_xf.txt
for ap in 1 0; do
LION_ACT_PASS=$ap python bench.py --allow-knobs --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/bench_r2d_actpass$ap.json 2> gpurun_out/bench_r2d.err; tail -c 400 gpurun_out/bench_r2d.err; cut -c1-500 gpurun_out/bench_r2d_actpass$ap.json
done
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
