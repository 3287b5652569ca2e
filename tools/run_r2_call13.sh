set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_net_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_xf_call13.log
(LION_BENCH_KG=4 TAPS=27 timeout 300 python tools/bench_convs.py; LION_BENCH_XF=1 TAPS=27 timeout 300 python tools/bench_convs.py) 2>&1 | cut -c1-200 | tee gpurun_out/convs_xf2.txt
timeout 600 python -m pytest tests/test_ode_gpu.py tests/test_point_ops_backward_gpu.py tests/test_encoder_gpu.py tests/test_trainer_gpu.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_misc_call13.log
for ap in 0 1; do
LION_ACT_PASS=$ap python bench.py --allow-knobs --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity --no-extra-configs --no-e2e > gpurun_out/bench_r2e_actpass$ap.json 2> gpurun_out/bench_r2e.err; tail -c 300 gpurun_out/bench_r2e.err; cut -c1-400 gpurun_out/bench_r2e_actpass$ap.json
done
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
