set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_net_gpu.py -m gpu -q -x -k "global" 2>&1 | tail -8 | tee gpurun_out/pytest_global_call9.log
timeout 120 python tools/bench_global.py 2>&1 | tail -2 | tee gpurun_out/bench_global_persistent.json
LION_GP_PERSISTENT=0 timeout 120 python tools/bench_global.py 2>&1 | tail -2 | tee gpurun_out/bench_global_layers.json
timeout 600 python -m pytest tests/test_encoder_gpu.py tests/test_trainer_gpu.py -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/pytest_encoder_call9.log
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_call9.log
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
