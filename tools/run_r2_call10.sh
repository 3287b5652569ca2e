set -x
mkdir -p gpurun_out
# 1. encoder + trainer tests (fixed)
timeout 600 python -m pytest tests/test_encoder_gpu.py tests/test_trainer_gpu.py -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_encoder_call10.log
# 2. global prior barrier experiments (timing only for mode 2)
for md in 0 1 2; do LION_GP_MODE=$md timeout 120 python tools/bench_global.py 2>&1 | tail -1 | cut -c1-300; done | tee gpurun_out/bench_global_modes.txt
# 3. transform-on-load conv: parity through the PVConv / network tests, then timing per shape
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_net_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_xf_call10.log
for xf in 0 1; do LION_BENCH_XF=$xf TAPS=27 timeout 300 python tools/bench_convs.py 2>&1 | cut -c1-200; done | tee gpurun_out/convs_xf.txt
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
