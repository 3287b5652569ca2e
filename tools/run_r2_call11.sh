set -x
mkdir -p gpurun_out
for md in 0 2; do LION_GP_MODE=$md timeout 120 python tools/bench_global.py 2>&1 | tail -1 | cut -c1-300; done | tee gpurun_out/bench_global_modes2.txt
LION_GP_PERSISTENT=0 timeout 120 python tools/bench_global.py 2>&1 | tail -1 | cut -c1-300 | tee -a gpurun_out/bench_global_modes2.txt
timeout 300 python -m pytest tests/test_net_gpu.py -m gpu -q -x -k "global" 2>&1 | tail -3
LION_ACT_PASS=1 timeout 600 python -m pytest tests/test_point_ops_backward_gpu.py tests/test_encoder_gpu.py tests/test_trainer_gpu.py -m gpu -q 2>&1 | tail -12 | tee gpurun_out/pytest_bwd_enc_call11.log
ONLY="sa1.0 conv2" LION_BENCH_XF=1 ITERS=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 8 python tools/bench_convs.py > gpurun_out/sanitizer_xf.log 2>&1; tail -60 gpurun_out/sanitizer_xf.log | cut -c1-260
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
