set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_rate tools/umma_rate.cu && timeout 300 /tmp/umma_rate | tee gpurun_out/umma_rate_r02.txt
timeout 600 python -m pytest tests/test_point_ops_gpu.py -m gpu -q 2>&1 | tail -30 | tee gpurun_out/pytest_point_ops.log
timeout 600 python -m pytest tests/test_blocks_gpu.py -m gpu -q -k "b32" 2>&1 | tail -15 | tee gpurun_out/pytest_conv_b32.log
