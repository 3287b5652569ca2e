set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_issue tools/umma_issue.cu && timeout 300 /tmp/umma_issue | tee gpurun_out/umma_issue.txt
timeout 600 python -m pytest tests/test_point_ops_gpu.py -m gpu -q 2>&1 | tail -30 | tee gpurun_out/pytest_point_ops.log
