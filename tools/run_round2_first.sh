# First GPU call of round 2: the premise and the first run of the experimental
# tap-stacked convolution, plus the GPU-side reference baseline.   gpurun --timeout 900 -- 'bash tools/run_round2_first.sh'
set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_rate tools/umma_rate.cu && timeout 120 /tmp/umma_rate | tee gpurun_out/umma_rate.txt
for s in 3 2; do
  LION_TC_STACK=$s timeout 300 python -m pytest tests/test_blocks_gpu.py -m gpu -q -x -k conv3d > gpurun_out/pytest_stack$s.log 2>&1; echo "stack=$s pytest rc=$?"; tail -5 gpurun_out/pytest_stack$s.log
  LION_TC_STACK=$s TAPS=27 timeout 300 python tools/bench_convs.py > gpurun_out/convs_stack$s.jsonl 2>&1; cat gpurun_out/convs_stack$s.jsonl | cut -c1-160
done
TAPS=27 python tools/bench_convs.py | cut -c1-160
timeout 300 python bench.py --impl reference-gpu | tee gpurun_out/gpu_baseline.json | cut -c1-600
LION_EXTRA_GPU_TESTS=1 timeout 300 python -m pytest tests/test_fullsize_gpu.py -m gpu -q > gpurun_out/pytest_fullsize.log 2>&1; echo "fullsize rc=$?"; tail -5 gpurun_out/pytest_fullsize.log
