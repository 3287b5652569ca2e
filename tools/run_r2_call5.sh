set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_blocks_gpu.py -m gpu -q -x -k "conv3d" 2>&1 | tail -5 | tee gpurun_out/pytest_conv_blockissue.log
TAPS=27 python tools/bench_convs.py 2>&1 | cut -c1-200 | tee gpurun_out/convs_blockissue.jsonl
TAPS=27 ITERS=2000 CLOCKS=1 python tools/bench_convs.py 2>&1 | cut -c1-300 | tee gpurun_out/convs_blockissue_clocks.jsonl
LION_TC_STACK=3 ONLY="fp3 r=32" ITERS=2000 CLOCKS=1 python tools/bench_convs.py 2>&1 | head -1 | cut -c1-300 | tee gpurun_out/convs_stack3_clocks.jsonl
