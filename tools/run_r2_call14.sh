set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; echo rc=$?; tail -c 1200 gpurun_out/bench_r2_n2.err; cut -c1-3000 gpurun_out/bench_r2_n2.json
nvidia-smi --query-gpu=index,name,temperature.gpu,clocks.sm --format=csv
