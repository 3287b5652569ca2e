set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_net_gpu.py tests/test_blocks_gpu.py tests/test_scheduler_route.py -m gpu -q -x > gpurun_out/pytest9.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest9.log
for cfg in "1 0" "0 1" "1 1"; do set -- $cfg; LION_GP_CLUSTER=$1 LION_AFFINE_PREP=$2 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_s4.json 2> gpurun_out/bench_s4.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_s4.json')); print('cluster=$1 affine_prep=$2', round(d['value'],3), d['phases'], d['gpu_launches'])
PY
done
