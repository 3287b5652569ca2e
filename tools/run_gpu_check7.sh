set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_net_gpu.py tests/test_scheduler_route.py -m gpu -q -x > gpurun_out/pytest10.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest10.log
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_s5.json 2> gpurun_out/bench_s5.err; echo rc=$?; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_s5.json')); print('bench', round(d['value'],3), d['ms_per_step'], d['phases'], 'e2e', round(d['e2e']['value'],3))
PY
