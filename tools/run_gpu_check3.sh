set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest6.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/pytest6.log
python tools/bench_convs.py > gpurun_out/convs3.jsonl 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/convs3.jsonl'):
    try: d=json.loads(l); print(d.get('shape','sum'), d.get('ms', d.get('sum_ms_per_step_listed')), d.get('tflops_algorithmic',''))
    except Exception: print(l.strip())
PY
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_s1.json')); print('bench', d['value'], d['phases'], d['roofline']['achieved'])
PY
