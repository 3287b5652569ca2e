set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest8.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/pytest8.log
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from lion_b200.third_party.pvcnn.functional import furthest_point_sample_indices
for N, M in ((2048, 1024), (1024, 256), (256, 64), (64, 16)):
    c = torch.randn(32, 3, N, device='cuda')
    furthest_point_sample_indices(c, M); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): furthest_point_sample_indices(c, M)
    e1.record(); torch.cuda.synchronize()
    print('fps', N, M, round(e0.elapsed_time(e1) / 5 * 1000, 1), 'us')
PY
python tools/bench_convs.py > gpurun_out/convs5.jsonl 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/convs5.jsonl'):
    try: d=json.loads(l); print(d.get('shape','sum'), d.get('ms', d.get('sum_ms_per_step_listed')), d.get('tflops_algorithmic',''))
    except Exception: print(l.strip())
PY
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_s3.json 2> gpurun_out/bench_s3.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_s3.json')); print('bench', d['value'], d['phases'], d['roofline']['achieved'])
PY
