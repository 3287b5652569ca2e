set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest5.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest5.log
LION_TC_MT=4 timeout 600 python -m pytest tests/test_blocks_gpu.py tests/test_net_gpu.py -m gpu -x -q -k "shared_mlp or attention or sa_module or fp_module or pvconv or golden" > gpurun_out/pytest5_mt4.log 2>&1; echo pytest-mt4 rc=$?; tail -3 gpurun_out/pytest5_mt4.log
for mt in 1 2 4; do echo MT=$mt; LION_TC_MT=$mt TAPS=1 python tools/bench_convs.py 2>&1 | tee gpurun_out/convs1x1_mt$mt.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    print(d.get('shape','sum'), d.get('ms', d.get('sum_ms_per_step_listed')))
"; done
ONLY="SA0 mlp1" timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 4 -c 1 -f -o gpurun_out/conv1x1_sa0mlp1 python tools/bench_convs.py > gpurun_out/ncu_1x1.log 2>&1
for mt in 1 4; do LION_TC_MT=$mt python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('MT', $mt, d['value'], d['phases'])
"; done
