# Round-end evidence on one B200: tests, smoke, the default bench line, the reference arm, the
# per-shape convolution timings and one `ncu --set full` capture of the dominant kernel.
# (The ncu launch list -- tools/run_launch_list.sh -- is a separate, 6-minute call.)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo smoke rc=$?
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo ref rc=$?
python tools/bench_convs.py > gpurun_out/final_convs.jsonl 2>&1
ONLY="fp3 r=32" timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 4 -c 1 -f -o gpurun_out/final_conv_fp3 python tools/bench_convs.py > gpurun_out/final_ncu_full.log 2>&1
python tools/bench_metrics.py 16 405 > gpurun_out/final_metrics.jsonl 2>&1
cat gpurun_out/final_bench_n1.json; cat gpurun_out/final_bench_ref.json; tail -1 gpurun_out/final_smoke.log; cat gpurun_out/final_metrics.jsonl
