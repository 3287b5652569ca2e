set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo smoke rc=$?
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo ref rc=$?
python tools/bench_convs.py > gpurun_out/final_convs.jsonl 2>&1
LION_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 1 --ddpm-steps 3 --no-e2e --no-cpu-baseline > gpurun_out/final_ncu_bench.log 2>&1
ONLY="fp3 r=32" timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 4 -c 1 -f -o gpurun_out/final_conv_fp3 python tools/bench_convs.py > gpurun_out/final_ncu_full.log 2>&1
cat gpurun_out/final_bench_n1.json; cat gpurun_out/final_bench_ref.json; tail -3 gpurun_out/final_smoke.log
