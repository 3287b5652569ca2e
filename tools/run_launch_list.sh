set -x
mkdir -p gpurun_out
LION_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches4.csv python bench.py --steps 1 --warmup 1 --ddpm-steps 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench4.log 2>&1
tail -2 gpurun_out/ncu_bench4.log
