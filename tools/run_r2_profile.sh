# round-2 evidence: per-kernel metrics of one step (every launch), full captures of three kernel classes
set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_step_metrics.csv python tools/profile_step.py > gpurun_out/r02_step_metrics.log 2>&1; tail -2 gpurun_out/r02_step_metrics.log
ONLY="fp3 r=32" timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 4 -c 1 -f -o gpurun_out/r02_conv_fp3 python tools/bench_convs.py > gpurun_out/r02_ncu_conv.log 2>&1; tail -2 gpurun_out/r02_ncu_conv.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_devox_fuse|k_act_grid|k_gp_partial|k_scatter" -c 12 -f -o gpurun_out/r02_hbm_kernels python tools/profile_step.py > gpurun_out/r02_ncu_hbm.log 2>&1; tail -2 gpurun_out/r02_ncu_hbm.log
ls -la gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
