#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_sparse_conv_gather" -s 1 -c 1 -f -o gpurun_out/r02_sparse_gather python tools/profile_step.py > gpurun_out/r02_ncu_gather.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_ygemm" -s 1 -c 1 -f -o gpurun_out/r02_ygemm python tools/profile_step.py > gpurun_out/r02_ncu_ygemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv
