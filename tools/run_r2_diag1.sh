# round-2 diagnostic: where does the fp3 / fp2 / r=8 convolution lose time?  (kernel-only timings)
set -x
mkdir -p gpurun_out
python tools/bench_conv_vs_cudnn.py > gpurun_out/r02_conv_vs_cudnn.jsonl 2> gpurun_out/cudnn_err.log; tail -3 gpurun_out/cudnn_err.log; cut -c1-260 gpurun_out/r02_conv_vs_cudnn.jsonl
for shape in "fp3 r=32" "fp2 r=16" "r=8 128->128" "sa0.x conv"; do
  for cfg in "" "LION_TC_DEBUG=1" "LION_TC_DEBUG=2" "LION_TC_DEBUG=4" "LION_TC_DEBUG=5" "LION_TC_ASTAGES=2" "LION_TC_KG64=4" \
             "LION_TC_STACK=3" "LION_TC_STACK=2"; do
    echo "== $shape :: $cfg"
    env $cfg ONLY="$shape" python tools/bench_convs.py 2>&1 | head -1 | cut -c1-200
  done
done 2>&1 | tee gpurun_out/diag1.txt
