#!/bin/bash
mkdir -p gpurun_out
for cfg in "fc1 1" "nofc1 0"; do
  set -- $cfg
  NOBS_WHISPER_FC1_FUSED=$2 NOBS_WHISPER_LANES=2 NOBS_WHISPER_TRACE=gpurun_out/trace_r2n_$1.bin NOBS_WHISPER_TRACE_SKIP=600000 NOBS_WHISPER_TRACE_COUNT=40000 NOBS_WHISPER_TRACE_CAP=800000 \
    timeout 600 python bench.py --steps 1 --warmup 0 --windows 120 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2n_trace_$1.json 2> gpurun_out/bench_r2n_trace_$1.err
  python tools/trace_lanes.py gpurun_out/trace_r2n_$1.bin | tee gpurun_out/lanes_r2n_$1.txt
  python tools/trace_lanes.py gpurun_out/trace_r2n_$1.bin 20000 21500 --dump > gpurun_out/lanes_dump_r2n_$1.txt
  rm -f gpurun_out/trace_r2n_$1.bin
done
