#!/bin/bash
mkdir -p gpurun_out
for L in 2 3; do
  NOBS_WHISPER_PROJ=0 NOBS_WHISPER_LANES=$L NOBS_WHISPER_TRACE=gpurun_out/trace_r2k_L$L.bin NOBS_WHISPER_TRACE_SKIP=600000 NOBS_WHISPER_TRACE_COUNT=60000 NOBS_WHISPER_TRACE_CAP=800000 \
    timeout 600 python bench.py --steps 1 --warmup 0 --windows 120 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2k_trace_L$L.json 2> gpurun_out/bench_r2k_trace_L$L.err
  python tools/trace_lanes.py gpurun_out/trace_r2k_L$L.bin | tee gpurun_out/lanes_r2k_L$L.txt
done
