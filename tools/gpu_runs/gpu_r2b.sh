#!/bin/bash
# round-2 second GPU pass: chain-kernel timelines (1 and 2 lanes), then the GPU test suite
mkdir -p gpurun_out
for L in 1 2; do
  NOBS_WHISPER_LANES=$L NOBS_WHISPER_TRACE=gpurun_out/trace_r2b_L$L.bin NOBS_WHISPER_TRACE_SKIP=300000 NOBS_WHISPER_TRACE_COUNT=6000 NOBS_WHISPER_TRACE_CAP=400000 \
    timeout 600 python bench.py --steps 1 --warmup 0 --windows 120 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2b_trace_L$L.json 2> gpurun_out/bench_r2b_trace_L$L.err
  tail -2 gpurun_out/bench_r2b_trace_L$L.err
  python tools/trace_dump.py gpurun_out/trace_r2b_L$L.bin 1000 1600 > gpurun_out/timeline_r2b_L$L.txt 2>&1
  tail -4 gpurun_out/timeline_r2b_L$L.txt
done
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/pytest_gpu_r2b.log 2>&1
tail -15 gpurun_out/pytest_gpu_r2b.log
