#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/grid_sync_bench.py > gpurun_out/grid_sync_r2c.txt 2>&1; cat gpurun_out/grid_sync_r2c.txt
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -3
for L in 1 2; do
  NOBS_WHISPER_LANES=$L NOBS_WHISPER_TRACE=gpurun_out/trace_r2c_L$L.bin NOBS_WHISPER_TRACE_SKIP=300000 NOBS_WHISPER_TRACE_COUNT=8000 NOBS_WHISPER_TRACE_CAP=400000 \
    timeout 600 python bench.py --steps 1 --warmup 0 --windows 120 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2c_trace_L$L.json 2> gpurun_out/bench_r2c_trace_L$L.err
  python tools/trace_dump.py gpurun_out/trace_r2c_L$L.bin 1000 1500 > gpurun_out/timeline_r2c_L$L.txt 2>&1
  tail -4 gpurun_out/timeline_r2c_L$L.txt
done
NOBS_WHISPER_LANES=2 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2c_L2.json 2> gpurun_out/bench_r2c_L2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2c_L2.json')); print('L2 chain v2', round(d['value'],1), round(d['ms_per_step'],1), d['config']['stage_ms_per_step'])"
( time timeout 1800 python -m pytest tests/ -q -m gpu ) > gpurun_out/pytest_gpu_r2c.log 2>&1
tail -15 gpurun_out/pytest_gpu_r2c.log
