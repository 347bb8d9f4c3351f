#!/bin/bash
# round 2, step f: cluster projections at 192 threads / 96 registers — kernel test, timelines, bench A/B
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_dec_proj.py -x -q -s ) > gpurun_out/pytest_proj_r2g.log 2>&1
tail -12 gpurun_out/pytest_proj_r2g.log
if ! grep -q " passed" gpurun_out/pytest_proj_r2g.log || grep -q "failed" gpurun_out/pytest_proj_r2g.log; then echo "KERNEL TEST FAILED"; exit 0; fi
for L in 1 2; do
  NOBS_WHISPER_LANES=$L NOBS_WHISPER_TRACE=gpurun_out/trace_r2g_L$L.bin NOBS_WHISPER_TRACE_SKIP=300000 NOBS_WHISPER_TRACE_COUNT=12000 NOBS_WHISPER_TRACE_CAP=400000 \
    timeout 600 python bench.py --steps 1 --warmup 0 --windows 120 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2g_trace_L$L.json 2> gpurun_out/bench_r2g_trace_L$L.err
  python tools/trace_dump.py gpurun_out/trace_r2g_L$L.bin 1000 1600 > gpurun_out/timeline_r2g_L$L.txt 2>&1
  tail -6 gpurun_out/timeline_r2g_L$L.txt
  rm -f gpurun_out/trace_r2g_L$L.bin
done
run() {  # tag, env...
  tag=$1; shift
  ( env "$@" NOBS_WHISPER_PROFILE_HOST=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2g_$tag.json ) 2> gpurun_out/bench_r2g_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2g_$tag.json"))
    print("$tag", round(d["value"],1), round(d["ms_per_step"],1), d["config"]["stage_ms_per_step"], round(d["roofline"]["achieved"]), round(d["roofline"]["avg_launch_us"],1), d["gpu_launches"])
except Exception as e:
    print("$tag failed", e)
PY
  tail -2 gpurun_out/bench_r2g_$tag.err
}
run proj NOBS_WHISPER_PROJ=1
run proj_l3 NOBS_WHISPER_PROJ=1 NOBS_WHISPER_LANES=3
( time timeout 900 python -m pytest tests/test_gpu_parity_bf16.py tests/test_gpu_lanes.py tests/test_gpu_headline_parity.py tests/test_gpu_full_size.py -q ) > gpurun_out/pytest_parity_r2g.log 2>&1
tail -8 gpurun_out/pytest_parity_r2g.log
