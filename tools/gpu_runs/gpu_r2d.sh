#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -q -m gpu ) > gpurun_out/pytest_gpu_r2d.log 2>&1
tail -12 gpurun_out/pytest_gpu_r2d.log
run() {  # tag, env...
  tag=$1; shift
  ( env "$@" NOBS_WHISPER_PROFILE_HOST=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2d_$tag.json ) 2> gpurun_out/bench_r2d_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2d_$tag.json"))
    print("$tag", round(d["value"],1), round(d["ms_per_step"],1), d["config"]["stage_ms_per_step"], round(d["roofline"]["achieved"]), round(d["roofline"]["avg_launch_us"],1), d["gpu_launches"])
except Exception as e:
    print("$tag failed", e)
PY
}
run default NOBS_X=1
run nosapre NOBS_WHISPER_SA_PREFETCH=0
run nok6cl NOBS_WHISPER_K6_CLUSTER=0
run lanes3 NOBS_WHISPER_LANES=3
