#!/bin/bash
# round 2, step e: fused cluster projections — kernel test, parity subset, bench A/B
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_dec_proj.py -x -q -s ) > gpurun_out/pytest_proj_r2e.log 2>&1
tail -25 gpurun_out/pytest_proj_r2e.log
if ! grep -q " passed" gpurun_out/pytest_proj_r2e.log || grep -q "failed" gpurun_out/pytest_proj_r2e.log; then echo "KERNEL TEST FAILED"; exit 0; fi
( time timeout 900 python -m pytest tests/test_gpu_parity_bf16.py tests/test_gpu_lanes.py tests/test_gpu_headline_parity.py -q ) > gpurun_out/pytest_parity_r2e.log 2>&1
tail -15 gpurun_out/pytest_parity_r2e.log
run() {  # tag, env...
  tag=$1; shift
  ( env "$@" NOBS_WHISPER_PROFILE_HOST=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2e_$tag.json ) 2> gpurun_out/bench_r2e_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2e_$tag.json"))
    print("$tag", round(d["value"],1), round(d["ms_per_step"],1), d["config"]["stage_ms_per_step"], round(d["roofline"]["achieved"]), round(d["roofline"]["avg_launch_us"],1), d["gpu_launches"])
except Exception as e:
    print("$tag failed", e)
PY
  tail -2 gpurun_out/bench_r2e_$tag.err
}
run proj NOBS_WHISPER_PROJ=1
run proj_l3 NOBS_WHISPER_PROJ=1 NOBS_WHISPER_LANES=3
run old NOBS_WHISPER_PROJ=0
