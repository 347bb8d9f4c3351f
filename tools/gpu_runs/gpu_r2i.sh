#!/bin/bash
# round 2, step i: one host thread per decode lane — correctness of the lanes tests, then the lanes sweep
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_parity_fp32.py tests/test_gpu_chained_parallel.py -q -x ) > gpurun_out/pytest_lanes_r2i.log 2>&1
tail -5 gpurun_out/pytest_lanes_r2i.log
run() {  # tag, env...
  tag=$1; shift
  ( env "$@" NOBS_WHISPER_PROFILE_HOST=1 timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2i_$tag.json ) 2> gpurun_out/bench_r2i_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2i_$tag.json"))
    print("$tag", round(d["value"],1), round(d["ms_per_step"],1), d["config"]["stage_ms_per_step"], round(d["roofline"]["achieved"]), round(d["roofline"]["avg_launch_us"],1), d["gpu_launches"])
except Exception as e:
    print("$tag failed", e)
PY
  tail -1 gpurun_out/bench_r2i_$tag.err
}
run old_l2 NOBS_WHISPER_PROJ=0
run old_l3 NOBS_WHISPER_PROJ=0 NOBS_WHISPER_LANES=3
run old_l4 NOBS_WHISPER_PROJ=0 NOBS_WHISPER_LANES=4
run proj_l3 NOBS_WHISPER_PROJ=1 NOBS_WHISPER_LANES=3
run proj_l4 NOBS_WHISPER_PROJ=1 NOBS_WHISPER_LANES=4
run old_l6 NOBS_WHISPER_PROJ=0 NOBS_WHISPER_LANES=6
