#!/bin/bash
# round 2, step j: what saturates with more lanes?  cross-attention queue depth and weight residency in L2
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  ( env "$@" NOBS_WHISPER_PROJ=0 NOBS_WHISPER_PROFILE_HOST=1 timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2j_$tag.json ) 2> gpurun_out/bench_r2j_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2j_$tag.json"))
    print("$tag", round(d["value"],1), round(d["ms_per_step"],1), d["config"]["stage_ms_per_step"], round(d["roofline"]["achieved"]), round(d["roofline"]["avg_launch_us"],1), d["gpu_launches"])
except Exception as e:
    print("$tag failed", e)
PY
  tail -1 gpurun_out/bench_r2j_$tag.err
}
run st2_l2 NOBS_WHISPER_CROSS_STAGES=2
run st2_l3 NOBS_WHISPER_CROSS_STAGES=2 NOBS_WHISPER_LANES=3
run wkeep_l2 NOBS_WHISPER_W_EVICT_LAST=1
run wkeep_l3 NOBS_WHISPER_W_EVICT_LAST=1 NOBS_WHISPER_LANES=3
run persm1_l3 NOBS_WHISPER_CROSS_PER_SM=1 NOBS_WHISPER_CROSS_STAGES=6 NOBS_WHISPER_LANES=3
run st2_wkeep_l3 NOBS_WHISPER_CROSS_STAGES=2 NOBS_WHISPER_W_EVICT_LAST=1 NOBS_WHISPER_LANES=3
