#!/bin/bash
# round 2, step h: cross-attention at 80 registers — lanes sweep on the multi-launch path and the cluster projections
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  ( env "$@" NOBS_WHISPER_PROFILE_HOST=1 timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2h_$tag.json ) 2> gpurun_out/bench_r2h_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2h_$tag.json"))
    print("$tag", round(d["value"],1), round(d["ms_per_step"],1), d["config"]["stage_ms_per_step"], round(d["roofline"]["achieved"]), round(d["roofline"]["avg_launch_us"],1), d["gpu_launches"])
except Exception as e:
    print("$tag failed", e)
PY
  tail -1 gpurun_out/bench_r2h_$tag.err
}
run old_hi_l2 NOBS_WHISPER_PROJ=0 NOBS_WHISPER_CROSS_LOW_REGS=0
run old_lo_l2 NOBS_WHISPER_PROJ=0
run old_lo_l3 NOBS_WHISPER_PROJ=0 NOBS_WHISPER_LANES=3
run old_lo_l4 NOBS_WHISPER_PROJ=0 NOBS_WHISPER_LANES=4
run proj_lo_l2 NOBS_WHISPER_PROJ=1
run proj_lo_l3 NOBS_WHISPER_PROJ=1 NOBS_WHISPER_LANES=3
run proj_lo_l4 NOBS_WHISPER_PROJ=1 NOBS_WHISPER_LANES=4
