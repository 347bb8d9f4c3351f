#!/bin/bash
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  ( env "$@" NOBS_WHISPER_PROFILE_HOST=1 timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2o_$tag.json ) 2> gpurun_out/bench_r2o_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2o_$tag.json"))
    c=d["config"]
    print("$tag", round(d["value"],1), round(d["ms_per_step"],1), c["stage_ms_per_step"], "rows", c["decoder_rows_per_step"], "rounds", c["decoder_rounds_per_step"], round(d["roofline"]["achieved"]), round(d["roofline"]["avg_launch_us"],1), d["gpu_launches"])
except Exception as e:
    print("$tag failed", e)
PY
  tail -1 gpurun_out/bench_r2o_$tag.err
}
run bal_l2 NOBS_WHISPER_LANES=2
run nobal_l2 NOBS_WHISPER_LANES=2 NOBS_WHISPER_CROSS_BALANCE=0
run bal_l3 NOBS_WHISPER_LANES=3
run bal_l2_nofc1 NOBS_WHISPER_LANES=2 NOBS_WHISPER_FC1_FUSED=0
