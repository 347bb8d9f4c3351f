#!/bin/bash
# bench.py under different lane counts / cross-attention configurations: "lanes stages spacing per_sm [mode]"
for cfg in "$@"; do
  set -- $cfg
  tag="L$1_s$2_sp$3_p$4_m${5:-2}"
  NOBS_WHISPER_LANES=$1 NOBS_WHISPER_CROSS_STAGES=$2 NOBS_WHISPER_CROSS_SPACING=$3 NOBS_WHISPER_CROSS_PER_SM=$4 NOBS_WHISPER_CROSS_MODE=${5:-2} NOBS_WHISPER_PROFILE_HOST=1 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("$tag", round(d["value"],1), round(d["ms_per_step"],1), d["config"]["stage_ms_per_step"], round(d["roofline"]["achieved"]), round(d["roofline"]["avg_launch_us"],1))
PY
  tail -1 gpurun_out/bench_$tag.err
done
