#!/bin/bash
# round 2, step l: speculative first fallback (shadow pass) — exactness tests, then bench A/B
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_speculation.py tests/test_gpu_lanes.py tests/test_gpu_parity_fp32.py -q -x ) > gpurun_out/pytest_spec_r2l.log 2>&1
tail -15 gpurun_out/pytest_spec_r2l.log
run() {  # tag, env...
  tag=$1; shift
  ( env "$@" NOBS_WHISPER_PROJ=0 NOBS_WHISPER_PROFILE_HOST=1 timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2l_$tag.json ) 2> gpurun_out/bench_r2l_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2l_$tag.json"))
    c=d["config"]
    print("$tag", round(d["value"],1), round(d["ms_per_step"],1), c["stage_ms_per_step"], "rows", c["decoder_rows_per_step"], "rounds", c["decoder_rounds_per_step"], round(d["roofline"]["achieved"]), round(d["roofline"]["avg_launch_us"],1), d["gpu_launches"])
except Exception as e:
    print("$tag failed", e)
PY
  tail -1 gpurun_out/bench_r2l_$tag.err
}
run spec_l3 NOBS_WHISPER_SPECULATE=1 NOBS_WHISPER_LANES=3
run spec_l2 NOBS_WHISPER_SPECULATE=1 NOBS_WHISPER_LANES=2
run spec_l4 NOBS_WHISPER_SPECULATE=1 NOBS_WHISPER_LANES=4
run nospec_l3 NOBS_WHISPER_SPECULATE=0 NOBS_WHISPER_LANES=3
