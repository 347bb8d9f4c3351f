#!/bin/bash
# round-2 first GPU pass: chain-kernel equality tests, then bench under lane / chain configurations
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
( time timeout 900 python -m pytest tests/test_gpu_chain.py -x -q ) > gpurun_out/pytest_chain_r2a.log 2>&1
tail -5 gpurun_out/pytest_chain_r2a.log
for cfg in "2 1" "3 1" "2 0" "1 1"; do
  set -- $cfg
  tag="r2a_L$1_c$2"
  ( time NOBS_WHISPER_LANES=$1 NOBS_WHISPER_CHAIN=$2 NOBS_WHISPER_PROFILE_HOST=1 timeout 600 \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_$tag.json ) 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$tag.json"))
    print("$tag", round(d["value"],1), round(d["ms_per_step"],1), d["config"]["stage_ms_per_step"], round(d["roofline"]["achieved"]), round(d["roofline"]["avg_launch_us"],1), d["gpu_launches"])
except Exception as e:
    print("$tag failed", e)
PY
  grep "nobs profile" gpurun_out/bench_$tag.err | tail -2
done
