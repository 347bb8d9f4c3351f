#!/bin/bash
# asynchronous encoder (windows that become due are encoded while the lane keeps decoding its other audios): parity tests + bench A/B
mkdir -p gpurun_out
( time timeout 420 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_parity_fp32.py tests/test_gpu_chained_parallel.py "tests/test_gpu_full_size.py::test_config3_small_batch_of_64_windows" -x -q --durations=6 ) > gpurun_out/pytest_r2u.log 2>&1
tail -14 gpurun_out/pytest_r2u.log
run() {  # tag, env...
  tag=$1; shift
  ( env "$@" timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2u_$tag.json ) 2> gpurun_out/bench_r2u_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2u_$tag.json"))
    c=d["config"]
    print("$tag", round(d["value"],1), round(d["ms_per_step"],1), c["stage_ms_per_step"], "rows", c["decoder_rows_per_step"], "rounds", c["decoder_rounds_per_step"], round(d["roofline"]["achieved"]), round(d["roofline"]["avg_launch_us"],1), d["gpu_launches"])
except Exception as e:
    print("$tag failed", e)
PY
  tail -1 gpurun_out/bench_r2u_$tag.err
}
run delay3 NOBS_WHISPER_ENC_DELAY=3
run delay0 NOBS_WHISPER_ENC_DELAY=0
run delay3_prio NOBS_WHISPER_ENC_DELAY=3 NOBS_WHISPER_DEC_PRIORITY=1
