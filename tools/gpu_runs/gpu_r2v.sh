#!/bin/bash
# encoder attention with two CTAs per SM (single S buffer): kernel + encoder parity tests, bench A/B; 3 lanes with row groups
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_kernels_bf16.py -k "encoder_attention" tests/test_gpu_parity_bf16.py -x -q ) > gpurun_out/pytest_r2v.log 2>&1
tail -6 gpurun_out/pytest_r2v.log
run() {  # tag, env...
  tag=$1; shift
  ( env "$@" timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2v_$tag.json ) 2> gpurun_out/bench_r2v_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2v_$tag.json"))
    c=d["config"]
    o=d["roofline"]["other_kernels"]
    print("$tag", round(d["value"],1), round(d["ms_per_step"],1), c["stage_ms_per_step"], round(d["roofline"]["achieved"]), round(d["roofline"]["avg_launch_us"],1), {k[:14]: (round(v["achieved"]), round(v["share_of_step"],3)) for k,v in o.items()})
except Exception as e:
    print("$tag failed", e)
PY
  tail -1 gpurun_out/bench_r2v_$tag.err
}
run att2 NOBS_WHISPER_ENC_ATT_PER_SM=2
run att1 NOBS_WHISPER_ENC_ATT_PER_SM=1
run att2_l3 NOBS_WHISPER_ENC_ATT_PER_SM=2 NOBS_WHISPER_LANES=3
