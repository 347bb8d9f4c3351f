#!/bin/bash
# row groups in the tcgen05 cross-attention (kernel test, speculation / beam tests, bench A/B) + turbo latency with the fused projection paths
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_kernels_bf16.py -k cross_attention -x -q ) > gpurun_out/pytest_r2t_kernel.log 2>&1
tail -6 gpurun_out/pytest_r2t_kernel.log
( time timeout 400 python -m pytest tests/test_gpu_speculation.py tests/test_gpu_parity_bf16.py -x -q ) > gpurun_out/pytest_r2t_spec.log 2>&1
tail -6 gpurun_out/pytest_r2t_spec.log
for cfg in "X=0" "NOBS_WHISPER_PROJ=1" "NOBS_WHISPER_CHAIN=1"; do
  echo "latency $cfg"
  ( env $cfg timeout 200 python tools/latency_trace.py 8 ) 2>&1 | tail -3
done
run() {  # tag, env...
  tag=$1; shift
  ( env "$@" timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --latency-clips 0 > gpurun_out/bench_r2t_$tag.json ) 2> gpurun_out/bench_r2t_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2t_$tag.json"))
    c=d["config"]
    print("$tag", round(d["value"],1), round(d["ms_per_step"],1), c["stage_ms_per_step"], "rows", c["decoder_rows_per_step"], "rounds", c["decoder_rounds_per_step"], round(d["roofline"]["achieved"]), round(d["roofline"]["avg_launch_us"],1), d["gpu_launches"])
except Exception as e:
    print("$tag failed", e)
PY
  tail -1 gpurun_out/bench_r2t_$tag.err
}
run groups1 NOBS_WHISPER_CROSS_GROUPS=1
run groups0 NOBS_WHISPER_CROSS_GROUPS=0
