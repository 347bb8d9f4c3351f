#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_speculation.py tests/test_gpu_chain.py tests/test_gpu_parity_bf16.py tests/test_gpu_headline_parity.py -q ) > gpurun_out/pytest_r2q.log 2>&1
tail -8 gpurun_out/pytest_r2q.log
for g in 8 0; do
  ( NOBS_WHISPER_GRAPH_ROWS=$g timeout 600 python bench.py --steps 1 --warmup 0 --windows 4 --no-cpu-baseline --latency-clips 200 > gpurun_out/bench_r2q_lat_g$g.json ) 2> gpurun_out/bench_r2q_lat_g$g.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2q_lat_g$g.json"))
print("graph_rows $g latency", d["latency"])
PY
  tail -2 gpurun_out/bench_r2q_lat_g$g.err
done
