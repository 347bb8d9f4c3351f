#!/bin/bash
# round 2: 2 x B200, weak scaling (120 clips per GPU) and strong scaling (the one-hour workload split over the GPUs)
mkdir -p gpurun_out
for mode in weak strong; do
  ( timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline --latency-clips 0 --scaling $mode > gpurun_out/bench_r2s_2gpu_$mode.json ) 2> gpurun_out/bench_r2s_2gpu_$mode.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2s_2gpu_$mode.json"))
    print("$mode", d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],1), d["scaling"], d["config"]["windows_per_gpu"], d["config"]["windows_total"])
except Exception as e:
    print("$mode failed", e)
PY
  tail -2 gpurun_out/bench_r2s_2gpu_$mode.err
done
