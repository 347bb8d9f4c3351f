#!/bin/bash
# round 2 evidence run: full GPU test suite, the headline bench line, ncu launch list + full sections of the decode kernels
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -q -m gpu ) > gpurun_out/pytest_gpu_r2p.log 2>&1
tail -6 gpurun_out/pytest_gpu_r2p.log
( time timeout 900 python bench.py > gpurun_out/bench_r2p.json ) 2> gpurun_out/bench_r2p.err
tail -3 gpurun_out/bench_r2p.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2p.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "iso", d["roofline"].get("isolated"))
print("cpu", d["cpu_baseline"]); print("latency", d["latency"]); print("clocks", d["clocks"])
PY
# launch list of the bench command at 16 windows (ends in minutes under ncu)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 2500 --csv --log-file gpurun_out/r2_launches_16win.csv \
  python bench.py --windows 16 --steps 1 --warmup 0 --no-cpu-baseline --latency-clips 0 > gpurun_out/ncu_launches_r2p.log 2>&1
tail -2 gpurun_out/ncu_launches_r2p.log
# full sections: cross-attention, the fused cluster projection (FC1), the split-K step GEMM and its epilogue, self-attention
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dec_cross_attention_tc|dec_proj_cluster|gemm_skinny_sm100|skinny_reduce|dec_self_attention" -s 600 -c 12 \
  -o gpurun_out/r2_ncu_full_decode python bench.py --windows 16 --steps 1 --warmup 0 --no-cpu-baseline --latency-clips 0 > gpurun_out/ncu_full_r2p.log 2>&1
tail -2 gpurun_out/ncu_full_r2p.log
ncu -i gpurun_out/r2_ncu_full_decode.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_decode_raw.csv 2>/dev/null
ls -la gpurun_out/r2_ncu_full_decode* | head
