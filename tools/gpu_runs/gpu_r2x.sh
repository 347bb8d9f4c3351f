#!/bin/bash
# round 2 final evidence: default bench line, ncu --set full of the grouped cross-attention and the two-CTA encoder attention (kernel tests),
# ncu launch list of smoke()
mkdir -p gpurun_out
( time timeout 420 python bench.py > gpurun_out/bench_r2x.json ) 2> gpurun_out/bench_r2x.err
tail -3 gpurun_out/bench_r2x.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_r2x.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "iso", d["roofline"].get("isolated"))
    print("cpu", d["cpu_baseline"]); print("latency", d["latency"]); print("beam5", d.get("latency_beam5")); print("clocks", d["clocks"])
except Exception as e:
    print("bench failed", e)
PY
T=tests/test_gpu_kernels_bf16.py
timeout 100 ncu --set full --clock-control none --import-source on -k regex:"dec_cross_attention_tc" -c 3 -o gpurun_out/r2_ncu_full_cross_groups \
  python -m pytest "$T::test_decoder_cross_attention_kernel[10-6-2-1500-25]" "$T::test_decoder_cross_attention_kernel[121-20-31-1500-22]" "$T::test_decoder_cross_attention_kernel[130-6-4-1500-2]" -q > gpurun_out/ncu_full_r2x_cross.log 2>&1
tail -2 gpurun_out/ncu_full_r2x_cross.log
ncu -i gpurun_out/r2_ncu_full_cross_groups.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_cross_groups_raw.csv 2>/dev/null
timeout 60 ncu --set full --clock-control none --import-source on -k regex:"enc_attention_sm100" -c 1 -o gpurun_out/r2_ncu_full_enc_attention \
  python -m pytest "$T::test_encoder_attention_kernel[2-3-0]" -q > gpurun_out/ncu_full_r2x_att.log 2>&1
tail -2 gpurun_out/ncu_full_r2x_att.log
ncu -i gpurun_out/r2_ncu_full_enc_attention.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_enc_attention_raw.csv 2>/dev/null
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r2_launches_smoke.csv \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_launches_r2x.log 2>&1
tail -2 gpurun_out/ncu_launches_r2x.log
ls -la gpurun_out/r2_ncu_full_* gpurun_out/r2_launches_smoke.csv | head
