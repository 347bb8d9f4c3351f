#!/bin/bash
# tcgen05 streaming cross-attention: ring depth / TMEM pitch / CTAs per SM ("stages spacing per_sm")
for cfg in "$@"; do
  set -- $cfg
  echo "== stages=$1 spacing=$2 per_sm=$3"
  export NOBS_WHISPER_CROSS_STAGES=$1 NOBS_WHISPER_CROSS_SPACING=$2 NOBS_WHISPER_CROSS_PER_SM=$3
  timeout 200 python -m pytest tests/test_gpu_kernels_bf16.py -x -q -k "cross and 2]" 2>&1 | tail -1
  for R in 120 60 40; do echo -n "R=$R "; timeout 100 python tools/time_decode_kernels.py $R 1280 | grep "tcgen05"; done
done
