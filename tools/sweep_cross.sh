#!/bin/bash
# Sweep of the streaming cross-attention kernel's ring depth / consumer warps / CTAs per SM (device us per launch).
for cfg in "6 8 1" "8 8 1" "10 8 1" "12 8 1" "6 16 1" "8 16 1" "12 16 1" "4 8 2" "6 8 2" "3 4 2" "4 4 2" "4 16 2" "3 8 2"; do
  set -- $cfg
  for R in 120 60; do
    echo -n "stages=$1 warps=$2 per_sm=$3 R=$R: "
    NOBS_WHISPER_CROSS_STAGES=$1 NOBS_WHISPER_CROSS_WARPS=$2 NOBS_WHISPER_CROSS_PER_SM=$3 python tools/time_decode_kernels.py $R 1280 | grep "stream"
  done
done
