#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -q -m gpu -x ) > gpurun_out/pytest_gpu_r2r.log 2>&1
tail -8 gpurun_out/pytest_gpu_r2r.log
