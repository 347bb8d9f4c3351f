#!/bin/bash
# full GPU test suite on the round's final code
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_r2w.log 2>&1
tail -8 gpurun_out/pytest_gpu_r2w.log
