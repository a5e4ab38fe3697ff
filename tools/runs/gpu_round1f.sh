#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/fwd_waits.py > gpurun_out/fwd_waits.txt 2>&1
CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "B4N7F9Fe5H3C5mean" > gpurun_out/pytest_odd.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_all.log
ls -la gpurun_out
