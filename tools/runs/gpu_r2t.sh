#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_all.log
rm -f gpurun_out/splits.log
for i in 1 2 3; do timeout 300 python tools/gemm_fill_probe.py 2>&1 | head -n 1 >> gpurun_out/splits.log; done
