#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_all.log
timeout 300 python tools/gemm_fill_probe.py > gpurun_out/probe_bwd.log 2>&1
timeout 300 python tools/fwd_waits.py > gpurun_out/waits.log 2>&1
