#!/bin/bash
# round 1n: large-universe (N > 32) path bring-up + regression of the whole GPU suite after the ABI v3 change
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "large_universe" > gpurun_out/pytest_large.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_large.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not large_universe" > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_all.log
