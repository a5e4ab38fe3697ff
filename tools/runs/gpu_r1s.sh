#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -p no:cacheprovider -k "config_c" > gpurun_out/pytest_c.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_c.log
