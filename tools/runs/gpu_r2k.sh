#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_all.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
