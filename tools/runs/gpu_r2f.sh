#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "structured or large_universe" > gpurun_out/pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_new.log
timeout 900 python bench.py --config D --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_D.json 2> gpurun_out/bench_D.err; echo "rc=$?" >> gpurun_out/bench_D.err
