#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "attention_stages or layer_matches or golden or structured" > gpurun_out/pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_new.log
timeout 300 python tools/fwd_waits.py > gpurun_out/fwd_waits.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
