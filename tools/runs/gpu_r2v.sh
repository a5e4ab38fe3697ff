#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config D --no-cpu-baseline --no-e2e > gpurun_out/bench_D.json 2> gpurun_out/bench_D.err; echo "rc=$?" >> gpurun_out/bench_D.err
timeout 600 python tools/bench_config_c.py > gpurun_out/bench_C.json 2> gpurun_out/bench_C.err; echo "rc=$?" >> gpurun_out/bench_C.err
