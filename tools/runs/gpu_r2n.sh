#!/bin/bash
# Final single-GPU lines: default bench (cpu_baseline + e2e), config D, config C.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?" >> gpurun_out/bench_full.err
timeout 600 python bench.py --config D --no-cpu-baseline --no-e2e > gpurun_out/bench_D.json 2> gpurun_out/bench_D.err; echo "rc=$?" >> gpurun_out/bench_D.err
timeout 600 python tools/bench_config_c.py > gpurun_out/bench_C.json 2> gpurun_out/bench_C.err; echo "rc=$?" >> gpurun_out/bench_C.err
