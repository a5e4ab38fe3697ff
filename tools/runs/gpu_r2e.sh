#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-structured > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "rc=$?" >> gpurun_out/bench_graph.err
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-structured --no-graph > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; echo "rc=$?" >> gpurun_out/bench_eager.err
