#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_all.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-structured --no-graph > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gat_attn_bwd2_kernel -s 3 -c 1 -o gpurun_out/bwd2_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-structured --no-graph > gpurun_out/ncu_bwd2.log 2>&1
