#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gat_attn_bwd2_kernel -s 3 -c 1 -o gpurun_out/bwd2_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bwd2.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_bwd2.log
