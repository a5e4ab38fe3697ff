#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lg_dv_kernel -s 1 -c 1 -o gpurun_out/lg_dv python bench.py --config D --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_lg_dv.log 2>&1
