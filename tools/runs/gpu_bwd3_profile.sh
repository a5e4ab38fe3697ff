#!/bin/bash
# GPU box: role counters of the tcgen05 attention backward, then one ncu --set full capture of it
python tools/bwd3_waits.py > gpurun_out/bwd3_waits.txt 2>&1
python tools/bwd3_waits.py --structured >> gpurun_out/bwd3_waits.txt 2>&1
cat gpurun_out/bwd3_waits.txt
ncu --set full --clock-control none --import-source on -k regex:gat_attn_bwd3 -s 3 -c 1 -o gpurun_out/bwd3 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --no-structured > gpurun_out/bwd3_ncu.log 2>&1
tail -3 gpurun_out/bwd3_ncu.log
