#!/bin/bash
# GPU box: ncu --set full captures of the p_format 1 attention kernels (forward, backward), materialised edge rows
ncu --set full --clock-control none --import-source on -k regex:gat_attn_fwd16 -s 3 -c 1 -o gpurun_out/r2m_fwd16 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --no-structured > gpurun_out/r2m_ncu_fwd.log 2>&1
tail -2 gpurun_out/r2m_ncu_fwd.log
ncu --set full --clock-control none --import-source on -k regex:gat_attn_bwd2 -s 3 -c 1 -o gpurun_out/r2m_bwd2p -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --no-structured > gpurun_out/r2m_ncu_bwd.log 2>&1
tail -2 gpurun_out/r2m_ncu_bwd.log
