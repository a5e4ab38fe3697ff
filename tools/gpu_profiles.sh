#!/bin/bash
# Round-1 profile set: plain bench first (must exit 0), then the ncu launch list of the same command and one
# --set full capture each of the attention forward / backward and the two projection GEMM kernels.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-structured --no-graph > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-structured --no-graph > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gat_attn_fwd -s 3 -c 1 -o gpurun_out/fwd_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-structured --no-graph > gpurun_out/ncu_fwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gat_attn_bwd2_kernel -s 3 -c 1 -o gpurun_out/bwd2_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-structured --no-graph > gpurun_out/ncu_bwd2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:gemm3x_f16_kernel -s 6 -c 2 -o gpurun_out/gemm_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-structured --no-graph > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
