#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/promo.log
for pr in 3 0 2 3 0; do echo "PROMO=$pr" >> gpurun_out/promo.log; SPOTV2_TMAP_PROMO=$pr timeout 300 python tools/gemm_fill_probe.py 2>&1 | head -n 1 >> gpurun_out/promo.log; done
for pr in 0 3; do
SPOTV2_TMAP_PROMO=$pr timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"gat_attn_bwd2_kernel|gat_attn_fwd_kernel" -s 6 -c 2 --csv --log-file gpurun_out/promo_dram_$pr.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-structured --no-graph > /dev/null 2>&1
done
