#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_all.log
for i in 1 2; do timeout 300 python tools/gemm_fill_probe.py 2>&1 | head -n 1 >> gpurun_out/ld16.log; done
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"gat_attn_bwd2_kernel" -s 3 -c 1 --csv --log-file gpurun_out/ld16_dram.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-structured --no-graph > /dev/null 2>&1
