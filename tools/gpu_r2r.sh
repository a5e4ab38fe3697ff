#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/stagger.log
for ns in 0 4000 8000 16000 0 8000; do echo "STAGGER=$ns" >> gpurun_out/stagger.log; SPOTV2_BWD_STAGGER_NS=$ns timeout 300 python tools/gemm_fill_probe.py 2>&1 | head -n 1 >> gpurun_out/stagger.log; done
