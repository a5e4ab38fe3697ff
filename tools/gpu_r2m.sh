#!/bin/bash
mkdir -p gpurun_out
for d in 0 8 0 8; do SPOTV2_GEMM_DBG=$d timeout 300 python tools/gemm_fill_probe.py > gpurun_out/probe_hint_$d.log 2>&1; cat gpurun_out/probe_hint_$d.log >> gpurun_out/probe_hint_all.log; done
