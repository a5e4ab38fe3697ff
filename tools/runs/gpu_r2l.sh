#!/bin/bash
# TMA-store epilogue check: GEMM-facing parity tests, then the projection phase probe.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "projection_gemms or f16_pair or layer_matches or gemm" > gpurun_out/pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_new.log
timeout 300 python tools/gemm_fill_probe.py > gpurun_out/probe_tma.log 2>&1; echo "rc=$?" >> gpurun_out/probe_tma.log
SPOTV2_GEMM_DBG=1 timeout 300 python tools/gemm_fill_probe.py > gpurun_out/probe_tma_dbg1.log 2>&1
