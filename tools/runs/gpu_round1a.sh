#!/bin/bash
# First GPU pass: parity tests, smoke, bench, tensor-core rounding probe, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_full_batch_properties > gpurun_out/pytest_gpu_all.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python tools/tc_probe.py > gpurun_out/tc_probe.txt 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gat_attn -s 3 -c 2 -o gpurun_out/attn_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
