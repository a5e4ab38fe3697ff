#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -p no:cacheprovider -k "dropout or large_universe or config_c or projection_gemms" > gpurun_out/pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_new.log
timeout 600 python tools/bench_config_c.py > gpurun_out/bench_C.json 2> gpurun_out/bench_C.err; echo "rc=$?" >> gpurun_out/bench_C.err
timeout 900 python bench.py --config D --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_D.json 2> gpurun_out/bench_D.err; echo "rc=$?" >> gpurun_out/bench_D.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_D.csv python bench.py --config D --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_D.log 2>&1
