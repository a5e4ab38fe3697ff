#!/bin/bash
# round 1o: config-D timing through bench.py, then the config-A bench line with both host-buffer legs
mkdir -p gpurun_out
timeout 900 python bench.py --config D --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_D.json 2> gpurun_out/bench_D.err; echo "rc=$?" >> gpurun_out/bench_D.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_D.csv python bench.py --config D --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_D.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
