#!/bin/bash
# round 1t: full GPU suite, smoke, bench line, config-D and config-C timings, then the profile set (A and D)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py --config D --steps 5 --warmup 3 > gpurun_out/bench_D.json 2> gpurun_out/bench_D.err; echo "rc=$?" >> gpurun_out/bench_D.err
timeout 600 python tools/bench_config_c.py > gpurun_out/bench_C.json 2> gpurun_out/bench_C.err
bash tools/gpu_profiles.sh > gpurun_out/profiles.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_D.csv python bench.py --config D --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_D.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lg_edge_logit|lg_dv|bgemm|lg_softmax" -s 9 -c 9 -o gpurun_out/large_full -f python bench.py --config D --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_large.log 2>&1
