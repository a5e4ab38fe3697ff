#!/bin/bash
# Round-2 check + profile set: GPU tests, the three bench lines, then the ncu launch list and --set full captures
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_tests.log 2>&1; tail -2 gpurun_out/r2z_tests.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --config C --steps 5 > gpurun_out/r2z_bench_C.json 2> gpurun_out/r2z_bench_C.err; echo "bench C rc=$?"
timeout 600 python bench.py --config D --steps 5 > gpurun_out/r2z_bench_D.json 2> gpurun_out/r2z_bench_D.err; echo "bench D rc=$?"
bash tools/gpu_profiles.sh 2>&1 | tail -6
