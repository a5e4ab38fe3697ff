#!/bin/bash
# Check + profile set: GPU tests, the three bench lines, the ncu launch list and --set full captures (tools/gpu_profiles.sh),
# and - when a bring-up build of the library was shipped as tools/ab/libspotv2_gat_bringup.so - the kernels' cycle counters.
T=${1:-r2z}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; tail -2 gpurun_out/${T}_tests.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --config C --steps 5 > gpurun_out/${T}_bench_C.json 2> gpurun_out/${T}_bench_C.err; echo "bench C rc=$?"
timeout 600 python bench.py --config D --steps 5 > gpurun_out/${T}_bench_D.json 2> gpurun_out/${T}_bench_D.err; echo "bench D rc=$?"
bash tools/gpu_profiles.sh 2>&1 | tail -6
if [ -f tools/ab/libspotv2_gat_bringup.so ]; then
  SPOTV2_GAT_LIB=$PWD/tools/ab/libspotv2_gat_bringup.so timeout 300 python tools/fwd_waits.py > gpurun_out/${T}_counters.txt 2>&1; tail -8 gpurun_out/${T}_counters.txt
fi
