#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "structured" > gpurun_out/pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_new.log
timeout 900 python bench.py --config D --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_D.json 2> gpurun_out/bench_D.err; echo "rc=$?" >> gpurun_out/bench_D.err
cat > /tmp/s.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch, bench
bench.CFG["N"] = 500
hp = bench.HotPath(32, torch.device("cuda", 0), 1234, structured=True)
for _ in range(2): hp.step()
torch.cuda.synchronize()
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"win_" -c 40 --csv --log-file gpurun_out/launches_win.csv python /tmp/s.py > gpurun_out/ncu_win.log 2>&1
