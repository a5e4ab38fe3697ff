#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/s.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch, bench
bench.CFG["N"] = 30
hp = bench.HotPath(4096, torch.device("cuda", 0), 1234, structured=True)
for _ in range(3): hp.step()
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none -k regex:"win_" -s 4 -c 2 -o gpurun_out/win_full -f python /tmp/s.py > gpurun_out/ncu_win.log 2>&1
