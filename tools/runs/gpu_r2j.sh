#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/s.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch, bench
N = int(sys.argv[1]); B = int(sys.argv[2])
bench.CFG["N"] = N
hp = bench.HotPath(B, torch.device("cuda", 0), 1234, structured=True)
for _ in range(3): hp.step()
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none -k regex:"win_|lg_dterms" -s 8 -c 8 -o gpurun_out/win_D_full -f python /tmp/s.py 500 32 > gpurun_out/ncu_winD.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"win_" -s 4 -c 2 -o gpurun_out/win_A_full -f python /tmp/s.py 30 4096 > gpurun_out/ncu_winA.log 2>&1
