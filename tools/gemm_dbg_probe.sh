#!/bin/bash
# GPU box, bring-up library (tools/ab/libspotv2_gat_bringup.so): forward projection time with parts of the pair epilogue
# switched off (SPOTV2_GEMM_DBG bits: 1 no TMA stores, 2 no scale/amax, 4 no per-chunk accumulation, 8 no pair epilogue)
export SPOTV2_GAT_LIB=$PWD/tools/ab/libspotv2_gat_bringup.so
for dbg in 0 1 8 9 15 0; do
  SPOTV2_GEMM_DBG=$dbg python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-structured --no-graph 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
p=d['phase_ms']
print('dbg $dbg', 'step %.3f' % d['ms_per_step'], ' '.join('%s %.3f' % (k, v) for k, v in p.items()))
"
done
