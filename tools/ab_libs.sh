#!/bin/bash
# GPU box: A/B two builds of the library on the same box (base = tools/ab/libspotv2_gat_base.so, new = the in-tree one):
# phase times of the default bench, alternating, so that box-to-box differences cancel.
# Base build: git stash; python -m spotv2net_b200.build; mkdir -p tools/ab; cp spotv2net_b200/libspotv2_gat.so tools/ab/libspotv2_gat_base.so;
# git stash pop; python -m spotv2net_b200.build   (tools/ab/*.so is git-ignored and travels with the gpurun snapshot; delete it afterwards)
for rep in 1 2 3; do
  for which in base new; do
    if [ $which = base ]; then export SPOTV2_GAT_LIB=$PWD/tools/ab/libspotv2_gat_base.so; else unset SPOTV2_GAT_LIB; fi
    python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-structured --no-graph 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
p=d['phase_ms']
print('$which', 'step %.3f' % d['ms_per_step'], ' '.join('%s %.3f' % (k, v) for k, v in p.items()))
"
  done
done
