#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipp_rollout -s 20 -c 1 -f -o $O/c15_rollout \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-sims 24 > $O/c15_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mcts_select -s 20 -c 1 -f -o $O/c15_select \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-sims 24 >> $O/c15_ncu.log 2>&1
ls -la $O/c15_*.ncu-rep
