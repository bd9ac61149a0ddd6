#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mcts_select -s 150 -c 1 -f -o $O/p2_select python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-sims 100 > $O/p2_select.log 2>&1; echo "ncu select rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mcts_expand -s 150 -c 1 -f -o $O/p2_expand python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-sims 100 > $O/p2_expand.log 2>&1; echo "ncu expand rc=$?"
ls -la $O/p2_*
