#!/bin/bash
O=gpurun_out; mkdir -p $O
bash tools/gpu_sweep.sh c13 default st1 st2 st3 ld1 ld2 ld1st1 default
timeout 600 python tools/mcts_gap_probe.py > $O/c13_mcts_gap.log 2>&1; cat $O/c13_mcts_gap.log | tail -5
