#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python tools/predict_probe.py > $O/c9_predict_probe.log 2>&1; cat $O/c9_predict_probe.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipp_step_bulk -s 16 -c 1 -f -o $O/c9_predict \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-trees 0 > $O/c9_ncu_predict.log 2>&1
ls -la $O/c9_predict.ncu-rep
