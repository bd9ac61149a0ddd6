#!/bin/bash
O=gpurun_out; mkdir -p $O
bash tools/gpu_sweep.sh c8 default td2 td4 g2 g4 ch4 eg1 eg3 w15 w14r default
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipp_step_bulk -s 60 -c 1 -f -o $O/c8_predict \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-trees 0 > $O/c8_ncu_predict.log 2>&1
ls -la $O/c8_predict.ncu-rep
