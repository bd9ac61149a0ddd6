#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python tools/debug_layout3.py > $O/c2_debug.log 2>&1; echo "debug rc=$?"; tail -40 $O/c2_debug.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipp_step_bulk -s 6 -c 1 -f -o $O/c2_bulk \
  python bench.py --steps 8 --warmup 3 --batch 65536 --no-cpu-baseline --e2e-steps 2 --mcts-trees 0 --layout super > $O/c2_ncu_bench.log 2>&1
ls -la $O/c2_bulk.ncu-rep
timeout 600 python -m pytest tests/test_gpu_full_size_parity.py -x -q > $O/c2_full.log 2>&1; echo "full rc=$?"; tail -12 $O/c2_full.log
