#!/bin/bash
# One full ncu capture of the step kernel (default layout).  usage (under gpurun): bash tools/gpu_ncu_step.sh TAG [extra bench args]
TAG=${1:-ncu}; shift
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipp_step_async -s 6 -c 1 -f -o $O/${TAG}_async \
  python bench.py --steps 8 --warmup 3 --batch 65536 --no-cpu-baseline --e2e-steps 2 --mcts-trees 0 "$@" > $O/${TAG}_ncu_bench.log 2>&1
ls -la $O/${TAG}_async.ncu-rep
