#!/bin/bash
# MCTS leg: tests, bench numbers, per-kernel times at the C4 per-GPU size.   usage (under gpurun): bash tools/gpu_mcts.sh TAG
TAG=${1:-mcts}; O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_mcts.py -x -q 2>&1 | tail -3
python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 20 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['mcts_rollouts']))"
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"mcts|rollout" -c 60 --csv \
  --log-file $O/${TAG}_kernels.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-sims 10 > /dev/null 2>&1
echo done
