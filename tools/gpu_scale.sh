#!/bin/bash
# Weak-scaling lines of bench.py on one multi-GPU box.   usage (under gpurun --gpus N): bash tools/gpu_scale.sh TAG N1 [N2 ...]
TAG=$1; shift
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 300 python bench.py --gpus 1 --steps 1000 --warmup 20 --e2e-steps 300 --mcts-trees 0 --no-cpu-baseline > $O/${TAG}_n$N.json 2> $O/${TAG}_n$N.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N \
      --steps 1000 --warmup 20 --e2e-steps 300 --mcts-trees 0 > $O/${TAG}_n$N.json 2> $O/${TAG}_n$N.err
  fi
  echo "N=$N rc=$?"; tail -1 $O/${TAG}_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'ms', round(d['ms_per_step'],4), d['clocks'])" || tail -5 $O/${TAG}_n$N.err
done
