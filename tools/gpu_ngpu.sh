#!/bin/bash
# the driver's own multi-GPU command line:  bash tools/gpu_ngpu.sh N   (under gpurun --gpus N)
N=$1; O=gpurun_out; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + N)) bench.py --gpus $N --steps 20 --warmup 5 \
   > $O/n${N}_bench.json 2> $O/n${N}_bench.err; echo "bench N=$N rc=$?"
tail -3 $O/n${N}_bench.err
python - <<PY
import json
l=json.loads(open('$O/n${N}_bench.json').read().strip().splitlines()[-1])
print('N', l['n_gpus'], 'value', round(l['value']/1e6,1), 'M | e2e', round(l['e2e']['value']/1e6,1), 'M', round(l['e2e']['us_per_step'],1), 'us | pipelined', round(l['e2e']['pipelined_value']/1e6,1))
m=l.get('mcts_rollouts'); print('mcts ms/sim', m['ms_per_lockstep_simulation'], m['tree_simulations_per_sec'])
p=l['roofline']['predict']; print('predict', p['value'], p['frac'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800 + N)) bench.py --impl reference --gpus $N --steps 3 --warmup 3 > $O/n${N}_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-250 $O/n${N}_ref.json
