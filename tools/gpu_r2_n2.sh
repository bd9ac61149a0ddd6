#!/bin/bash
O=gpurun_out; mkdir -p $O
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > $O/n${N}_bench.json 2> $O/n${N}_bench.err; echo "bench N=$N rc=$?"; tail -5 $O/n${N}_bench.err
python - <<PY
import json
d=json.loads(open("$O/n${N}_bench.json").read().strip().splitlines()[-1])
print("N=%d value %.1f M | e2e %.1f M (%.0f us) pipelined %.1f M | mcts %.4f ms/sim | path: %s" % (d["n_gpus"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["us_per_step"], d["e2e"]["pipelined_value"]/1e6,
      d["mcts_rollouts"]["ms_per_lockstep_simulation"], d["e2e"]["path"][-120:]))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 3 --warmup 3 > $O/n${N}_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-300 $O/n${N}_ref.json
