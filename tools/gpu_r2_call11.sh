#!/bin/bash
O=gpurun_out; mkdir -p $O
T=c11
for L in planes mv tiled super; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 5 --mcts-layout $L > $O/${T}_bench_$L.json 2> $O/${T}_bench_$L.err; echo "bench $L rc=$?"; tail -2 $O/${T}_bench_$L.err
python - <<PY
import json
d=json.load(open("$O/${T}_bench_$L.json"))
print("$L", "mcts %.4f ms/sim %.0f M pred/s" % (d["mcts_rollouts"]["ms_per_lockstep_simulation"], d["mcts_rollouts"]["prediction_steps_per_sec"]/1e6))
PY
done
