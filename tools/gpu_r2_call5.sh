#!/bin/bash
O=gpurun_out; mkdir -p $O
T=c5
timeout 900 python -m pytest tests/test_gpu_paths_and_scale.py tests/test_gpu_parity.py tests/test_gpu_mcts.py -x -q > $O/${T}_paths.log 2>&1; echo "paths+parity+mcts rc=$?"; tail -6 $O/${T}_paths.log
timeout 900 python -m pytest tests/test_gpu_full_size_parity.py -x -q > $O/${T}_full.log 2>&1; echo "full rc=$?"; tail -8 $O/${T}_full.log
for L in super tiled; do
timeout 300 python bench.py --layout $L --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 100 > $O/${T}_bench_$L.json 2> $O/${T}_bench_$L.err; echo "bench $L rc=$?"; python - <<PY
import json
d=json.load(open("$O/${T}_bench_$L.json"))
print("$L", d["value"]/1e6, d["value_trace_reduction"]/1e6, d["roofline"]["frac"], "e2e", d["e2e"]["value"]/1e6, d["e2e"]["pipelined_value"]/1e6, "mcts ms/sim", d["mcts_rollouts"]["ms_per_lockstep_simulation"], d["mcts_rollouts"]["prediction_steps_per_sec"]/1e6)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"mcts_|rollout" -c 120 --csv --log-file $O/${T}_mcts_launches.csv \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-sims 24 --layout super > /dev/null 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open("$O/${T}_mcts_launches.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
H=rows[hdr]; kn=H.index("Kernel Name"); mn=H.index("Metric Name"); mv=H.index("Metric Value")
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[hdr+1:]:
    if len(r)>mv: agg[r[kn][:40]][r[mn]].append(float(r[mv].replace(",","")))
for k,v in agg.items():
    print(k, {m:(len(x), sum(x)/len(x), max(x)) for m,x in v.items()})
PY
