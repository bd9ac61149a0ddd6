#!/bin/bash
O=gpurun_out; mkdir -p $O
T=c10
timeout 900 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_full_size_parity.py -x -q > $O/${T}_mcts.log 2>&1; echo "mcts+full rc=$?"; tail -4 $O/${T}_mcts.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; tail -3 $O/${T}_bench.err
python - <<PY
import json
d=json.load(open("$O/${T}_bench.json")); r=d["roofline"]
print("value %.1f M frac %.3f | predict %.1f M / eval-only %.1f M | e2e %.1f M (%.0f us) pipelined %.1f M | mcts %.4f ms/sim %.0f M pred/s" % (
    d["value"]/1e6, r["frac"], r["predict"]["value"]/1e6, r["predict"]["evaluate_only"]["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["us_per_step"], d["e2e"]["pipelined_value"]/1e6,
    d["mcts_rollouts"]["ms_per_lockstep_simulation"], d["mcts_rollouts"]["prediction_steps_per_sec"]/1e6))
PY
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"mcts_|rollout" -c 120 --csv --log-file $O/${T}_mcts_launches.csv \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-sims 24 > /dev/null 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open("$O/${T}_mcts_launches.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
H=rows[hdr]; kn=H.index("Kernel Name"); mn=H.index("Metric Name"); mv=H.index("Metric Value")
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[hdr+1:]:
    if len(r)>mv: agg[r[kn][:40]][r[mn]].append(float(r[mv].replace(",","")))
for k,v in agg.items():
    print(k, {m:(len(x), round(sum(x)/len(x)), max(x)) for m,x in v.items()})
PY
