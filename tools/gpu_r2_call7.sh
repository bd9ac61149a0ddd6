#!/bin/bash
O=gpurun_out; mkdir -p $O
T=c7
timeout 900 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_full_size_parity.py -x -q > $O/${T}_mcts.log 2>&1; echo "mcts+full rc=$?"; tail -6 $O/${T}_mcts.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${T}_bench_k20.json 2> $O/${T}_bench_k20.err; echo "bench k20 rc=$?"; tail -3 $O/${T}_bench_k20.err
timeout 600 python bench.py --steps 200 --warmup 10 > $O/${T}_bench_k200.json 2> $O/${T}_bench_k200.err; echo "bench k200 rc=$?"; tail -3 $O/${T}_bench_k200.err
python - <<PY
import json
for k in ("k20","k200"):
    try:
        d=json.load(open("$O/${T}_bench_%s.json"%k))
    except Exception as e:
        print(k, "no json", e); continue
    r=d["roofline"]
    print(k, "value %.1f M frac %.3f | entropy %.1f M | predict %.1f M frac %.3f (%s) | e2e %.1f M (%.0f us) pipelined %.1f M | mcts %.4f ms/sim %.0f M pred/s | clocks %s" % (
        d["value"]/1e6, r["frac"], r["modes"]["gauss_entropy"]["value"]/1e6, r["predict"]["value"]/1e6, r["predict"]["frac"], r["predict"]["kernel"],
        d["e2e"]["value"]/1e6, d["e2e"]["us_per_step"], d["e2e"]["pipelined_value"]/1e6, d["mcts_rollouts"]["ms_per_lockstep_simulation"], d["mcts_rollouts"]["prediction_steps_per_sec"]/1e6, d["clocks"]))
    print("   cpu", d.get("cpu_baseline"))
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
