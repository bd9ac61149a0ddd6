#!/bin/bash
# bench with the search engine on SPLIT; predict knobs
mkdir -p gpurun_out
for w in 16 20 22 24; do echo "== default lib, predict warps $w"; IPP_BULK_PREDICT_WARPS=$w timeout 300 python tools/predict_probe.py split 2>&1 | grep predict; done
for v in g2 g6 td4; do echo "== variant $v"; IPP_B200_LIB=build/variants/libipp_$v.so timeout 300 python tools/predict_probe.py split 2>&1 | grep -v layout; done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/s3_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/s3_bench.json').read().strip().splitlines()[-1])
print('value',l['value'],'frac',l['roofline']['frac'],'e2e',l['e2e']['value'], l['e2e']['us_per_step'], 'pipelined', l['e2e']['pipelined_value'])
m=l.get('mcts_rollouts'); print('mcts ms/sim', m['ms_per_lockstep_simulation'], m['layout'], m['tree_simulations_per_sec'])
p=l['roofline']['predict']; print('predict', p['value'], p['frac'], p['layout'], 'eval-only', p['evaluate_only']['value'], '| on step layout', p['on_step_layout']['value'])
PY
timeout 600 python bench.py --steps 20 --warmup 5 --search-layout planes --no-cpu-baseline > gpurun_out/s3_bench_planes.json 2> gpurun_out/s3_bench_planes.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/s3_bench_planes.json').read().strip().splitlines()[-1])
m=l.get('mcts_rollouts'); print('mcts ms/sim', m['ms_per_lockstep_simulation'], m['layout'])
p=l['roofline']['predict']; print('predict', p['value'], p['frac'], p['layout'])
PY
