#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths_and_scale.py -m gpu -x -q > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s7_pytest.log
echo "== poll"; timeout 300 python tools/e2e_breakdown.py 2>&1 | head -1
echo "== no poll"; IPP_POLL_DONE=0 timeout 300 python tools/e2e_breakdown.py 2>&1 | head -1
echo "== no poll no fetch"; IPP_POLL_DONE=0 IPP_ZERO_COPY=r timeout 300 python tools/e2e_breakdown.py 2>&1 | head -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s7_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/s7_bench.json').read().strip().splitlines()[-1])
print('value',l['value'],'frac',l['roofline']['frac'],'e2e',l['e2e']['value'], l['e2e']['us_per_step'], 'pipelined', l['e2e']['pipelined_value'])
m=l.get('mcts_rollouts'); print('mcts ms/sim', m['ms_per_lockstep_simulation'], m['layout'], m['tree_simulations_per_sec'])
p=l['roofline']['predict']; print('predict', p['value'], p['frac'], p['layout'], 'eval-only', p['evaluate_only']['value'], '| on step layout', p['on_step_layout']['value'])
print(l['cpu_baseline']['value'], l['clocks'])
PY
