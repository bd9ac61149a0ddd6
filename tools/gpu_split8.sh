#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_parity.py -m gpu -x -q -k "memoised or zero_copy or search" > gpurun_out/s9_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s9_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s9_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/s9_bench.json').read().strip().splitlines()[-1])
print('value',l['value'],'frac',l['roofline']['frac'],'e2e',l['e2e']['value'], l['e2e']['us_per_step'], 'pipelined', l['e2e']['pipelined_value'])
m=l.get('mcts_rollouts'); print('mcts ms/sim', m['ms_per_lockstep_simulation'], m['simulations'], m['tree_simulations_per_sec'], m['prediction_steps_per_sec'], m['tree_bytes_per_gpu'])
p=l['roofline']['predict']; print('predict', p['value'], p['frac'], p['layout'], 'eval-only', p['evaluate_only']['value'], '| on step layout', p['on_step_layout']['value'])
PY
