#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_missions.py -m gpu -x -q > gpurun_out/m2_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/m2_pytest.log
timeout 600 python -m pytest tests/test_gpu_full_size_parity.py -m gpu -x -q -k "search or mcts or tree" 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 10 > gpurun_out/m2_bench.json 2> gpurun_out/m2_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/m2_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/m2_bench.json').read().strip().splitlines()[-1])
m=l.get('mcts_rollouts'); print('mcts ms/sim', m['ms_per_lockstep_simulation'], m['simulations'], m['tree_simulations_per_sec'], m['gpu_launches'])
PY
