#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "zero_copy or pipelined or edge" > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s5_pytest.log
timeout 300 python tools/e2e_breakdown.py 2>&1 | tee gpurun_out/s5_e2e_breakdown.txt
IPP_ZERO_COPY=r timeout 300 python tools/e2e_breakdown.py 2>&1 | head -1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --mcts-trees 0 > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s5_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/s5_bench.json').read().strip().splitlines()[-1])
print('value',l['value'],'frac',l['roofline']['frac'],'e2e',l['e2e']['value'], l['e2e']['us_per_step'], 'pipelined', l['e2e']['pipelined_value'])
print(l['e2e']['path'])
p=l['roofline']['predict']; print('predict', p['value'], p['frac'], p['layout'], 'eval-only', p['evaluate_only']['value'])
PY
