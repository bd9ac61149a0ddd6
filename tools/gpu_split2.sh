#!/bin/bash
# memoised MCTS rollout + predict-warps variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_missions.py tests/test_gpu_full_size_parity.py -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/s2_pytest.log
for v in pw28 pw32; do
  for w in 24 28 32; do
    echo "== variant $v warps $w"; IPP_B200_LIB=build/variants/libipp_$v.so IPP_BULK_PREDICT_WARPS=$w timeout 300 python tools/predict_probe.py split 2>&1 | grep predict
  done
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/s2_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/s2_bench.json').read().strip().splitlines()[-1])
print('value',l['value'],'e2e',l['e2e']['value'])
print(json.dumps(l.get('mcts_rollouts'),indent=1))
print(json.dumps(l['roofline'].get('predict'),indent=1))
PY
