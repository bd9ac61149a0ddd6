#!/bin/bash
# GPU visit: full GPU test-suite + the e2e leg under each zero-copy mode.   usage (under gpurun): bash tools/gpu_zc.sh TAG
TAG=${1:-zc}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
for Z in "" r ri; do
  python bench.py --steps 600 --warmup 20 --e2e-steps 600 --no-cpu-baseline --mcts-trees 0 --zero-copy "$Z" 2>$O/${TAG}_bench_$Z.err | tail -1 > $O/${TAG}_bench_zc_$Z.json
  python -c "import json; d=json.load(open('$O/${TAG}_bench_zc_$Z.json')); print('zero-copy=[$Z]', 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), d['e2e']['path'][-60:])"
done
