#!/bin/bash
# usage: [BENCH_ARGS="..."] tools/run_ablation.sh VARIANT...   (libs in build/exp/lib_<VARIANT>.so)
for V in "$@"; do
  IPP_B200_LIB=$PWD/build/exp/lib_$V.so timeout 200 python -O bench.py --steps 300 --warmup 10 --no-cpu-baseline --e2e-steps 20 --mcts-trees 0 $BENCH_ARGS 2>/dev/null | tail -1 > /tmp/abl.json
  python -c "import json; d=json.load(open('/tmp/abl.json')); print('$V', round(d['value']/1e6,1), round(d['value_trace_reduction']/1e6,1), round(d['e2e']['value']/1e6,1), 'frac', round(d['roofline']['frac'],3))"
done
