#!/bin/bash
O=gpurun_out; mkdir -p $O
T=r02b
timeout 900 compute-sanitizer --tool memcheck --log-file $O/${T}_sanitizer_memcheck.log python tools/sanitize_small.py > $O/${T}_sanitizer_memcheck.out 2>&1; echo "memcheck rc=$?"; tail -2 $O/${T}_sanitizer_memcheck.log; tail -4 $O/${T}_sanitizer_memcheck.out
timeout 900 compute-sanitizer --tool racecheck --log-file $O/${T}_sanitizer_racecheck.log python tools/sanitize_small.py 1 3 > $O/${T}_sanitizer_racecheck.out 2>&1; echo "racecheck rc=$?"; tail -2 $O/${T}_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --log-file $O/${T}_sanitizer_synccheck.log python tools/sanitize_small.py 3 > $O/${T}_sanitizer_synccheck.out 2>&1; echo "synccheck rc=$?"; tail -2 $O/${T}_sanitizer_synccheck.log
for Z in r ri; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --mcts-trees 0 --zero-copy $Z > $O/c12_bench_zc_$Z.json 2>/dev/null
python - <<PY
import json
d=json.load(open("$O/c12_bench_zc_$Z.json"))
print("zero-copy '$Z': e2e %.1f M (%.1f us/step) pipelined %.1f M" % (d["e2e"]["value"]/1e6, d["e2e"]["us_per_step"], d["e2e"]["pipelined_value"]/1e6))
PY
done
