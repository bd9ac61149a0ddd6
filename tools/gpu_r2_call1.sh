#!/bin/bash
# round 2, call 1: first run of the bulk-copy kernel (IPP_LAYOUT_SUPER): smoke, parity, bench A/B vs the tiled layout, sanitizer
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/c1_smoke.log 2>&1; echo "smoke rc=$?"; tail -6 $O/c1_smoke.log
timeout 900 python -m pytest tests/test_gpu_paths_and_scale.py -x -q > $O/c1_paths.log 2>&1; echo "paths rc=$?"; tail -6 $O/c1_paths.log
for L in super tiled; do
  timeout 300 python bench.py --layout $L --steps 200 --warmup 10 --no-cpu-baseline --mcts-trees 0 --e2e-steps 100 > $O/c1_bench_$L.json 2> $O/c1_bench_$L.err; echo "bench $L rc=$?"; cut -c1-900 $O/c1_bench_$L.json
done
timeout 1200 python -m pytest tests/test_gpu_full_size_parity.py -x -q > $O/c1_full.log 2>&1; echo "full rc=$?"; tail -12 $O/c1_full.log
timeout 600 compute-sanitizer --tool memcheck --log-file $O/c1_memcheck.log python tools/sanitize_small.py > $O/c1_memcheck.out 2>&1; echo "memcheck rc=$?"; tail -3 $O/c1_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --log-file $O/c1_racecheck.log python tools/sanitize_small.py 3 > $O/c1_racecheck.out 2>&1; echo "racecheck rc=$?"; tail -3 $O/c1_racecheck.log
