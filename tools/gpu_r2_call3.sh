#!/bin/bash
O=gpurun_out; mkdir -p $O
T=${1:-c3}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/${T}_smoke.log
timeout 900 python -m pytest tests/test_gpu_paths_and_scale.py tests/test_gpu_parity.py -x -q > $O/${T}_paths.log 2>&1; echo "paths+parity rc=$?"; tail -6 $O/${T}_paths.log
timeout 300 python bench.py --layout super --steps 200 --warmup 10 --no-cpu-baseline --mcts-trees 0 --e2e-steps 100 > $O/${T}_bench_super.json 2> $O/${T}_bench_super.err; echo "bench rc=$?"; cut -c1-400 $O/${T}_bench_super.json
timeout 900 python -m pytest tests/test_gpu_full_size_parity.py -x -q > $O/${T}_full.log 2>&1; echo "full rc=$?"; tail -8 $O/${T}_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipp_step_bulk -s 6 -c 1 -f -o $O/${T}_bulk \
  python bench.py --steps 8 --warmup 3 --batch 65536 --no-cpu-baseline --e2e-steps 2 --mcts-trees 0 --layout super > $O/${T}_ncu_bench.log 2>&1
ls -la $O/${T}_bulk.ncu-rep
