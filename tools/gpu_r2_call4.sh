#!/bin/bash
O=gpurun_out; mkdir -p $O
T=c4
timeout 900 python -m pytest tests/test_gpu_paths_and_scale.py tests/test_gpu_parity.py -x -q > $O/${T}_paths.log 2>&1; echo "paths+parity rc=$?"; tail -6 $O/${T}_paths.log
timeout 900 python -m pytest tests/test_gpu_full_size_parity.py -x -q > $O/${T}_full.log 2>&1; echo "full rc=$?"; tail -8 $O/${T}_full.log
timeout 900 python -m pytest tests/test_gpu_missions.py -q > $O/${T}_missions.log 2>&1; echo "missions rc=$?"; tail -25 $O/${T}_missions.log
for V in default w14 w12; do
  if [ $V = default ]; then unset IPP_B200_LIB; else export IPP_B200_LIB=$PWD/build/libipp_b200_$V.so; fi
  timeout 300 python bench.py --layout super --steps 200 --warmup 10 --no-cpu-baseline --mcts-trees 0 --e2e-steps 100 > $O/${T}_bench_$V.json 2> $O/${T}_bench_$V.err; echo "bench $V rc=$?"; cut -c1-160 $O/${T}_bench_$V.json
  timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:ipp_step_bulk -s 6 -c 2 --csv --log-file $O/${T}_inst_$V.csv \
     python bench.py --steps 8 --warmup 3 --batch 65536 --no-cpu-baseline --e2e-steps 2 --mcts-trees 0 --layout super > /dev/null 2>&1
  tail -6 $O/${T}_inst_$V.csv | cut -d, -f5,12-20
done
unset IPP_B200_LIB
timeout 300 python bench.py --layout tiled --steps 200 --warmup 10 --no-cpu-baseline --mcts-trees 0 --e2e-steps 100 > $O/${T}_bench_tiled.json 2>/dev/null; cut -c1-160 $O/${T}_bench_tiled.json
