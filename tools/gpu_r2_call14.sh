#!/bin/bash
O=gpurun_out; mkdir -p $O
T=c14
timeout 900 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_missions.py tests/test_gpu_full_size_parity.py::test_mcts_full_size_sampled_trees_vs_oracle -x -q > $O/${T}_mcts.log 2>&1; echo "mcts rc=$?"; tail -4 $O/${T}_mcts.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; tail -3 $O/${T}_bench.err
python - <<PY
import json
d=json.load(open("$O/${T}_bench.json"))
print("mcts %.4f ms/sim %.0f M pred/s" % (d["mcts_rollouts"]["ms_per_lockstep_simulation"], d["mcts_rollouts"]["prediction_steps_per_sec"]/1e6))
PY
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:"mcts_|rollout" -c 120 --csv --log-file $O/r02b_mcts_launches.csv \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-sims 24 > /dev/null 2>&1
grep -c mcts_select $O/r02b_mcts_launches.csv
