#!/bin/bash
O=gpurun_out; mkdir -p $O
# full capture of the covariance-only persistent kernel on SPLIT (committing), and per-launch metrics of the search kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ipp_step_bulk_kernelILi1ELb0ELb0ELb0ELb1" -s 8 -c 1 -f -o $O/p1_predict_split \
  python tools/predict_probe.py split > $O/p1_ncu_predict.log 2>&1; echo "ncu predict rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,smsp__warps_active.avg.per_cycle_active --clock-control none -k regex:"mcts_" -c 400 --csv --log-file $O/p1_mcts_launches.csv \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-sims 60 > /dev/null 2>&1; echo "ncu mcts rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mcts_select" -s 150 -c 1 -f -o $O/p1_select python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-sims 60 > /dev/null 2>&1; echo "ncu select rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mcts_expand" -s 150 -c 1 -f -o $O/p1_expand python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-sims 60 > /dev/null 2>&1; echo "ncu expand rc=$?"
ls -la $O/p1_*
