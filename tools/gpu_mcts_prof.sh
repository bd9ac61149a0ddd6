#!/bin/bash
# per-launch metrics and full captures of the search kernels on growing trees (tools/mcts_probe.py ... peaked)
O=gpurun_out; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"mcts_(select|expand)" -c 200 --csv --log-file $O/mp_launches.csv python tools/mcts_probe.py split 100 peaked > /dev/null 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mcts_select -s 60 -c 1 -f -o $O/mp_select python tools/mcts_probe.py split 100 peaked > /dev/null 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mcts_expand -s 60 -c 1 -f -o $O/mp_expand python tools/mcts_probe.py split 100 peaked > /dev/null 2>&1; echo "rc=$?"
ls -la $O/mp_*
