#!/bin/bash
# Evidence for profiles/: launch list of the bench command, DRAM traffic of the step kernel at the full batch, DRAM probe.
O=gpurun_out
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r01_launches.csv \
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-sims 8 > $O/r01_launches_bench.log 2>&1
for L in mv tiled; do
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ipp_step_async -s 8 -c 4 --csv \
  --log-file $O/r01_traffic_$L.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-trees 0 --layout $L > $O/r01_traffic_bench_$L.log 2>&1
done
timeout 300 build/dram_probe2 24 32 > $O/r01_dram_probe2.txt 2>&1
ls -la $O
