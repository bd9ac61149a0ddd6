#!/bin/bash
# round 2 evidence: full GPU test suite, sanitizer, launch list, full ncu capture + DRAM traffic of the step kernel, MCTS kernels, bench lines
O=gpurun_out; mkdir -p $O
T=${1:-r02}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/${T}_smoke.log
timeout 900 compute-sanitizer --tool memcheck --log-file $O/${T}_sanitizer_memcheck.log python tools/sanitize_small.py > $O/${T}_sanitizer_memcheck.out 2>&1; echo "memcheck rc=$?"; tail -2 $O/${T}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --log-file $O/${T}_sanitizer_racecheck.log python tools/sanitize_small.py 1 3 > $O/${T}_sanitizer_racecheck.out 2>&1; echo "racecheck rc=$?"; tail -2 $O/${T}_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --log-file $O/${T}_sanitizer_synccheck.log python tools/sanitize_small.py 3 > $O/${T}_sanitizer_synccheck.out 2>&1; echo "synccheck rc=$?"; tail -2 $O/${T}_sanitizer_synccheck.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${T}_launches.csv \
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 4 --mcts-sims 8 > $O/${T}_launches_bench.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:ipp_step_bulk -s 4 -c 6 --csv \
  --log-file $O/${T}_traffic_super.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-trees 0 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipp_step_bulk -s 6 -c 1 -f -o $O/${T}_bulk_step \
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-trees 0 > $O/${T}_ncu_bench.log 2>&1
# the covariance-only persistent kernel on the split layout: launches 14.. of predict_probe.py split are the committing predict leg
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:ipp_step_bulk -s 114 -c 6 --csv \
  --log-file $O/${T}_traffic_predict_split.csv python tools/predict_probe.py split > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipp_step_bulk -s 116 -c 1 -f -o $O/${T}_predict_split \
  python tools/predict_probe.py split > $O/${T}_ncu_predict.log 2>&1
# the search kernels on growing trees (simulation 60 of tools/mcts_probe.py ... peaked)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mcts_select -s 60 -c 1 -f -o $O/${T}_mcts_select \
  python tools/mcts_probe.py split 100 peaked > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mcts_expand -s 60 -c 1 -f -o $O/${T}_mcts_expand \
  python tools/mcts_probe.py split 100 peaked > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:"mcts_(select|expand)" -c 200 --csv --log-file $O/${T}_mcts_launches.csv \
   python tools/mcts_probe.py split 100 peaked > /dev/null 2>&1
timeout 300 python tools/predict_probe.py split > $O/${T}_predict_probe_split.txt 2>&1
timeout 300 python tools/predict_probe.py super > $O/${T}_predict_probe_super.txt 2>&1
timeout 300 python tools/e2e_ab.py > $O/${T}_e2e_ab.txt 2>&1
timeout 300 python tools/mcts_probe.py split 100 uniform > $O/${T}_mcts_probe.txt 2>&1
timeout 300 python tools/mcts_probe.py split 100 peaked >> $O/${T}_mcts_probe.txt 2>&1
timeout 600 python tools/bench_configs.py > $O/${T}_configs_throughput.jsonl 2>&1
python bench.py --impl reference --steps 5 --warmup 3 > $O/${T}_bench_ref.json 2>/dev/null; cut -c1-200 $O/${T}_bench_ref.json
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_k20.json 2> $O/${T}_bench_k20.err; cut -c1-200 $O/${T}_bench_k20.json
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; cut -c1-300 $O/${T}_bench.json
ls -la $O | tail -20
