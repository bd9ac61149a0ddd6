#!/bin/bash
# Evidence for profiles/: launch list of the bench command, DRAM traffic of the step kernel at the full batch (both layouts),
# one full ncu capture of the step kernel (default layout), DRAM probe.   usage (under gpurun): bash tools/gpu_profiles.sh TAG
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-sims 8 > $O/${TAG}_launches_bench.log 2>&1
for L in mv tiled; do
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ipp_step_async -s 8 -c 4 --csv \
  --log-file $O/${TAG}_traffic_$L.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 2 --mcts-trees 0 --layout $L > $O/${TAG}_traffic_bench_$L.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipp_step_async -s 6 -c 1 -f -o $O/${TAG}_async_tiled \
  python bench.py --steps 8 --warmup 3 --batch 65536 --no-cpu-baseline --e2e-steps 2 --mcts-trees 0 --layout tiled > $O/${TAG}_ncu_bench.log 2>&1
timeout 300 build/dram_probe2 24 32 > $O/${TAG}_dram_probe2.txt 2>&1
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; tail -2 $O/${TAG}_pytest.log
python bench.py --impl reference --steps 5 --warmup 3 > $O/${TAG}_bench_ref.json 2>/dev/null; cat $O/${TAG}_bench_ref.json
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; cat $O/${TAG}_bench.json
ls -la $O | tail -12
