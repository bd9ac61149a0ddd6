#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, one full ncu capture of the step kernel.
# usage (under gpurun): bash tools/gpu_round.sh TAG
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cat $O/${TAG}_bench.json
python bench.py --impl reference --steps 5 --warmup 3 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; cat $O/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $O/${TAG}_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipp_step_async -s 6 -c 1 -f -o $O/${TAG}_async \
  python bench.py --steps 8 --warmup 3 --batch 16384 --no-cpu-baseline --e2e-steps 2 > $O/${TAG}_ncu_bench.log 2>&1
ls -la $O
