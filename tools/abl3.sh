python -m pytest tests/test_gpu_paths_and_scale.py -x -q 2>&1 | tail -2
IPP_ASYNC_WARPS=15 bash tools/run_ablation.sh w16
bash tools/run_ablation.sh w16 w20 w24 w28
IPP_ASYNC_WARPS=18 bash tools/run_ablation.sh w20
IPP_ASYNC_WARPS=22 bash tools/run_ablation.sh w24
bash tools/run_ablation.sh w20
python tools/bench_configs.py 2>&1 | tail -8
