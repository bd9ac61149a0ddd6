#!/bin/bash
# SPLIT layout bring-up: parity subset + predict probe on both layouts
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths_and_scale.py tests/test_gpu_reset.py tests/test_gpu_mcts.py tests/test_gpu_full_size_parity.py -m gpu -x -q -k "4 or split or persistent_paths" > gpurun_out/s1_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/s1_pytest.log
timeout 300 python tools/predict_probe.py super > gpurun_out/s1_probe_super.txt 2>&1; cat gpurun_out/s1_probe_super.txt
timeout 300 python tools/predict_probe.py split > gpurun_out/s1_probe_split.txt 2>&1; cat gpurun_out/s1_probe_split.txt
