#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_full_size_parity.py tests/test_gpu_paths_and_scale.py tests/test_gpu_parity.py -m gpu -x -q -k "predict or persistent_paths or edge" > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s6_pytest.log
echo "== tma store"; timeout 300 python tools/predict_probe.py split 2>&1 | grep predict
echo "== stg"; IPP_B200_LIB=build/variants/libipp_notma.so timeout 300 python tools/predict_probe.py split 2>&1 | grep predict
