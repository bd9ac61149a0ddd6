#!/bin/bash
mkdir -p gpurun_out
echo "== default (guide predict 2)"; timeout 300 python tools/predict_probe.py split 2>&1 | grep predict
echo "== gp1"; IPP_B200_LIB=build/variants/libipp_gp1.so timeout 300 python tools/predict_probe.py split 2>&1 | grep predict
echo "== gp1 w20"; IPP_BULK_PREDICT_WARPS=20 IPP_B200_LIB=build/variants/libipp_gp1.so timeout 300 python tools/predict_probe.py split 2>&1 | grep predict
timeout 300 python tools/e2e_breakdown.py 2>&1 | tee gpurun_out/s4_e2e_breakdown.txt
lscpu | grep -E "Model name|^CPU\(s\)|MHz" | head -5
