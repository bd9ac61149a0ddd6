#!/bin/bash
O=gpurun_out; mkdir -p $O
T=c6
timeout 600 python -m pytest tests/test_gpu_reset.py -q > $O/${T}_reset.log 2>&1; echo "reset rc=$?"; tail -15 $O/${T}_reset.log
timeout 300 python tools/e2e_probe.py super > $O/${T}_probe.log 2>&1; cat $O/${T}_probe.log
