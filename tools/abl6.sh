IPP_B200_LIB=$PWD/build/exp/lib_d16.so python -m pytest tests/test_gpu_paths_and_scale.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
bash tools/run_ablation.sh s16 d16 d20 d24 d32 d20
IPP_B200_LIB=$PWD/build/exp/lib_d20.so python -m pytest tests/test_gpu_paths_and_scale.py -x -q 2>&1 | tail -2
