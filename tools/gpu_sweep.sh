#!/bin/bash
# usage: bash tools/gpu_sweep.sh TAG name1 name2 ...   (variants under build/variants/, "default" = the in-tree library)
O=gpurun_out; mkdir -p $O
T=$1; shift
for V in "$@"; do
  ENVV=""
  case $V in
    default) unset IPP_B200_LIB;;
    w15) unset IPP_B200_LIB; ENVV="IPP_BULK_WARPS=15";;
    w14r) unset IPP_B200_LIB; ENVV="IPP_BULK_WARPS=14";;
    *) export IPP_B200_LIB=$PWD/build/variants/$V.so;;
  esac
  env $ENVV timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --mcts-trees 0 --e2e-steps 2 > $O/${T}_sweep_$V.json 2> $O/${T}_sweep_$V.err
  python - <<PY
import json
try:
    d=json.load(open("$O/${T}_sweep_$V.json")); r=d["roofline"]
    print("%-8s step %.1f M (entropy %.1f) predict %.1f M" % ("$V", d["value"]/1e6, r["modes"]["gauss_entropy"]["value"]/1e6, r["predict"]["value"]/1e6))
except Exception as e:
    print("$V", "failed", e)
PY
done
unset IPP_B200_LIB
