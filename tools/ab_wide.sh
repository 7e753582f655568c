#!/bin/bash
# A/B of kernel variants on the wide workload (nac = 1000): tools/ab_wide.sh <nstep> <variant dir> ...
nstep=$1; shift
for v in "$@"; do
  DCP_B200_LIB=decipher_b200/csrc/$v/libdecipher_b200.so timeout -k 5 300 python bench.py --nac 1000 --steps 3 --warmup 3 --nstep $nstep --no-cpu-baseline --no-exact > gpurun_out/abw_$v.json 2> gpurun_out/abw_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/abw_$v.json"))
    print("%-22s nac 1000 value %.4g  launch_ms %.1f" % ("$v", d["value"], d["roofline"]["launch_ms"]))
except Exception as e:
    print("$v failed", e)
PY
done
