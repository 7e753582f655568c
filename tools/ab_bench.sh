#!/bin/bash
# A/B of kernel variants on the bench workload: tools/ab_bench.sh <nstep> <variant dir> [<variant dir> ...]   (dirs under decipher_b200/csrc)
nstep=$1; shift
for v in "$@"; do
  DCP_B200_LIB=decipher_b200/csrc/$v/libdecipher_b200.so timeout -k 5 300 python bench.py --steps 3 --warmup 3 --nstep $nstep --no-cpu-baseline --no-exact > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_$v.json"))
    print("%-22s value %.4g  launch_ms %.1f  frac %.4f" % ("$v", d["value"], d["roofline"]["launch_ms"], d["roofline"]["frac"]))
except Exception as e:
    print("$v failed", e)
PY
done
