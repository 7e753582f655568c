#!/bin/bash
# One GPU visit: resident-path tests, a short full-length bench, the ncu launch list and one full capture of the hot kernel.
#   tools/gpu_check.sh <tag> [pytest -k expression]
tag=${1:-x}; sel=${2:-"resident or full_length or small_and_ragged or sharding or fast"}
out=gpurun_out
timeout -k 10 240 python tools/gpu_sanity.py > $out/${tag}_sanity.log 2>&1; rc=$?
tail -8 $out/${tag}_sanity.log
if [ $rc -ne 0 ]; then echo "SANITY FAILED rc=$rc"; exit 1; fi
timeout -k 10 900 python -m pytest tests -m gpu -x -q -k "$sel" 2>&1 | tail -15 > $out/${tag}_tests.log
tail -5 $out/${tag}_tests.log
timeout -k 10 400 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || { tail -5 $out/${tag}_bench.err; exit 1; }
python - <<PY
import json
d = json.load(open("$out/${tag}_bench.json"))
print("value %.4g e2e %.4g frac %.4f launch_ms %.1f ms/step %.1f launches %d parity %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["launch_ms"], d["ms_per_step"], d["gpu_launches"], d.get("parity", {}).get("max_rel")))
PY
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --members 1184 --nstep 8760 --steps 2 --warmup 3 --no-cpu-baseline --no-exact > $out/${tag}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_physics_resident_v2 -c 1 -f -o $out/${tag}_v2 \
    python bench.py --members 592 --nstep 2048 --steps 1 --warmup 3 --no-cpu-baseline --no-exact > $out/${tag}_ncu_full.log 2>&1
python tools/ncu_summary.py $out/${tag}_v2.ncu-rep $((592*500*2048)) > $out/${tag}_ncu_v2.txt 2>&1
head -30 $out/${tag}_ncu_v2.txt
fi
if [ -n "$DRAM_FULL" ]; then
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name regex:k_physics_resident_v2 -c 1 --csv \
    --log-file $out/${tag}_dram_fullsize.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-exact > $out/${tag}_ncu_dram.log 2>&1
tail -4 $out/${tag}_dram_fullsize.csv
fi
