#!/bin/bash
# BASELINE config 4 bench lines (5 000 HRUs; kW = 64 / 512 / dense) + the DFMA-vs-DMMA A/B of the dense contraction with ncu.
out=gpurun_out
run() { # tag, extra env, args...
  tag=$1; shift; envs=$1; shift
  env $envs timeout -k 10 600 python bench.py --config C4 "$@" > $out/c4_$tag.json 2> $out/c4_$tag.err || { echo "$tag FAILED"; tail -3 $out/c4_$tag.err; return; }
  python - <<PY
import json
d = json.load(open("$out/c4_$tag.json"))
r = d["roofline"]
print("%-14s value %.4g e2e %.4g  contraction %.2f TFLOP/s  launches %d  parity %s  cpu %.3g" % ("$tag", d["value"], d["e2e"]["value"], r["contraction_tflops"], d["gpu_launches"], d.get("parity", {}).get("max_rel"), d.get("cpu_baseline", {}).get("value", 0)))
PY
}
run kw64 X=1 --kw 64 --steps 2 --warmup 3
run kw512 X=1 --kw 512 --nstep 4380 --steps 2 --warmup 3
run dense_mma DCP_DENSE_MMA=1 --kw dense --nstep 438 --steps 2 --warmup 3
run dense_dfma DCP_DENSE_MMA=0 --kw dense --nstep 438 --steps 2 --warmup 3 --no-cpu-baseline
for v in 1 0; do
  DCP_DENSE_MMA=$v timeout -k 10 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_redistribute_dense -c 1 -s 20 -f -o $out/c4_gemm_mma$v \
     python bench.py --config C4 --kw dense --nstep 32 --members 1024 --steps 1 --warmup 3 --no-cpu-baseline > $out/c4_ncu_mma$v.log 2>&1
  python tools/ncu_summary.py $out/c4_gemm_mma$v.ncu-rep > $out/c4_ncu_gemm_mma$v.txt 2>&1
  grep -E "gpu__time_duration|pipe_fp64|issue_active|registers_per_thread|bank_conflicts|wavefronts_mem_shared|warps_active" $out/c4_ncu_gemm_mma$v.txt
done
