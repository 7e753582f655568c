"""BASELINE config 2 in full, once: 500-HRU synthetic catchment, 87 600 hourly steps, 10^5 RANMAR-sampled members on one
B200 through the public host-buffer API (parameters in, five measures + status + init diagnostics out).
usage: c2_full.py [members]"""
import json, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from decipher_b200 import synth
from decipher_b200.capi import DeviceCatchment, OPT_FAST_MATH, PERF_NSE, PERF_KGE

M = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
cat = synth.synth_catchment("C2")
mc = synth.sample_parameters(cat.npt, M)
dev = DeviceCatchment(cat)
dev.run(np.asfortranarray(mc[:, :, :592].copy()), flags=OPT_FAST_MATH)          # warm-up: one wave
t0 = time.perf_counter()
out = dev.run(mc, flags=OPT_FAST_MATH, want_q=False, want_perf=True)
wall = time.perf_counter() - t0
s = out["stats"]
units = M * cat.nac * cat.nstep
nse = out["perf"][PERF_NSE]
print(json.dumps({"config": "C2 full", "members": M, "nac": cat.nac, "nstep": cat.nstep, "units": units,
                  "wall_s": wall, "device_s": s["ms_total"] * 1e-3, "units_per_s_wall": units / wall,
                  "units_per_s_device": units / (s["ms_total"] * 1e-3), "ms_init": s["ms_init"], "ms_physics": s["ms_physics"],
                  "ms_route": s["ms_route"], "n_physics_launches": s["n_physics_launches"], "n_launches": s["n_launches"],
                  "status_nonzero_solflag_hist": np.bincount(out["status"] & 3, minlength=3).tolist(),
                  "members_with_nonfinite": int(((out["status"] & 0x20) != 0).sum()),
                  "nse_gauge0_quantiles": np.nanquantile(nse[0], [0.0, 0.5, 0.9, 1.0]).tolist(),
                  "kge_gauge0_max": float(np.nanmax(out["perf"][PERF_KGE][0]))}))
dev.close()
