"""the CPU port of the reference path timed on this box's host cores."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from decipher_b200 import synth, formats, host
from oracle_binding import oracle_run
cores = os.cpu_count()
cat = synth.synth_catchment("C2", nstep=8760)
def t(n, threads, opt):
    mc = synth.sample_parameters(1, n)
    t0 = time.perf_counter(); oracle_run(cat, mc, want_q=False, want_totals=False, n_threads=threads, opt=opt); dt = time.perf_counter() - t0
    return n * cat.nac * cat.nstep / dt, dt
print("cores", cores)
v, dt = t(2, 1, "O0"); print(f"B1 oracle -O0, 1 thread: {v:.3e} units/s ({dt:.1f} s)")
v, dt = t(4, 1, "O2"); print(f"B2 oracle -O2 strict FP, 1 thread: {v:.3e} units/s ({dt:.1f} s)")
v, dt = t(4 * cores, cores, "O2"); print(f"B3 oracle -O2, {cores} threads: {v:.3e} units/s ({dt:.1f} s)")
# B4: C1 with text output
c1 = synth.make_raw(**{k: v for k, v in synth.CONFIGS["C1"].items() if k != "members"})
root = tempfile.mkdtemp()
ll, ul = synth.example_bounds(3)
formats.write_project(c1, root, ll, ul, numsim=10)
p = host.Project(root); cat1 = p.catchment(); mc = p.genpar(10)
t0 = time.perf_counter()
r = oracle_run(cat1, mc, n_threads=1)
for m in range(10):
    p.write_member(m + 1, r["mcpar"][:, :, m], r["q"][:, :, m], r["reach_totals"][:, :, m], r["init_diag"][:, m], r["status"][m])
dt = time.perf_counter() - t0
print(f"B4 C1 (10 members x 24 HRUs x 8760 steps) incl. .res/.flow text output, 1 thread: {10*24*8760/dt:.3e} units/s ({dt:.2f} s)")
