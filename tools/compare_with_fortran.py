#!/usr/bin/env python
"""Pins the oracle (and through it the GPU path) against the REAL reference wherever a Fortran compiler exists.

No Fortran compiler is available in the build image, so parity is unpinned upstream; this tool is
the missing link for a maintainer who has one:

  1. `--make-project DIR`  writes a small synthetic project in the reference's own file formats
     (project.dat, settings.dat, params.dat, model_structure.dat, DTA files, forcing; decipher_b200/formats.py).
  2. The maintainer builds the unmodified reference (RRMODEL/debug/makefile or RRMODEL/release/makefile, gfortran)
     and runs it inside DIR:   cd DIR && /path/to/DECIPHeR_v1.exe -auto auto.dat
     (`--run-exe EXE` does that for them), which writes <prefix>mc_id_%08d.flow / .res per member.
  3. `--compare DIR` reads those files and checks them against the CPU oracle run on the same project:
     parameter sets of the .res table to the printed 4 decimals, flows to the f0.8 text precision
     (0.5e-8 absolute: the text files cannot carry the 1e-9 relative check).

Exit status 0 = every member within tolerance."""
import argparse
import glob
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def make_project(root, members):
    from decipher_b200 import formats, synth
    os.makedirs(root, exist_ok=True)
    raw = synth.make_raw(nac=24, nriv=2, npt=3, ngp=2, nge=1, nstep=600, kW=4, seed=synth.BASE_SEED + 1)
    ll, ul = synth.example_bounds(3)
    formats.write_project(raw, root, ll, ul, numsim=members, seeds=(5, 2000))
    print(f"project written to {root}: run the reference there with `-auto auto.dat`")


def compare(root, flow_tol=0.5e-8 + 1e-12):
    from decipher_b200 import host
    from oracle_binding import oracle_run
    p = host.Project(root, 1, 1, 1)
    cat = p.catchment()
    files = sorted(glob.glob(p.out_prefix + "mc_id_*.flow"))
    if not files:
        print(f"no {p.out_prefix}mc_id_*.flow files: run the reference executable in {root} first")
        return 2
    n = len(files)
    mc = p.genpar(n)
    ref = oracle_run(cat, mc.copy(order="F"), flags=0)
    worst, bad = 0.0, 0
    for m, f in enumerate(files):
        flow = np.loadtxt(f, ndmin=2).T                                   # (nriv, nstep)
        if flow.shape != (cat.nriv, cat.nstep):
            print(f"{os.path.basename(f)}: shape {flow.shape}, expected {(cat.nriv, cat.nstep)}")
            bad += 1
            continue
        err = float(np.nanmax(np.abs(flow - ref["q"][:, :, m])))
        worst = max(worst, err)
        res = open(f[:-5] + ".res").read().splitlines()
        i = res.index("! PARAMETERS")
        row = res[i + 2]
        par_ok = abs(float(row[12:22]) - ref["mcpar"][0, 0, m]) <= 5.1e-5     # SZM of layer 1, F10.4
        if not (err <= flow_tol) or not par_ok:
            print(f"{os.path.basename(f)}: max |flow - oracle| = {err:.3e} m/ts, parameters {'ok' if par_ok else 'DIFFER'}")
            bad += 1
    p.close()
    print(f"{n - bad}/{n} members within the text precision; worst flow difference {worst:.3e} m/timestep")
    return 0 if bad == 0 else 1


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--make-project", metavar="DIR")
    ap.add_argument("--members", type=int, default=6)
    ap.add_argument("--run-exe", metavar="EXE", help="reference executable to run inside DIR before comparing")
    ap.add_argument("--compare", metavar="DIR")
    a = ap.parse_args()
    if a.make_project:
        make_project(a.make_project, a.members)
    if a.compare:
        if a.run_exe:
            subprocess.run([os.path.abspath(a.run_exe), "-auto", "auto.dat"], cwd=a.compare, check=True)
        sys.exit(compare(a.compare))


if __name__ == "__main__":
    main()
