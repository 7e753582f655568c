"""End to end through the drop-in boundary: a project in the reference's file formats -> the
decipher_b200_run driver (C++ host + C ABI + CUDA) -> mc_id_*.flow/.res and the binary measure
table, compared with the CPU oracle run on the arrays the C++ reader produced."""
import os
import subprocess

import numpy as np
import pytest

from decipher_b200 import formats, host, synth
from decipher_b200.capi import OPT_CHECK_BALANCE
from oracle_binding import oracle_run

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "decipher_b200", "csrc", "build", "decipher_b200_run")


def test_cli_matches_oracle(tmp_path):
    root = str(tmp_path)
    raw = synth.make_raw(nac=30, nriv=2, npt=3, ngp=2, nge=1, nstep=600, kW=4, seed=99)
    ll, ul = synth.example_bounds(3)
    formats.write_project(raw, root, ll, ul, numsim=6, seeds=(5, 2000))
    r = subprocess.run([EXE, "-dir", root, "-auto", os.path.join(root, "auto.dat")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    p = host.Project(root)
    cat = p.catchment()
    mc = p.genpar(6)
    ref = oracle_run(cat, mc.copy(order="F"), flags=0)
    for m in range(6):
        flow = np.loadtxt(p.out_prefix + f"mc_id_{m + 1:08d}.flow", ndmin=2).T          # (nriv, nstep)
        assert flow.shape == (2, 600)
        assert np.abs(flow - ref["q"][:, :, m]).max() <= 0.5e-8 + 1e-12                 # f0.8 text precision
        res = open(p.out_prefix + f"mc_id_{m + 1:08d}.res").read().splitlines()
        i = res.index("! PARAMETERS")
        row = res[i + 2]
        assert float(row[12:22]) == pytest.approx(ref["mcpar"][0, 0, m], abs=5.1e-5)
        j = res.index("! ET TOTALS (mm)")
        assert [float(x) for x in res[j + 1].split()] == pytest.approx(list(ref["reach_totals"][2, :, m]), abs=5.1e-3)
    raw_tab = open(p.out_prefix + "ensemble_perf.bin", "rb").read()
    hdr = np.frombuffer(raw_tab[:24], dtype=np.int64)
    from decipher_b200.capi import NMEASURE
    from parity import compare_perf
    assert hdr.tolist() == [6, 2, NMEASURE]
    n = NMEASURE * 2 * 6
    perf = np.frombuffer(raw_tab[24:24 + 8 * n], dtype=np.float64).reshape((NMEASURE, 2, 6), order="F")
    compare_perf(perf, ref["perf"])
    p.close()
