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
    from parity import compare_perf
    tab = formats.read_perf_table(p.out_prefix + "ensemble_perf.bin")
    compare_perf(tab["perf"], ref["perf"])
    assert np.array_equal(tab["mcpar"], ref["mcpar"])            # the parameter sets sit next to their measures (after the clip)
    assert np.array_equal(tab["status"], ref["status"])
    p.close()


def test_cli_selected_members_binary_flows_and_two_ranks(tmp_path):
    """Ensemble-scale output: text files for the selected members only, their flows as FP64,
    measures of every member in the table; the same run over two ranks (dcp_run_ensemble_multi; both on device 0 when
    the box has one GPU is not possible from the CLI, so -gpus 2 is exercised only with two devices)."""
    import glob
    root = str(tmp_path)
    raw = synth.make_raw(nac=40, nriv=3, npt=2, ngp=2, nge=1, nstep=400, kW=4, seed=101, tree=True)
    ll, ul = synth.example_bounds(2)
    formats.write_project(raw, root, ll, ul, numsim=40, seeds=(5, 2000))
    r = subprocess.run([EXE, "-dir", root, "-ids", "1", "1", "1", "-write-members", "3:5,17,40", "-flow-bin", "-fast"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    p = host.Project(root)
    cat = p.catchment()
    ref = oracle_run(cat, p.genpar(40).copy(order="F"))
    want = [3, 4, 5, 17, 40]
    files = sorted(glob.glob(p.out_prefix + "mc_id_*.flow"))
    assert [int(f[-13:-5]) for f in files] == want and len(glob.glob(p.out_prefix + "mc_id_*.res")) == 5
    from parity import compare_flows, compare_perf
    ft = formats.read_flow_table(p.out_prefix + "ensemble_flow.bin")
    assert ft["members"].tolist() == want
    compare_flows(ft["q"], ref["q"][:, :, [m - 1 for m in want]])
    flow = np.loadtxt(files[3], ndmin=2).T
    assert np.abs(flow - ft["q"][:, :, 3]).max() <= 0.5e-8 + 1e-12
    tab = formats.read_perf_table(p.out_prefix + "ensemble_perf.bin")
    compare_perf(tab["perf"], ref["perf"])
    assert np.array_equal(tab["mcpar"], ref["mcpar"]) and tab["perf"].shape[2] == 40
    from decipher_b200.capi import load_library
    if load_library().dcp_device_count() >= 2:
        r2 = subprocess.run([EXE, "-dir", root, "-ids", "1", "1", "1", "-no-flow", "-fast", "-gpus", "2"], capture_output=True, text=True, timeout=300)
        assert r2.returncode == 0, r2.stdout + r2.stderr
        tab2 = formats.read_perf_table(p.out_prefix + "ensemble_perf.bin")
        # (the single-device pass also produced reach totals: another instantiation of the FAST kernel, so 1e-9, not bits)
        compare_perf(tab2["perf"], ref["perf"])
        assert np.array_equal(tab2["status"], tab["status"]) and np.array_equal(tab2["mcpar"], tab["mcpar"])
        assert "NCCL gather" in r2.stdout
    bad = subprocess.run([EXE, "-dir", root, "-ids", "1", "1", "1", "-write-members", "0:3"], capture_output=True, text=True, timeout=60)
    assert bad.returncode == 2
    p.close()
