"""The CUDA path against the reference's own source, executed (tests/golden/f90_exec/, see tests/test_f90_pin_cpu.py):
flows within BASELINE's 1e-9 relative, the reach totals of the .res file, parameter sets and flags exactly.  No oracle in
between: the committed numbers are what the reference's Fortran statements produce for these inputs."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import f90_cases  # noqa: E402
from decipher_b200.capi import (DeviceCatchment, KERNEL_RESIDENT, KERNEL_STREAMED, OPT_CHECK_BALANCE, OPT_FAST_MATH,
                                ST_RZ_BALANCE, ST_SOLFLAG_MASK, ST_TS_OUTSUM, ST_TS_WBAL)  # noqa: E402
from parity import compare_close, compare_flows  # noqa: E402

pytestmark = pytest.mark.gpu
VEC = os.path.join(HERE, "golden", "f90_exec")


@pytest.mark.parametrize("name", sorted(f90_cases.CASES))
@pytest.mark.parametrize("kernel,flags", [(KERNEL_STREAMED, 0), (KERNEL_RESIDENT, 0), (KERNEL_RESIDENT, OPT_FAST_MATH)])
def test_cuda_reproduces_the_executed_reference(name, kernel, flags):
    v = np.load(os.path.join(VEC, name + ".npz"), allow_pickle=False)
    cat, _ = f90_cases.build(name)
    mc = np.asfortranarray(v["mcpar_in"])
    dev = DeviceCatchment(cat)
    out = dev.run(mc.copy(order="F"), flags=flags | OPT_CHECK_BALANCE, kernel=kernel, want_q=True, want_totals=True)
    dev.close()
    stop = [str(s) for s in v["stop"]]
    for m in range(mc.shape[2]):
        st = int(out["status"][m])
        assert (st & ST_SOLFLAG_MASK) == int(v["sol_flag"][m]), (m, st)
        assert bool(st & ST_RZ_BALANCE) == ("dyna_root_zone1" in stop[m]), (m, st, stop[m])
        assert bool(st & (ST_TS_OUTSUM | ST_TS_WBAL)) == ("dyna_results_ts" in stop[m]), (m, st, stop[m])
        if stop[m]:
            continue
        assert np.array_equal(out["mcpar"][:, :, m], v["mcpar_written"][:, :, m])
        compare_close(out["init_diag"][:, m] * 1000, v["init"][:, m], what="init block of the .res file")
        qr = v["q"][:, :, m:m + 1]
        if np.abs(qr).mean() < 1e-7:                              # below 1 mm per YEAR at hourly steps
            # h_dry_start: flows of 1e-8 .. 1e-18 m per step are (sd_new - sd_old)/dtt of stores of 1e-2 .. 3 m
            # (dyna_satzone_analyticalsolver.f90:98): the rounding of the stores (>= 1e-18) IS the flow, in the reference too,
            # and a relative contract means nothing there -- such members are held to 1e-15 m per step absolutely
            assert np.abs(out["q"][:, :, m:m + 1] - qr).max() <= 1e-15
        else:
            compare_flows(out["q"][:, :, m:m + 1], qr, floor=1e-2 if name.startswith(("h_", "d_", "i_")) else 1e-3)
        compare_close(out["reach_totals"][:, :, m], v["totals"][:, :, m], what="reach totals")


@pytest.mark.parametrize("name", sorted(f90_cases.PROJECTS))
@pytest.mark.parametrize("fast", [False, True])
def test_cli_on_project_files_reproduces_the_executed_program(name, fast, tmp_path):
    """The whole drop-in (decipher_b200_run: C++ readers + RANMAR/genpar + C ABI + CUDA) on a project directory against
    `program main` of the reference executed on the same files (tests/golden/f90_exec/p_*.npz)."""
    import subprocess
    from decipher_b200 import formats, host
    exe = os.path.join(os.path.dirname(HERE), "decipher_b200", "csrc", "build", "decipher_b200_run")
    v = np.load(os.path.join(VEC, name + ".npz"), allow_pickle=False)
    root = str(tmp_path)
    f90_cases.write_project(name, root)
    n = int(v["numsim"])
    cmd = [exe, "-dir", root, "-ids", "1", "1", "1", "-write-members", f"1:{n}", "-flow-bin"] + (["-fast"] if fast else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    p = host.Project(root)
    ft = formats.read_flow_table(p.out_prefix + "ensemble_flow.bin")
    assert ft["members"].tolist() == list(range(1, n + 1))
    compare_flows(ft["q"], v["q"])
    flow = np.loadtxt(p.out_prefix + f"mc_id_{n:08d}.flow", ndmin=2).T                     # the reference's own text format
    assert np.abs(flow - v["q"][:, :, n - 1]).max() <= 0.5e-8 + 1e-12
    tab = formats.read_perf_table(p.out_prefix + "ensemble_perf.bin")
    unclipped = tab["mcpar"][:, 3, :] == v["mcpar"][:, 3, :]                                # srinit may be clipped to srmax afterwards
    assert np.array_equal(np.delete(tab["mcpar"], 3, axis=1), np.delete(v["mcpar"], 3, axis=1)) and unclipped.mean() > 0.5
    p.close()


@pytest.mark.parametrize("sea", [False, True])
def test_router_cli_reproduces_the_executed_router_program(sea, tmp_path):
    """decipher_route_run (files -> C++ front end -> GPU router -> <series>_routed.txt) against the reference's main program
    route_run_standalone executed from source on the same files (its velocity is hard-wired to 0.133)."""
    import subprocess
    from decipher_b200 import formats
    exe = os.path.join(os.path.dirname(HERE), "decipher_b200", "csrc", "build", "decipher_route_run")
    v = np.load(os.path.join(VEC, "r_router_program_%s.npz" % ("sea" if sea else "plain")), allow_pickle=False)
    cat, x = f90_cases.router_network(sea)
    paths = formats.write_router_inputs(cat, x, str(tmp_path))
    r = subprocess.run([exe, "-river_data", paths["river_data"], "-tree", paths["flow_conn"], "-timeseries_in", paths["timeseries"],
                        "-velocity", repr(f90_cases.ROUTER_VELOCITY)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    out = os.path.join(str(tmp_path), str(v["out_file"]))                         # the name the reference derives (:191)
    lines = open(out).read().splitlines()
    assert lines[0].split("\t") == [str(h) for h in v["headers"]]
    vals = np.array([[float(a) for a in ln.split(" ")] for ln in lines[1:]])
    assert vals.shape == v["routed"].shape and np.abs(vals - v["routed"]).max() <= 0.5e-6 + 1e-12       # F0.6 text
