"""CPU tests of the C++ host library: the reference's text formats parse to exactly the arrays the
generator assembled (reader arithmetic of dyna_tread_dyna / dyna_inputs / dyna_read_routingdata
restated independently in decipher_b200.synth.assemble), RANMAR/genpar are bit-exact with the oracle,
and the .res/.flow writers follow the reference's edit descriptors."""
import dataclasses
import os

import numpy as np
import pytest

from decipher_b200 import formats, host, synth
from oracle_binding import oracle_genpar, oracle_ranmar


@pytest.fixture(scope="module")
def project(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("proj"))
    raw = synth.make_raw(nac=40, nriv=3, npt=3, ngp=2, nge=2, nstep=300, kW=5, seed=1234, tree=True)
    ll, ul = synth.example_bounds(3)
    formats.write_project(raw, root, ll, ul, numsim=7, seeds=(5, 2000))
    return root, raw, ll, ul


def test_reader_matches_generator_bit_exact(project):
    root, raw, ll, ul = project
    p = host.Project(root, 1, 1, 1)
    got, want = p.catchment(), synth.assemble(raw)
    assert p.numsim == 7 and p.seeds == (5, 2000)
    assert np.array_equal(p.ll, ll) and np.array_equal(p.ul, ul)
    for f in dataclasses.fields(want):
        if f.name == "meta":
            continue
        a, b = getattr(got, f.name), getattr(want, f.name)
        if isinstance(b, np.ndarray):
            assert a.shape == b.shape, f.name
            assert np.array_equal(a, b), f"{f.name} differs (max abs {np.abs(a - b).max()})"
        else:
            assert a == b, f.name
    p.close()


def test_ranmar_and_genpar_bit_exact(project):
    root, raw, ll, ul = project
    assert (host.ranmar(1802, 9373, 20000, 6) * 2.0 ** 24).tolist() == [6533892.0, 14220222.0, 7275067.0, 6172232.0, 8354498.0, 10633180.0]
    assert np.array_equal(host.ranmar(5, 2000, 0, 500), oracle_ranmar(5, 2000, 0, 500))
    p = host.Project(root)
    assert np.array_equal(p.genpar(50), oracle_genpar(5, 2000, ll, ul, 50))
    assert np.array_equal(p.genpar(50), synth.sample_parameters(3, 50))
    p.close()


def test_auto_file_and_errors(project, tmp_path):
    root, *_ = project
    lib = host.load_host_library()
    import ctypes as C
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    assert lib.dcph_read_auto_file(os.path.join(root, "auto.dat").encode(), C.byref(a), C.byref(b), C.byref(c)) == 0
    assert (a.value, b.value, c.value) == (1, 1, 1)
    with pytest.raises(Exception):
        host.Project(str(tmp_path))                 # no project.dat
    with pytest.raises(Exception):
        host.Project(root, 1, 2, 1)                 # catchment setup id out of range


def test_res_and_flow_writers(project):
    root, raw, ll, ul = project
    p = host.Project(root)
    npt, nriv, nstep = 3, 3, 300
    mc = p.genpar(2)
    rng = np.random.default_rng(0)
    q = np.asfortranarray(rng.uniform(0, 2e-4, (nriv, nstep)))
    q[0, 0] = 0.0
    q[1, 5] = 1.23456789012
    # totals[k, r]: k = precip / PET / actual ET, r = river
    totals = np.asfortranarray(np.array([[981.602, 415.8675, 410.04], [0.4, 410.0, 0.5], [1200.0, 0.004, 399.999]]))
    diag = np.array([6.9393193900000002e-05, 5.9e-05, 1.7399e-02])
    p.write_member(3, mc[:, :, 1], q, totals, diag, status=1)
    res = open(p.out_prefix + "mc_id_00000003.res").read().splitlines()
    flow = open(p.out_prefix + "mc_id_00000003.flow").read().splitlines()
    assert len(flow) == nstep
    assert flow[0].split()[0] == ".00000000"                 # gfortran F0.8 drops the leading zero
    assert flow[5].split()[1] == "1.23456789"
    vals = np.array([[float(x) for x in ln.split()] for ln in flow]).T
    assert np.abs(vals - q).max() <= 0.5e-8 + 1e-15
    assert res[1] == " ! INITIALISATION" and res[2] == " Solution found!"
    assert float(res[3].split()[2]) == pytest.approx(6.9393193900000002e-02, rel=1e-15)
    i = res.index("! PARAMETERS")
    assert res[i + 1] == "Param_class     szm_def   lnto_def    srmax_def srinit_def chv_def     td_def     smax_def  "
    row = res[i + 2]
    assert len(row) == 12 + 10 + 12 + 10 + 10 + 13 + 11 + 10 and int(row[:12]) == 1
    assert float(row[12:22]) == pytest.approx(mc[0, 0, 1], abs=5e-5)
    j = res.index("! PRECIP TOTALS (mm)")
    assert res[j + 1] == " 981.60 415.87 410.04"
    k = res.index("! PET TOTALS (mm)")
    assert res[k + 1] == " .40 410.00 .50"
    assert res[res.index("! ET TOTALS (mm)") + 1] == " 1200.00 .00 400.00"
    p.write_member(4, mc[:, :, 1], q, totals, diag, status=2, f0_leading_zero=True)
    res2 = open(p.out_prefix + "mc_id_00000004.res").read().splitlines()
    assert res2[res2.index("! PET TOTALS (mm)") + 1] == " 0.40 410.00 0.50"
    assert res2[2].startswith(" Maximum QBF is less than")
    p.close()


def test_compare_with_fortran_tool_on_oracle_written_files(tmp_path):
    """tools/compare_with_fortran.py (the link to the real reference for anyone with a Fortran compiler): its
    project writer and comparator agree when the member files are written from the oracle's own results with
    the reference's edit descriptors (C++ host writers)."""
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location("cwf", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                      "tools", "compare_with_fortran.py"))
    cwf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cwf)
    from oracle_binding import oracle_run
    root = str(tmp_path / "proj")
    cwf.make_project(root, 4)
    assert cwf.compare(root) == 2                                     # nothing to compare yet
    p = host.Project(root, 1, 1, 1)
    mc = p.genpar(4)
    ref = oracle_run(p.catchment(), mc.copy(order="F"), flags=0)
    for m in range(4):
        p.write_member(m + 1, ref["mcpar"][:, :, m], ref["q"][:, :, m], ref["reach_totals"][:, :, m],
                       ref["init_diag"][:, m], int(ref["status"][m]))
    p.close()
    assert cwf.compare(root) == 0


def test_member_selection_syntax():
    """`-write-members` of decipher_b200_run (ensemble-scale output: text files for selected members only)."""
    assert host.parse_member_list("3:5,17,40", 40).tolist() == [2, 3, 4, 16, 39]
    assert host.parse_member_list("7,7,5:8,1", 10).tolist() == [0, 4, 5, 6, 7]
    assert host.parse_member_list("1:100000", 100000).size == 100000
    for bad in ("0:3", "5:4", "3:41", "a", "3;4", "3:", ""):
        if bad == "":
            assert host.parse_member_list(bad, 40).size == 0
            continue
        with pytest.raises(Exception):
            host.parse_member_list(bad, 40)
