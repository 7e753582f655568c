"""The oracle against the reference's OWN source, executed.

`tests/golden/f90_exec/*.npz` were produced by `tests/golden/make_f90_vectors.py`: the reference's Fortran files (RRMODEL/
dyna_*.f90, DTA/dta_route_processing.f90), translated statement by statement by `oracle/f90_translate.py` and run with
Fortran arithmetic semantics on small synthetic catchments -- RANMAR and genpar, param_setup, init_satzone (every exit of
its search), Topmod with its three zones, dynamic_dist, river / route_process_run_step, results_ts, write_output.  The C++
oracle has to reproduce every flow, total, parameter and flag BIT FOR BIT; the GPU kernels are then held to the oracle
(tests/test_gpu_*.py) and to these vectors directly (tests/test_gpu_f90_pin.py).

When the reference sources are present (this container, not the GPU box) one case is regenerated live, which keeps the
generator and the committed files honest.
"""
import ast
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import f90_cases  # noqa: E402
from decipher_b200 import capi, synth  # noqa: E402
from oracle_binding import oracle_build_hist, oracle_genpar, oracle_ranmar, oracle_route_timeseries, oracle_run  # noqa: E402

VEC = os.path.join(HERE, "golden", "f90_exec")


def bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.uint64)


def same_bits(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(bits(a), bits(b))


def load(name):
    return np.load(os.path.join(VEC, name + ".npz"), allow_pickle=False)


def test_ranmar_stream_of_the_executed_source():
    """ranin(5, 2000) + ranaru (RRMODEL/dyna_random.f90), the seeds of RRMODEL_EXAMPLE_FILES/params.dat"""
    u = load("ranmar_5_2000")["u"]
    assert same_bits(oracle_ranmar(5, 2000, 0, 200), u)
    assert 0.0 <= u.min() and u.max() < 1.0


@pytest.mark.parametrize("name", sorted(f90_cases.CASES))
def test_oracle_reproduces_the_executed_reference(name):
    case = f90_cases.CASES[name]
    v = load(name)
    cat, explicit = f90_cases.build(name)
    mc_in = v["mcpar_in"]
    if explicit is None:                                           # sampled by the reference's genpar: bit-exact parameter sets
        ll, ul = synth.example_bounds(case["raw"]["npt"])
        assert same_bits(oracle_genpar(case["seeds"][0], case["seeds"][1], ll, ul, case["members"]), mc_in)
    else:
        assert same_bits(explicit, mc_in)
    M = mc_in.shape[2]
    ref = oracle_run(cat, np.asfortranarray(mc_in).copy(order="F"), flags=capi.OPT_CHECK_BALANCE)
    stop = [str(s) for s in v["stop"]]
    warn = ast.literal_eval(str(v["warnings"]))
    for m in range(M):
        st = int(ref["status"][m])
        assert (st & capi.ST_SOLFLAG_MASK) == int(v["sol_flag"][m]), (m, st)
        # the three numbers init_satzone writes to the .res file (x 1000, dyna_init_satzone.f90:188-190)
        assert same_bits(ref["init_diag"][:, m] * 1000, v["init"][:, m]), (m, ref["init_diag"][:, m] * 1000, v["init"][:, m])
        rz = "dyna_root_zone1" in stop[m]
        ts = "dyna_results_ts" in stop[m]
        assert bool(st & capi.ST_RZ_BALANCE) == rz, (m, st, stop[m])
        assert bool(st & (capi.ST_TS_OUTSUM | capi.ST_TS_WBAL)) == ts, (m, st, stop[m])
        assert (" Storebal SRZ wrong " in warn[m]) == rz
        if stop[m]:
            continue                                               # the reference program ends there: nothing was written
        assert same_bits(ref["q"][:, :, m], v["q"][:, :, m]), (m, np.abs(ref["q"][:, :, m] - v["q"][:, :, m]).max())
        assert same_bits(ref["reach_totals"][:, :, m], v["totals"][:, :, m]), m
        assert same_bits(ref["mcpar"][:, :, m], v["mcpar_written"][:, :, m]), m       # after the srinit clip
        assert np.isfinite(v["q"][:, :, m]).all() == (not (st & capi.ST_NONFINITE))


def test_branches_the_vectors_reach():
    """the explicit parameter sets exist to drive the reference into its rare branches: make sure they still do"""
    d = load("d_init_branches")
    assert sorted(set(d["sol_flag"].tolist())) == [0, 1, 2]
    assert d["mcpar_written"][0, 3, 0] == d["mcpar_written"][0, 2, 0] != d["mcpar_in"][0, 3, 0]      # srinit clipped to srmax
    assert "dyna_root_zone1.f90" in str(load("f_huge_rain")["stop"][0])
    assert str(load("g_negative_rain")["stop"][0]) == ""            # `.lt. - 0000000001` is the integer -1
    w = ast.literal_eval(str(load("b_tree_ntt2")["warnings"]))
    assert all(x.get(" Storebal SD wrong KINEMATIC ", 0) > 0 for x in w)      # the reference's own check misfires for ntt = 2


def test_routing_of_the_executed_source():
    """hist_add_value on the inputs of the reference's add_accumulation_test (dta_route_processing.f90:444-468),
    route_processing_build_hist at the four velocities of build_hist_test (:470-534) and route_process_run_timeseries"""
    v = load("routing")
    acc = v["accumulation"]
    # the reference's own comment (:392-397) describes the split: first cell (1 - fraction), full middle cells, last cell
    assert same_bits(acc[0], [0, 0, 1.0 * (2.4 - 2.2), 0, 0, 0])
    assert same_bits(acc[1], [0, 0, 1 - (2.2 - 3 + 1), 3.1 - 4 + 1, 0, 0])
    assert same_bits(acc[2], [0, 0, 1 - (2.2 - 3 + 1), 1.0, 4.2 - 5 + 1, 0])
    assert same_bits(acc[3], [0, 0, 1 - (2.2 - 3 + 1), 1.0, 1.0, 5.2 - 6 + 1])
    raw = synth.make_raw(nac=10, nriv=4, npt=1, ngp=1, nge=1, nstep=96, kW=3, seed=int(v["route_seed"][0]), tree=True)
    cat = synth.assemble(raw)
    inflow = v["inflow"]                                            # (timestep, series)
    for k, vel in enumerate(v["velocity"]):
        rc, hist, idx, T = oracle_build_hist(cat, float(vel))
        assert rc == 0 and T == int(v["hist_len"][k])
        assert same_bits(hist, v[f"hist_{k}"])
        assert np.array_equal(idx, v["hist_idx"][k])
        routed, st = oracle_route_timeseries(cat, np.array([vel]), np.asfortranarray(inflow.T[:, :, None]))
        assert same_bits(routed[:, :, 0], v["routed"][k].T) and st[0] == 0


def _check_against_executed_program(p, v, work):
    """p: host.Project (C++ reader + sampler); v: what `program main` of the reference produced on the same files"""
    cat = p.catchment()
    assert (cat.nac, cat.nriv, cat.nstep, cat.npt, cat.ntt) == tuple(int(v[k]) for k in ("nac", "nriv", "nstep", "npt", "ntt"))
    assert cat.dt == float(v["dt"]) and p.numsim == int(v["numsim"]) and tuple(p.seeds) == tuple(v["seeds"])
    for k in ("ac", "st", "mslp", "wtrans", "rivers", "sum_ac_riv", "qobs_start", "rain", "pet", "frac_sub_area", "river_data"):
        assert same_bits(getattr(cat, k), v[k]), k
    for k in ("ippt", "ipet", "ipar", "ipriv", "ims_rz", "ims_uz", "ntrans", "itrans", "node_gauge_id", "node_type",
              "uptree_count", "uptree_index"):
        assert np.array_equal(getattr(cat, k), v[k]), k
    assert np.array_equal(cat.node_to_flow_mapping, v["mapping"])
    assert same_bits(p.ll, v["ll"]) and same_bits(p.ul, v["ul"])
    # `project` itself (dyna_project.f90): the output prefix and the files it connects to units 11-17 and 501-503
    assert os.path.normpath(p.out_prefix) == os.path.normpath(os.path.join(work, str(v["out_dir_path"]).replace("<work>", work)))
    # names of the per-member output files (dyna_file_open.f90: "('mc_id_',I8.8)")
    for m in range(p.numsim):
        for ext, key in ((".res", "res_files"), (".flow", "flow_files")):
            want = os.path.join(work, str(v[key][m]).replace("<work>", work))
            assert os.path.normpath(p.out_prefix + "mc_id_%08d%s" % (m + 1, ext)) == os.path.normpath(want)
    mc = p.genpar(p.numsim)
    assert same_bits(mc, v["mcpar"])
    ref = oracle_run(cat, mc.copy(order="F"))
    assert same_bits(ref["q"], v["q"]) and same_bits(ref["reach_totals"], v["totals"])
    assert same_bits(ref["q"], v["q_program"])           # ... as written by the reference's unmodified `program main` (dyna_main.f90)
    return cat


@pytest.mark.parametrize("name", sorted(f90_cases.PROJECTS))
def test_host_reader_sampler_and_oracle_reproduce_the_executed_program(name, tmp_path):
    """`program main` of the reference executed on project FILES (its own `project`, Tread_dyna, Inputs, mc_setup,
    modelstruct_setup, read_routingdata / riv_tree_read_impl, then ranin / genpar / mainloop) against the drop-in chain on
    the same files: the C++ host reader, its RANMAR / genpar and the oracle.  Every array the readers leave behind and every
    flow: bit for bit."""
    from decipher_b200 import host
    v = load(name)
    root = str(tmp_path)
    f90_cases.write_project(name, root)
    p = host.Project(root)
    _check_against_executed_program(p, v, root)
    p.close()
    want = f90_cases.project_units(root)
    got = dict(zip(v["unit_numbers"].tolist(), [str(x) for x in v["unit_files"]]))
    assert {u: os.path.normpath(os.path.join(root, f)) for u, f in got.items() if u != 10} == {u: os.path.normpath(f) for u, f in want.items()}
    if name == "p_missing_flow":
        # (0.001/24) is a single-precision quotient (dyna_inputs.f90:94): not 0.001/24 in double
        assert v["qobs_start"][1] == float(np.float32(0.001) / np.float32(24)) * float(v["dt"]) != (0.001 / 24) * float(v["dt"])


@pytest.mark.parametrize("ids", f90_cases.REF_EXAMPLE_IDS)
def test_the_references_own_control_files_executed_by_its_own_source(ids, tmp_path):
    """RRMODEL_EXAMPLE_FILES/{project,settings,params,model_structure}.dat -- the only fixtures the reference ships for this
    path, staged verbatim (tests/golden/ref_example/) -- read by the reference's own `project` / `mc_setup` /
    `modelstruct_setup` ... executed from source, then 10 members through `mainloop`; the C++ host on the same staging must
    arrive at the same setup selection, bounds, seeds, sampled parameter sets and flows."""
    from decipher_b200 import host
    v = load("x_ref_example_%d_%d_%d" % tuple(ids))
    work = str(tmp_path)
    f90_cases.stage_reference_example(work, os.path.join(HERE, "golden", "ref_example"))
    p = host.Project(work, *ids)
    _check_against_executed_program(p, v, work)
    assert p.numsim == 10 and tuple(p.seeds) == (5, 2000)            # params.dat:1-2 as the reference itself parsed them
    files = [str(x) for x in v["unit_files"]]
    assert any(f.endswith("Catch1_HRU%d_dyna_hru.dat" % ids[1]) for f in files)        # settings.dat's setup selection
    assert any(f.endswith("Catch1_%s_obsq.dat" % ["10k", "lump", "lump95"][ids[2] - 1]) for f in files)
    p.close()


@pytest.mark.parametrize("sea", [False, True])
def test_router_front_end_against_the_executed_router_program(sea, tmp_path):
    """The main program DTA/route_run_standalone.f90 executed from source on the file set it reads (tree, river data, ASCII-grid
    series; with and without an ungauged sea outlet as first node) against the C++ front end + the oracle's router: the
    column -> node mapping, the headers of the output file, its name, and every routed value bit for bit."""
    from decipher_b200 import formats, host
    v = load("r_router_program_%s" % ("sea" if sea else "plain"))
    cat, x = f90_cases.router_network(sea)
    paths = formats.write_router_inputs(cat, x, str(tmp_path))
    r = host.RouterInputs(paths["river_data"], paths["flow_conn"], paths["timeseries"])
    net = r.network()
    assert np.array_equal(net["node_to_flow_mapping"], np.arange(cat.nriv) + (2 if sea else 1))      # route_run_standalone.f90:139-151
    assert same_bits(r.inflow, x)
    routed, st = oracle_route_timeseries(cat, np.array([f90_cases.ROUTER_VELOCITY]), np.asfortranarray(r.inflow[:, :, None]))
    assert st[0] == 0 and same_bits(routed[:, :, 0].T, v["routed"])
    assert int(v["precision"]) == 6 and str(v["out_file"]) == "net_flow_routed.txt"
    out = str(tmp_path / "routed.txt")
    r.write_routed(routed[:, :, 0], out)
    assert open(out).readline().rstrip("\n").split("\t") == [str(h) for h in v["headers"]]
    if sea:
        # the reference stores the COUNTER in flow_output_indexes (dta_route_processing.f90:141-148): column k reads node k,
        # so the first column is the SEA node's flow under the first gauge's header -- reproduced, not corrected
        assert int(cat.node_type[0]) == 1 and [str(h) for h in v["headers"]] == [str(g) for g in cat.node_gauge_id[1:]]
    r.close()


def _reference_present():
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_f90_vectors
    return make_f90_vectors.reference_available()


@pytest.mark.skipif(not _reference_present(), reason="reference sources are not on this machine (GPU box)")
def test_live_translation_reproduces_the_committed_vectors():
    import make_f90_vectors as mk
    ns = mk.load_reference()
    assert not [p for p in mk.NEEDED if p in ns["_NOT_TRANSLATED"]]
    import tempfile
    with tempfile.TemporaryDirectory() as root:
        f90_cases.write_project("p_tree", root)
        r = mk.run_project(ns, "p_tree", root)            # includes a run of the unmodified main program (run_main_program)
        v = load("p_tree")
        for k in ("ac", "wtrans", "qobs_start", "frac_sub_area", "mcpar", "q", "totals"):
            assert same_bits(r[k], v[k]), k
    from decipher_b200 import formats
    with tempfile.TemporaryDirectory() as root:
        cat, x = f90_cases.router_network(True)
        r = mk.run_router_program(ns, root, formats.write_router_inputs(cat, x, root))
        assert same_bits(r["routed"], load("r_router_program_sea")["routed"])
    for name in ("a_two_reaches", "d_init_branches"):
        v = load(name)
        cat, mc, res = mk.make_case(ns, name)
        assert same_bits(mc, v["mcpar_in"])
        for m, r in enumerate(res):
            assert same_bits(r["q"], v["q"][:, :, m]) and same_bits(r["totals"], v["totals"][:, :, m])
            assert r["sol_flag"] == int(v["sol_flag"][m])


def test_statement_coverage_of_the_executed_source():
    """coverage.json (make_f90_vectors.py --coverage): which statements of the reference the vectors run.  The process
    equations must be covered completely; what is left are error exits and diagnostics."""
    import json
    with open(os.path.join(VEC, "coverage.json")) as f:
        cov = json.load(f)
    for proc in ("param_setup", "initialise_run", "init_satzone", "calculate_qb", "calculate_qbmax", "rain", "dynamic_dist",
                 "root_zone1", "satzone_analyticalsolver", "river", "topmod", "mainloop", "write_output", "genpar", "ranin", "ranaru",
                 "riv_node_full_upstream", "list_add_item_reserve", "read_routingdata", "read_numeric_list_fid"):
        assert cov[proc]["statements"] > 0 and cov[proc]["not_executed"] == [], (proc, cov[proc]["not_executed"])
    assert cov["unsat_zone1"]["not_executed"] == ["dyna_unsat_zone1.f90:85"]            # the " Storebal SUZ wrong " print
    assert cov["results_ac"]["not_executed"] == ["dyna_results_ac.f90:31"]              # the " Problem with EX " print
    assert cov["_total"]["executed"] >= 0.88 * cov["_total"]["statements"]
