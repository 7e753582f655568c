"""Inputs of the executed-reference vectors (tests/golden/f90_exec/): shared by the generator (make_f90_vectors.py, needs the
reference sources) and the test (test_f90_pin_cpu.py, needs only the committed .npz files).  Catchments come from
decipher_b200.synth; parameter sets are either sampled by the reference's own RANMAR / genpar (`members`, `seeds`) or given
explicitly (`sets`, columns szm, lnto, srmax, srinit, chv, td, smax -- dyna_param_setup.f90:58-64) to drive the reference
into branches a random sample rarely reaches."""
import numpy as np

from decipher_b200 import synth

CASES = {
    "a_two_reaches": dict(raw=dict(nac=12, nriv=2, npt=3, ngp=2, nge=1, nstep=240, kW=4, seed=4101), members=4, seeds=(5, 2000)),
    "b_tree_ntt2": dict(raw=dict(nac=16, nriv=3, npt=2, ngp=2, nge=2, nstep=200, kW=5, seed=4102, tree=True, ntt=2),
                        members=3, seeds=(17, 903)),
    "c_one_reach_dt24": dict(raw=dict(nac=9, nriv=1, npt=1, ngp=1, nge=1, nstep=150, kW=3, seed=4103, dt=24.0),
                             members=3, seeds=(5, 2000)),
    # init_satzone: srinit > srmax clip (:201-204); the search oscillating down to ndiff <= 1e-9 (sol_flag 0, both exits,
    # :150-166); sd >= smax in calculate_qb (:282-285); maximum-QBF start (sol_flag 2, :120-126); very dry stores
    # (u2*dtt/m2 < 1e-10 and u2 == 0 in the solver, dyna_satzone_analyticalsolver.f90:72-86)
    "d_init_branches": dict(raw=dict(nac=10, nriv=1, npt=1, ngp=1, nge=1, nstep=50, kW=3, seed=4105),
                            sets=[[0.02, 0.0, 0.05, 0.2, 1000, 10, 1.0],
                                  [0.06, 4.6875, 0.05, 0.0, 1000, 10, 0.2],
                                  [0.06, 4.75, 0.05, 0.0, 1000, 10, 0.2],
                                  [0.07, -2.0, 0.05, 0.0, 1000, 10, 0.2],
                                  [0.001, -3.0, 0.005, 0.0, 1000, 0.1, 3.0],
                                  [0.001, 1.0, 0.005, 0.0, 1000, 0.1, 3.0]]),
    # evaporation limited by the store (dyna_root_zone1.f90:58-59): daily steps, 7.5 mm of PET every fourth day, srmax at its lower bound
    "e_high_pet": dict(raw=dict(nac=8, nriv=2, npt=2, ngp=1, nge=1, nstep=120, kW=3, seed=4107, dt=24.0), tweak="high_pet",
                       sets=[[0.02, 1.0, 0.005, 0.0, 1000, 5, 1.0, 0.03, -1.0, 0.006, 0.001, 800, 2, 0.5]]),
    # an absurd rainfall value (3e6 m in one step) makes the store-balance residuals exceed their tolerances: the reference
    # STOPs (dyna_root_zone1.f90:73-77 or dyna_results_ts.f90:88-101); a NEGATIVE value does not (the lower tolerance there is
    # the integer -1: `.lt. - 0000000001`)
    "f_huge_rain": dict(raw=dict(nac=8, nriv=1, npt=1, ngp=1, nge=1, nstep=60, kW=3, seed=4108), tweak="huge_rain",
                        members=1, seeds=(5, 2000)),
    "g_negative_rain": dict(raw=dict(nac=8, nriv=1, npt=1, ngp=1, nge=1, nstep=60, kW=3, seed=4108), tweak="negative_rain",
                            members=1, seeds=(5, 2000)),
    # a start 1e11 times drier than observed: stores begin at (or beyond) smax, upslope units receive nothing, so the solver
    # runs its u2 == 0, u2*dtt/m2 < 1e-10 and sd > smax branches (dyna_satzone_analyticalsolver.f90:72-97)
    "h_dry_start": dict(raw=dict(nac=10, nriv=1, npt=1, ngp=1, nge=1, nstep=120, kW=3, seed=4109), tweak="dry_start",
                        sets=[[0.001, 1.0, 0.005, 0.0, 1000, 0.1, 3.0],
                              [0.03, 0.0, 0.05, 0.0, 1000, 10, 0.2],
                              [0.005, 2.0, 0.02, 0.0, 2000, 1.0, 0.2],
                              [0.001, -2.0, 0.005, 0.0, 1000, 40, 0.2],
                              [0.03, 1.2910231073712835, 0.05, 0.0, 1000, 10, 0.3]]),      # reaches tc > dtt (:90-91)
    # model-structure switches other than 1 skip the zone (dyna_topmod.f90:146,161) and four sub-steps per step (ntt = 4)
    "i_zones_off_ntt4": dict(raw=dict(nac=14, nriv=2, npt=2, ngp=2, nge=1, nstep=100, kW=4, seed=4110, ntt=4), tweak="zones_off",
                             members=3, seeds=(5, 2000)),
}


def _fuzz_cases(n_cases=6, n_members=12, seed=777):
    """randomised catchments (size, reaches, parameter layers, sub-steps, daily / hourly steps, tree / chain networks) and
    parameter sets drawn uniformly from the bounds of the reference's params.dat, with srinit beyond srmax and a tiny smax
    mixed in; a run of 240 such members (40 per catchment) gave no mismatch, 72 of them are committed"""
    rng = np.random.default_rng(seed)
    out = {}
    for c in range(n_cases):
        kw = dict(nac=int(rng.integers(4, 12)), nriv=int(rng.integers(1, 4)), npt=int(rng.integers(1, 4)), ngp=2, nge=1,
                  nstep=int(rng.integers(30, 80)), kW=int(rng.integers(2, 5)), seed=5000 + c, tree=bool(c % 2),
                  ntt=int(rng.choice([1, 1, 2, 3])), dt=float(rng.choice([1.0, 1.0, 24.0])))
        kw["nriv"] = min(kw["nriv"], kw["nac"])
        ll, ul = synth.example_bounds(kw["npt"])
        sets = []
        for m in range(n_members):
            mc = ll + rng.random((kw["npt"], 7)) * (ul - ll)
            if m % 5 == 0:
                mc[:, 3] = rng.uniform(0, 0.3, kw["npt"])
            if m % 7 == 0:
                mc[:, 6] = rng.uniform(0.01, 0.2, kw["npt"])
            sets.append(mc.reshape(-1).tolist())
        out[f"z_fuzz_{c}"] = dict(raw=kw, sets=sets)
    return out


CASES.update(_fuzz_cases())


def build(name):
    """-> (Catchment, explicit parameter table (npt, 7, M) or None)"""
    c = CASES[name]
    cat = synth.assemble(synth.make_raw(**c["raw"]))
    tw = c.get("tweak")
    if tw == "high_pet":
        cat.pet = np.zeros_like(cat.pet, order="F")
        cat.pet[:, 5::4] = 0.0075
    elif tw == "negative_rain":
        cat.rain = np.asfortranarray(cat.rain.copy())
        cat.rain[0, 37] = -0.001
    elif tw == "huge_rain":
        cat.rain = np.asfortranarray(cat.rain.copy())
        cat.rain[0, 37] = 3.0e6
    elif tw == "zones_off":
        cat.ims_uz = cat.ims_uz.copy()
        cat.ims_uz[::5] = 2
        cat.ims_rz = cat.ims_rz.copy()
        cat.ims_rz[3::7] = 0
    elif tw == "dry_start":
        cat.qobs_start = cat.qobs_start * 1e-11
    mc = None
    if "sets" in c:
        npt = c["raw"]["npt"]
        mc = np.zeros((npt, 7, len(c["sets"])), order="F")
        for m, s in enumerate(c["sets"]):
            mc[:, :, m] = np.asarray(s, dtype=np.float64).reshape(npt, 7)
    return cat, mc


# ---------------------------------------------------------------------------------------------------- whole projects
# Projects in the reference's FILE formats (decipher_b200.formats.write_project): the reference's readers (Tread_dyna, Inputs,
# mc_setup, modelstruct_setup, read_routingdata with riv_tree_read_impl / read_numeric_list_fid) are executed on the files, then
# ranin / genpar / mainloop for `numsim` members -- `program main` (dyna_main.f90:95-206), `project` included.
PROJECTS = {
    "p_tree": dict(raw=dict(nac=14, nriv=3, npt=2, ngp=2, nge=2, nstep=50, kW=4, seed=77, tree=True), numsim=3, seeds=(11, 222)),
    # observed flow: gauge 1 misses its first value (start = mean of the valid ones, dyna_inputs.f90:96-99), gauge 2 has no valid
    # value at all (start = (0.001/24)*dt, the quotient taken in SINGLE precision, :94)
    "p_missing_flow": dict(raw=dict(nac=10, nriv=2, npt=1, ngp=1, nge=1, nstep=40, kW=3, seed=78), numsim=2, seeds=(5, 2000),
                           tweak="missing_flow"),
}


def write_project(name, root):
    """writes the project files under root; -> raw dict"""
    from decipher_b200 import formats
    c = PROJECTS[name]
    raw = synth.make_raw(**c["raw"])
    if c.get("tweak") == "missing_flow":
        raw["qobs"] = np.asfortranarray(raw["qobs"].copy())
        raw["qobs"][0, 0] = raw["flow_err"]
        raw["qobs"][1, :] = raw["flow_err"]
    ll, ul = synth.example_bounds(c["raw"]["npt"])
    formats.write_project(raw, root, ll, ul, numsim=c["numsim"], seeds=c["seeds"])
    return raw


def project_units(root, name="Catch1"):
    """unit number -> file, as `project` connects them (dyna_project.f90:247-320)"""
    import os
    inp, st = os.path.join(root, "INPUTS"), os.path.join(root, "SETTINGS")
    return {12: f"{st}/params.dat", 11: f"{st}/model_structure.dat", 13: f"{inp}/{name}_dyna_hru.dat", 17: f"{inp}/{name}_hru_meta.dat",
            501: f"{inp}/{name}_flow_conn.dat", 502: f"{inp}/{name}_riv_data.dat", 503: f"{inp}/{name}_flow_point.dat",
            14: f"{inp}/{name}_obsq.dat", 15: f"{inp}/{name}_PET.dat", 16: f"{inp}/{name}_rain.dat"}


# ---------------------------------------------------------------------------------------------------- the reference's own control files
REF_EXAMPLE_IDS = [(1, 1, 1), (2, 3, 2)]          # (project, catchment setup, input setup) triples that are executed


def stage_reference_example(work, gold):
    """work/project.dat (the reference's text, its two absolute project_dir values pointed at work/CatchmentN/) plus, per
    project, SETTINGS/ holding the reference's params.dat, model_structure.dat and settings.dat BYTE FOR BYTE (copies under
    `gold` = tests/golden/ref_example) and INPUTS/ with synthetic data under exactly the names settings.dat lists (the data
    files are not shipped upstream).  -> (dirs, raws)"""
    import os
    import shutil
    from decipher_b200 import formats
    ll, ul = synth.example_bounds(3)
    dirs = []
    raws = {}
    for ic, cname in enumerate(["Catchment1", "Catchment2"]):
        root = os.path.join(work, cname)
        # three different HRU discretisations and three forcing sets, as settings.dat describes
        for k in range(3):
            raw = synth.make_raw(nac=12 + 6 * k + ic, nriv=2, npt=3, ngp=2, nge=1, nstep=120 + 24 * k, kW=3, seed=77 + 10 * ic + k)
            raws[(ic, k)] = raw
            tmp = os.path.join(work, f"tmp_{ic}_{k}")
            formats.write_project(raw, tmp, ll, ul, name="X")
            os.makedirs(os.path.join(root, "INPUTS"), exist_ok=True)
            inp = os.path.join(tmp, "INPUTS")
            shutil.copy(os.path.join(inp, "X_dyna_hru.dat"), os.path.join(root, "INPUTS", f"Catch1_HRU{k + 1}_dyna_hru.dat"))
            shutil.copy(os.path.join(inp, "X_hru_meta.dat"), os.path.join(root, "INPUTS", f"Catch1_HRU{k + 1}_hru_meta.dat"))
            if k == 0:      # the three catchment setups share one river network (settings.dat:14-16, 21-23, 28-30)
                for n in ("flow_conn", "riv_data", "flow_point"):
                    shutil.copy(os.path.join(inp, f"X_{n}.dat"), os.path.join(root, "INPUTS", f"Catch1_{n}.dat"))
            tag = ["10k", "lump", "lump95"][k]
            shutil.copy(os.path.join(inp, "X_obsq.dat"), os.path.join(root, "INPUTS", f"Catch1_{tag}_obsq.dat"))
            shutil.copy(os.path.join(inp, "X_rain.dat"), os.path.join(root, "INPUTS", f"Catch1_{tag}_rain.dat"))
            shutil.copy(os.path.join(inp, "X_PET.dat"), os.path.join(root, "INPUTS", f"Catch1_{tag}_PET.dat"))
            shutil.rmtree(tmp)
        os.makedirs(os.path.join(root, "SETTINGS"), exist_ok=True)
        os.makedirs(os.path.join(root, "OUTPUT"), exist_ok=True)
        for f in ("params.dat", "model_structure.dat", "settings.dat"):
            shutil.copyfile(os.path.join(gold, f), os.path.join(root, "SETTINGS", f))          # verbatim
        dirs.append(root + "/")
    text = open(os.path.join(gold, "project.dat")).read()
    assert "/data/MODEL/Catchment1/" in text and "/data/MODEL/Catchment2/" in text
    text = text.replace("/data/MODEL/Catchment1/", dirs[0]).replace("/data/MODEL/Catchment2/", dirs[1])
    with open(os.path.join(work, "project.dat"), "w") as f:
        f.write(text)
    return dirs, raws


# ---------------------------------------------------------------------------------------------------- standalone router program
# DTA/route_run_standalone.f90 executed on the file set it reads (decipher_b200.formats.write_router_inputs).  Its velocity is
# hard-wired (v_param(1) = 0.133 per timestep, :53), so the synthetic reaches are metres long to keep the delays within the series.
ROUTER_VELOCITY = 0.133


def router_network(sea):
    """-> (Catchment describing the river network, inflow (n_points, nstep))"""
    import dataclasses
    cat = synth.synth_catchment(nac=24, nriv=4, npt=1, nstep=90, tree=True, seed=4401, reach_km=(0.002, 0.012), cell_m=0.05)
    if sea:                       # an ungauged sea outlet as first node (DTA/route_run_standalone.f90:139-151): one input column less than nodes
        reach_m, cell_m = 3.0, 0.05
        down_old = np.asarray(cat.meta["down"])
        down = np.concatenate([[0], np.where(down_old > 0, down_old + 1, 1)]).astype(np.int32)
        _, tree = synth.full_upstream(down)
        nc = int(round(reach_m / cell_m))
        rd = cat.river_data.copy()
        rd[:, 2] += reach_m
        rd[:, 6] += reach_m
        seac = np.zeros((nc, 8))
        seac[:, 0] = 900.0
        seac[:, 1] = 2500.0
        seac[:, 6] = np.arange(nc) * cell_m
        seac[:, 2] = seac[:, 6] + cell_m
        seac[:, 3] = np.arange(nc) * cell_m
        seac[:, 5] = 0.002
        cat = dataclasses.replace(
            cat, node_gauge_id=np.concatenate([[900], cat.node_gauge_id]).astype(np.int32),
            node_type=np.concatenate([[1], cat.node_type]).astype(np.int32),
            uptree_count=np.array([len(t) for t in tree], dtype=np.int32),
            uptree_index=np.array([u for t in tree for u in t], dtype=np.int32),
            frac_sub_area=np.concatenate([[1.0], cat.frac_sub_area]),
            river_data=np.asfortranarray(np.round(np.vstack([seac, rd]), 3)),
            node_to_flow_mapping=(np.arange(cat.nriv) + 2).astype(np.int32), meta=dict(down=down))
    else:
        cat = dataclasses.replace(cat, river_data=np.asfortranarray(np.round(cat.river_data, 3)))
    rd = cat.river_data.copy()
    rd[:, 1] = 20.0 + (np.arange(rd.shape[0]) % 7)         # cell areas (the 5 cm cells of this toy network round to 0 otherwise)
    cat = dataclasses.replace(cat, river_data=np.asfortranarray(rd))
    rng = np.random.default_rng(4402)
    x = rng.exponential(1e-1, size=(cat.nriv, cat.nstep)) * (rng.random((cat.nriv, cat.nstep)) < 0.3)
    return cat, np.asfortranarray(x)
