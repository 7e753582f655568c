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
                              [0.001, -2.0, 0.005, 0.0, 1000, 40, 0.2]]),
}


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
    elif tw == "dry_start":
        cat.qobs_start = cat.qobs_start * 1e-11
    mc = None
    if "sets" in c:
        npt = c["raw"]["npt"]
        mc = np.zeros((npt, 7, len(c["sets"])), order="F")
        for m, s in enumerate(c["sets"]):
            mc[:, :, m] = np.asarray(s, dtype=np.float64).reshape(npt, 7)
    return cat, mc
