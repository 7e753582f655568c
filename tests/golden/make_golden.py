"""Generates tests/golden/c1_small.npz and tree_ntt2.npz: outputs of the CPU oracle on two seeded cases with observed flow,
i.e. including the performance measures (NSE, log-NSE, KGE, PBIAS, RMSE), which do not exist upstream.

These two files are ORACLE-made regression targets.  The link to the reference itself is tests/golden/f90_exec/ (written by
make_f90_vectors.py from the reference's own source, executed): the oracle reproduces those bit for bit
(tests/test_f90_pin_cpu.py), so what these goldens add is the objective functions and a committed target for the GPU tests at
a somewhat larger size.  Run: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from decipher_b200 import synth  # noqa: E402
from oracle_binding import oracle_run  # noqa: E402

CASES = {
    "c1_small": dict(cfg=dict(nac=24, nriv=2, npt=3, ngp=2, nge=1, nstep=600, kW=4, seed=synth.BASE_SEED + 1), members=6),
    "tree_ntt2": dict(cfg=dict(nac=40, nriv=4, npt=2, ngp=2, nge=2, nstep=400, kW=5, seed=4242, tree=True, ntt=2, dt=24.0), members=5),
}


def build(name):
    c = CASES[name]
    cat = synth.synth_catchment(**c["cfg"])
    mc = synth.sample_parameters(cat.npt, c["members"])
    return cat, mc


if __name__ == "__main__":
    for name in CASES:
        cat, mc = build(name)
        r = oracle_run(cat, mc.copy(order="F"), flags=1)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), q=r["q"], perf=r["perf"], reach_totals=r["reach_totals"],
                            init_diag=r["init_diag"], status=r["status"], mcpar_out=r["mcpar"], mcpar_in=mc)
        print(name, r["q"].shape, "status", r["status"])
