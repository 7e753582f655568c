"""Generates decipher_b200/data/c2_qobs_oracle.npz: the "observed" flow of BASELINE config 2 / 3 as SURVEY §8(d)
defines it -- an ORACLE run of the same catchment with mid-range parameters, 10 % multiplicative noise, 1 % of the
values set to the missing-data code.  (The linear-reservoir proxy synth.make_raw falls back to starts half the
ensemble from the degenerate `sol_flag = 2` initialisation)

    python tests/golden/make_qobs_c2.py          # needs oracle/build/libdecipher_oracle.so

The first SPIN steps of the oracle run still carry the initialisation (routing pipeline empty, stores adjusting), so
they take the simulated values of exactly one year later (same season, same diurnal phase).  Stored as int32 multiples
of 1e-9 m/timestep = the 9 decimals the flow file is written with (formats.write_project)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from decipher_b200 import synth            # noqa: E402
from oracle_binding import oracle_run      # noqa: E402

SPIN = 24 * 60


def main():
    cfg = dict(synth.CONFIGS["C2"]); cfg.pop("members")
    raw = synth.make_raw(**cfg, qobs_fixture=None)
    cat = synth.assemble(raw)
    ll, ul = synth.example_bounds(1)
    mc = np.asfortranarray(((ll + ul) / 2.0).reshape(1, 7, 1))
    out = oracle_run(cat, mc, want_q=True, want_perf=False, want_totals=False)
    q = out["q"][:, :, 0].copy()
    assert np.isfinite(q).all() and (out["status"][0] & ~0x3) == 0, out["status"]
    q[:, :SPIN] = q[:, 8760:8760 + SPIN]
    rng = np.random.default_rng(synth.BASE_SEED + 1002)
    q = np.maximum(q * (1.0 + 0.1 * rng.standard_normal(q.shape)), 1e-9)
    qi = np.rint(q * 1e9).astype(np.int64)
    assert qi.max() < 2 ** 31
    qi = qi.astype(np.int32)
    qi[rng.uniform(size=qi.shape) < 0.01] = -1                      # -> flow_err
    path = os.path.join(ROOT, "decipher_b200", "data", "c2_qobs_oracle.npz")
    np.savez_compressed(path, qobs_e9=qi, seed=cfg["seed"], nstep=cfg["nstep"], nriv=cfg["nriv"], sol_flag=int(out["status"][0] & 3),
                        mcpar=mc[0, :, 0])
    print(path, os.path.getsize(path), "bytes; mean flow per gauge (m/h):", q.mean(axis=1), "sol_flag", out["status"][0] & 3)


if __name__ == "__main__":
    main()
