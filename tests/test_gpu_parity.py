"""GPU (C ABI) vs CPU oracle on the same seeded inputs.  Tolerance: 1e-9 relative on flows and
performance measures; sampled parameters, status words and srinit clipping bit-exact."""
import numpy as np
import pytest

from decipher_b200 import synth
from decipher_b200.capi import (DeviceCatchment, KERNEL_RESIDENT, KERNEL_STREAMED, NMEASURE, OPT_CHECK_BALANCE, OPT_FAST_MATH)
from oracle_binding import oracle_run
from parity import compare_close, compare_flows, compare_perf

pytestmark = pytest.mark.gpu


def _run_both(cat, M, flags=0, kernel=KERNEL_STREAMED, **kw):
    mc = synth.sample_parameters(cat.npt, M)
    ref = oracle_run(cat, mc.copy(order="F"), flags=flags & OPT_CHECK_BALANCE, n_threads=8)
    dev = DeviceCatchment(cat)
    out = dev.run(mc.copy(order="F"), flags=flags, kernel=kernel, want_q=True, want_perf=True, want_totals=True, **kw)
    dev.close()
    return ref, out


def _check(ref, out, rtol=1e-9, exact_status=True):
    assert np.array_equal(ref["mcpar"], out["mcpar"]), "parameter sets (incl. srinit clip) must be bit-exact"
    if exact_status:
        assert np.array_equal(ref["status"], out["status"]), (ref["status"], out["status"])
    d = compare_flows(out["q"], ref["q"], rtol)
    compare_perf(out["perf"], ref["perf"], rtol)
    fin = np.isfinite(ref["q"]).all(axis=(0, 1))
    compare_close(out["reach_totals"][:, :, fin], ref["reach_totals"][:, :, fin], rtol, "reach totals")
    compare_close(out["init_diag"][:, fin], ref["init_diag"][:, fin], rtol, "init diagnostics")
    return d


def test_c1_streamed_exact_checked():
    """Config 1: example-like project, 10 members with the shipped seeds, balance checks on."""
    cat = synth.synth_catchment("C1")
    ref, out = _run_both(cat, 10, flags=OPT_CHECK_BALANCE)
    d = _check(ref, out)
    print("C1 exact:", d)
    assert d["max_rel"] < 1e-10


def test_c1_streamed_small_tiles():
    """Time tiling must not change results (ring buffer + halo of the routing convolution)."""
    cat = synth.synth_catchment("C1", nstep=3000)
    ref, out = _run_both(cat, 10, steps_per_tile=97)
    _check(ref, out)


def test_c1_fast_math():
    cat = synth.synth_catchment("C1")
    ref, out = _run_both(cat, 10, flags=OPT_FAST_MATH)
    d = _check(ref, out)
    print("C1 fast:", d)


def test_c2_subset_streamed():
    """Config 2 catchment (500 HRUs, kW=8, 3 reaches), 2000 steps, 96 members, batches of 40."""
    cat = synth.synth_catchment("C2", nstep=2000)
    ref, out = _run_both(cat, 96, members_per_batch=40)
    d = _check(ref, out)
    print("C2 exact:", d)
    ref, out = _run_both(cat, 96, flags=OPT_FAST_MATH)
    d = _check(ref, out)
    print("C2 fast:", d)


def test_tree_network_multilayer_ntt():
    """Branching river tree, 3 parameter layers, ntt = 3 sub-steps, daily timestep (dt = 24)."""
    cat = synth.synth_catchment(nac=60, nriv=5, npt=3, ngp=3, nge=2, nstep=1500, kW=5, seed=77, tree=True, ntt=3, dt=24.0)
    ref, out = _run_both(cat, 24, flags=OPT_CHECK_BALANCE)
    _check(ref, out)


def test_device_buffers_and_stats():
    import torch
    cat = synth.synth_catchment("C1", nstep=1000)
    M = 64
    mc = synth.sample_parameters(cat.npt, M)
    ref = oracle_run(cat, mc.copy(order="F"), want_q=False, want_totals=False)
    dev = DeviceCatchment(cat)
    t_mc = torch.from_numpy(np.ascontiguousarray(mc.transpose(2, 1, 0))).cuda()     # (M, 7, npt) C-order == (npt,7,M) F-order
    t_perf = torch.empty((M, cat.nriv, NMEASURE), dtype=torch.float64, device="cuda")
    t_status = torch.empty(M, dtype=torch.int32, device="cuda")
    st = dev.run_device(t_mc.data_ptr(), M, perf_ptr=t_perf.data_ptr(), status_ptr=t_status.data_ptr())
    torch.cuda.synchronize()
    perf = t_perf.cpu().numpy().transpose(2, 1, 0)
    compare_perf(perf, ref["perf"])
    assert np.array_equal(t_status.cpu().numpy(), ref["status"])
    # AUTO -> resident kernel: slowest-velocity reduction, init, the persistent time loop (routing + objective inside), finalize
    assert st["kernel_used"] == KERNEL_RESIDENT and st["n_launches"] == 4 and st["n_physics_launches"] == 1
    assert st["ms_total"] > 0 and st["ms_h2d"] < st["ms_total"]
    dev.close()


# ---------------------------------------------------------------------------------------------
# resident (persistent, on-chip state) kernel
# ---------------------------------------------------------------------------------------------
from decipher_b200.capi import KERNEL_RESIDENT  # noqa: E402


def test_resident_c1_exact_and_fast():
    """nac=24 -> 16 member sub-groups per CTA, 3 parameter layers, ragged last group (M=70)."""
    cat = synth.synth_catchment("C1", nstep=2500)
    ref, out = _run_both(cat, 70, kernel=KERNEL_RESIDENT)
    assert out["stats"]["kernel_used"] == KERNEL_RESIDENT
    d = _check(ref, out)
    print("resident C1 exact:", d)
    assert d["max_rel"] < 1e-10
    ref, out = _run_both(cat, 70, flags=OPT_FAST_MATH, kernel=KERNEL_RESIDENT)
    print("resident C1 fast:", _check(ref, out))


def test_resident_c2_subset():
    """Config 2 catchment (500 HRUs -> one HRU per compute thread, 4 members per CTA), 1500 steps,
    150 members = 38 member groups over the persistent grid (last group ragged)."""
    cat = synth.synth_catchment("C2", nstep=1500)
    ref, out = _run_both(cat, 150, kernel=KERNEL_RESIDENT)
    assert out["stats"]["kernel_used"] == KERNEL_RESIDENT and out["stats"]["n_physics_launches"] == 1
    d = _check(ref, out)
    print("resident C2 exact:", d)
    ref, out = _run_both(cat, 150, flags=OPT_FAST_MATH, kernel=KERNEL_RESIDENT)
    print("resident C2 fast:", _check(ref, out))


def test_resident_fast_daily_timestep():
    """dt = 24 (daily series), ntt = 1: the lean step folds dt and dtt into per-member constants (td/dt, dtt*cos/szm),
    which is invisible at dt = 1; this case makes the folding visible."""
    cat = synth.synth_catchment(nac=120, nriv=3, npt=2, ngp=2, nge=1, nstep=900, kW=6, seed=23, tree=True, dt=24.0)
    ref, out = _run_both(cat, 48, flags=OPT_FAST_MATH, kernel=KERNEL_RESIDENT)
    assert out["stats"]["kernel_used"] == KERNEL_RESIDENT
    print("resident fast dt=24:", _check(ref, out))


def test_resident_matches_streamed_bitwise_in_exact_mode():
    """Both kernels evaluate the same operations in the same order: exact mode must agree to the bit."""
    cat = synth.synth_catchment(nac=100, nriv=3, npt=2, ngp=2, nge=1, nstep=700, kW=6, seed=11, tree=True)
    mc = synth.sample_parameters(cat.npt, 40)
    dev = DeviceCatchment(cat)
    a = dev.run(mc.copy(order="F"), kernel=KERNEL_STREAMED, want_q=True, want_totals=True)
    b = dev.run(mc.copy(order="F"), kernel=KERNEL_RESIDENT, want_q=True, want_totals=True)
    dev.close()
    assert np.array_equal(a["q"], b["q"], equal_nan=True)
    assert np.array_equal(a["perf"], b["perf"], equal_nan=True)
    assert np.array_equal(a["reach_totals"], b["reach_totals"], equal_nan=True)
    assert np.array_equal(a["status"], b["status"])


@pytest.mark.parametrize("nac,npt,M", [(513, 1, 9), (700, 1, 70), (1000, 1, 301), (1024, 1, 5), (600, 3, 40)])
def test_resident_wide_kernel_513_to_1024_hrus(nac, npt, M):
    """No cliff at 512 HRUs: catchments of 513 .. 1024 HRUs stay on chip (wide build of the second-generation kernel: one
    team of 1024 columns per CTA, FAST policy), AUTO picks it, results within the 1e-9 contract of the oracle; ragged
    member groups, one and several parameter layers, every column busy at nac = 1024."""
    cat = synth.synth_catchment(nac=nac, nriv=3, npt=npt, ngp=4, nge=1, nstep=900, kW=8, seed=synth.BASE_SEED + 60 + nac, tree=(npt > 1))
    ref, out = _run_both(cat, M, flags=OPT_FAST_MATH, kernel=0)
    assert out["stats"]["kernel_used"] == KERNEL_RESIDENT and out["stats"]["n_physics_launches"] == 1
    print(f"wide resident nac={nac}:", _check(ref, out))


def test_wide_w_large_catchment_leaves_the_chip():
    """Config-4-like: more HRUs than the resident kernel holds (nac = 700 > 512) and a wide weighting
    matrix (64 targets per HRU): AUTO must pick the stepwise path (the streamed kernel when the balance
    checks are requested) and stay within tolerance."""
    cat = synth.synth_catchment(nac=700, nriv=5, npt=1, ngp=4, nge=1, nstep=500, kW=64, seed=synth.BASE_SEED + 4, tree=True)
    ref, out = _run_both(cat, 48, kernel=0)
    assert out["stats"]["kernel_used"] == 3
    print("C4-like:", _check(ref, out))
    ref, out = _run_both(cat, 48, flags=OPT_CHECK_BALANCE, kernel=0)
    assert out["stats"]["kernel_used"] == KERNEL_STREAMED
    print("C4-like checked:", _check(ref, out))


def test_multi_catchment_partition():
    """Config-5-like: several catchments of different size, dealt to 2 ranks (both emulated on this GPU)."""
    from decipher_b200.ensemble import run_catchments
    cats = [synth.synth_catchment(nac=n, nriv=2, npt=1, ngp=2, nge=1, nstep=300, kW=4, seed=100 + n) for n in (30, 120, 64, 200)]
    mcs = [synth.sample_parameters(1, 12, seeds=(5 + i, 2000)) for i in range(4)]
    got = {}
    for rank in range(2):
        got.update(run_catchments(cats, mcs, rank=rank, world=2, want_q=True))
    assert sorted(got) == [0, 1, 2, 3]
    for i, c in enumerate(cats):
        ref = oracle_run(c, mcs[i].copy(order="F"))
        compare_flows(got[i]["q"], ref["q"])
        compare_perf(got[i]["perf"], ref["perf"])


# ---------------------------------------------------------------------------------------------
# stepwise path (catchments too large for the chip; dense dynamic_dist matrix as an FP64 GEMM)
# ---------------------------------------------------------------------------------------------
from decipher_b200.capi import KERNEL_STEPWISE  # noqa: E402


def _bitwise(a, b):
    for k in ("q", "perf", "reach_totals"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    assert np.array_equal(a["status"], b["status"])


@pytest.mark.parametrize("tile", [150, 77])
@pytest.mark.parametrize("min_density", ["2.0", "0.0"], ids=["csr", "gemm"])
def test_stepwise_exact_is_bitwise_streamed(min_density, tile, monkeypatch):
    """EXACT mode: redistribute (sparse CSR, or the dense GEMM with one ascending accumulation chain per output),
    physics over (HRU, member) and ordered river sums perform the reference's operations in the reference's order,
    so the result must equal the streamed kernel's bit for bit.  nac = 300 is not a multiple of the GEMM tile,
    M = 37 is odd (ragged member tile), several time tiles -- of 150 steps (four 32-step CUDA graphs + 22 direct steps per
    tile) and of 77 steps (every second tile starts on an odd step: one direct step before the graphs)."""
    monkeypatch.setenv("DCP_DENSE_MIN_DENSITY", min_density)
    cat = synth.synth_catchment(nac=300, nriv=4, npt=2, ngp=3, nge=2, nstep=400, kW=40, seed=synth.BASE_SEED + 41, tree=True)
    mc = synth.sample_parameters(cat.npt, 37)
    dev = DeviceCatchment(cat)
    a = dev.run(mc.copy(order="F"), kernel=KERNEL_STREAMED, want_q=True, want_totals=True, steps_per_tile=tile)
    b = dev.run(mc.copy(order="F"), kernel=KERNEL_STEPWISE, want_q=True, want_totals=True, steps_per_tile=tile)
    dev.close()
    assert b["stats"]["kernel_used"] == KERNEL_STEPWISE
    _bitwise(a, b)


def test_stepwise_dense_fast_vs_oracle(monkeypatch):
    """Config-4-like dense weighting matrix (every HRU drains to every down-slope HRU), FAST arithmetic
    (one FMA per term, hoisted reciprocals): within the 1e-9 contract of the oracle; AUTO picks the stepwise
    path for a catchment larger than the resident kernel holds."""
    monkeypatch.setenv("DCP_DENSE_MIN_DENSITY", "0.05")
    cat = synth.synth_catchment(nac=600, nriv=5, npt=1, ngp=4, nge=1, nstep=300, kW=600, seed=synth.BASE_SEED + 42, tree=True)
    ref, out = _run_both(cat, 40, flags=OPT_FAST_MATH, kernel=0)
    assert out["stats"]["kernel_used"] == KERNEL_STEPWISE
    print("C4-dense fast:", _check(ref, out))
    ref, out = _run_both(cat, 40, kernel=KERNEL_STEPWISE)
    print("C4-dense exact:", _check(ref, out))


def test_stepwise_nonfinite_members_match_streamed(monkeypatch):
    """Non-finite baseflows must reach exactly the HRUs the reference's sparse scatter reaches (0 * inf would
    poison every destination in a plain GEMM): the dense redistribute stages zeros, flags the member and the
    flagged members are redone in sparse form.  Two triggers: szm = 0 for two members (NaN from the first step,
    the other members stay finite), and an infinite rainfall depth on one rain grid at step 100 (every member)."""
    monkeypatch.setenv("DCP_DENSE_MIN_DENSITY", "0.0")
    cat = synth.synth_catchment(nac=200, nriv=4, npt=1, ngp=2, nge=1, nstep=300, kW=30, seed=synth.BASE_SEED + 43, tree=True)
    mc = synth.sample_parameters(1, 40)
    mc[0, 0, 3] = 0.0
    mc[0, 0, 21] = 0.0
    for inf_rain in (False, True):
        if inf_rain:
            cat.rain[1, 100] = np.inf
        dev = DeviceCatchment(cat)
        a = dev.run(mc.copy(order="F"), kernel=KERNEL_STREAMED, want_q=True, want_totals=True)
        b = dev.run(mc.copy(order="F"), kernel=KERNEL_STEPWISE, want_q=True, want_totals=True)
        dev.close()
        bad = (~np.isfinite(a["q"])).any(axis=(0, 1))
        print("non-finite members:", int(bad.sum()), "of 40")
        assert bad[3] and bad[21] and (inf_rain or bad.sum() == 2)
        _bitwise(a, b)


@pytest.mark.parametrize("kw,nstep,M", [("64", 150, 6), ("512", 60, 6), ("dense", 16, 5)])
def test_config4_full_size_5000_hrus(kw, nstep, M):
    """BASELINE config 4 at its stated size: 5 000 HRUs, kW = 64 / 512 targets per HRU or a dense weighting matrix (every
    HRU drains to every down-slope HRU: 12.5 M non-zeros).  AUTO runs it on the stepwise path (sparse CSR SpMM, or the
    FP64 GEMM over the non-empty tiles when W is dense enough); both policies against the oracle."""
    cfg = dict(synth.CONFIGS["C4"]); cfg.pop("members")
    cfg.update(nstep=nstep)
    if kw == "dense":
        cfg.update(dense_w=True)
    else:
        cfg.update(kW=int(kw))
    cat = synth.synth_catchment(**cfg)
    assert cat.nac == 5000
    for flags in (OPT_FAST_MATH, 0):
        ref, out = _run_both(cat, M, flags=flags, kernel=0)
        assert out["stats"]["kernel_used"] == KERNEL_STEPWISE
        print(f"C4 nac=5000 kW={kw} flags={flags}:", _check(ref, out))
