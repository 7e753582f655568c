"""The multi-rank entry point of the C ABI (dcp_comm_init / dcp_run_ensemble_multi): one host thread per rank, contiguous
member blocks of ONE parameter table, one gather of the measures at the end.  On a one-GPU box the ranks share device 0
(the gather then uses device copies instead of NCCL); with >= 2 GPUs visible the same test also runs over NCCL."""
import numpy as np
import pytest

from decipher_b200 import synth
from decipher_b200.capi import Communicator, DeviceCatchment, KERNEL_AUTO, OPT_FAST_MATH, load_library

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_ranks", [1, 2, 3])
def test_multi_rank_run_equals_single_handle_bit_for_bit(n_ranks):
    cat = synth.synth_catchment("C2", nstep=600)
    M = 101                                          # ragged blocks
    table = synth.sample_parameters(cat.npt, M)
    dev = DeviceCatchment(cat)
    want = dev.run(table.copy(order="F"), flags=OPT_FAST_MATH, kernel=KERNEL_AUTO, want_q=True, want_totals=True)
    dev.close()
    comm = Communicator(n_ranks, devices=[0] * n_ranks)
    assert not comm.uses_nccl
    got = comm.run(cat, table.copy(order="F"), flags=OPT_FAST_MATH)                       # measures only: device buffers + gather
    for k in ("perf", "init_diag", "status", "mcpar"):
        assert np.array_equal(got[k], want[k], equal_nan=True), k
    st = got["stats"]
    assert st["n_ranks"] == n_ranks and st["ms_busy_max"] > 0 and len(st["ms_rank_device"]) == n_ranks
    got2 = comm.run(cat, table.copy(order="F"), flags=OPT_FAST_MATH, want_q=True, want_totals=True)   # flows: straight to the caller's arrays
    for k in ("q", "perf", "reach_totals", "init_diag", "status", "mcpar"):
        assert np.array_equal(got2[k], want[k], equal_nan=True), k
    comm.close()


def test_multi_rank_over_nccl_when_two_gpus_are_visible():
    if load_library().dcp_device_count() < 2:
        pytest.skip("one GPU visible: the NCCL leg needs two")
    cat = synth.synth_catchment("C1", nstep=500)
    table = synth.sample_parameters(cat.npt, 64)
    dev = DeviceCatchment(cat)
    want = dev.run(table.copy(order="F"))
    dev.close()
    comm = Communicator(2)
    assert comm.uses_nccl
    got = comm.run(cat, table.copy(order="F"))
    assert got["stats"]["used_nccl"]
    for k in ("perf", "init_diag", "status", "mcpar"):
        assert np.array_equal(got[k], want[k], equal_nan=True), k
    comm.close()


def test_ensemble_driver_all_devices_with_signatures():
    """decipher_b200.ensemble.run_members_all_devices (two ranks on device 0 here) incl. the flow-duration signatures,
    which every rank writes straight into its block of the caller's table."""
    from decipher_b200.ensemble import run_members_all_devices
    cat = synth.synth_catchment(nac=700, nriv=3, npt=1, ngp=4, nge=1, nstep=500, kW=8, seed=4242)      # wide resident kernel
    table = synth.sample_parameters(cat.npt, 33)
    dev = DeviceCatchment(cat)
    want = dev.run(table.copy(order="F"), flags=OPT_FAST_MATH, want_sig=True)
    dev.close()
    got = run_members_all_devices(cat, table.copy(order="F"), devices=[0, 0], flags=OPT_FAST_MATH, want_sig=True)
    assert np.array_equal(got["perf"], want["perf"], equal_nan=True) and got["stats"]["kernel_used"] == 2
    assert np.array_equal(got["sig"], want["sig"], equal_nan=True)
    assert np.isfinite(want["sig"][:3]).all() and (want["sig"][0] >= want["sig"][2]).all()


def test_communicator_rejects_bad_devices():
    from decipher_b200.capi import DecipherError
    with pytest.raises(DecipherError):
        Communicator(2, devices=[0, 99])
