"""Standalone router: DTA/route_run_standalone.f90 -> route_processing_build_hist +
route_process_run_timeseries (DTA/dta_route_processing.f90:154-354, 583-630).

CPU: the oracle's shifted-flow-matrix restatement against an independent closed form (the FIR DESIGN.md §3.2
derives) built from the oracle's own histograms, plus linearity / unit-impulse properties.
GPU: dcp_route_timeseries through the C ABI against the oracle (same operations, same order: bit-exact)."""
import numpy as np
import pytest

from decipher_b200 import synth
from oracle_binding import oracle_build_hist, oracle_route_timeseries


def _network(tree=True, nriv=4, nstep=400, seed=7):
    return synth.synth_catchment(nac=24, nriv=nriv, npt=1, nstep=nstep, tree=tree, seed=seed)


def _inflow(cat, n, seed=3):
    rng = np.random.default_rng(seed)
    x = rng.exponential(1e-4, size=(cat.nriv, cat.nstep, n)) * (rng.random((cat.nriv, cat.nstep, n)) < 0.3)
    return np.asfortranarray(x)


def _fir_reference(cat, v, inflow):
    """q_i(t) = sum over n in {i} + upstream(i) of sum_k H(n, c_i + k) * in_n(t - k), c_i = node_hist_indexes(i,1)."""
    rc, hist, idx, T = oracle_build_hist(cat, v)
    assert rc == 0
    nriv, nstep = cat.nriv, cat.nstep
    node_of_riv = cat.node_to_flow_mapping - 1
    riv_of_node = {int(n): r for r, n in enumerate(node_of_riv)}
    up_off = np.concatenate([[0], np.cumsum(cat.uptree_count)])
    out = np.zeros((nriv, nstep))
    for i in range(nriv):
        c = idx[i, 0] - 1
        nodes = [i] + [int(u) - 1 for u in cat.uptree_index[up_off[i]:up_off[i + 1]]]
        for n in nodes:
            if n not in riv_of_node:
                continue
            taps = hist[n, c:]
            out[i] += np.convolve(inflow[riv_of_node[n]], taps)[:nstep]
    return out


@pytest.mark.parametrize("tree", [False, True])
def test_oracle_router_matches_closed_form(tree):
    cat = _network(tree=tree)
    v = np.array([300.0, 1111.1, 5000.0])
    x = _inflow(cat, len(v))
    routed, status = oracle_route_timeseries(cat, v, x)
    assert (status == 0).all()
    for s, vs in enumerate(v):
        want = _fir_reference(cat, vs, x[:, :, s])
        np.testing.assert_allclose(routed[:, :, s], want, rtol=1e-12, atol=1e-18)


def test_oracle_router_linearity_and_mass():
    """Routing is linear in the inflow, and every unit of inflow leaves through the outlet gauge exactly once
    (histograms are normalised, dta_route_processing.f90:347) when the series is long enough."""
    cat = _network(tree=True, nstep=600)
    v = np.array([800.0])
    a, b = _inflow(cat, 1, seed=1), _inflow(cat, 1, seed=2)
    a[:, 300:, :] = 0.0
    b[:, 300:, :] = 0.0
    ra, _ = oracle_route_timeseries(cat, v, a)
    rb, _ = oracle_route_timeseries(cat, v, b)
    rab, _ = oracle_route_timeseries(cat, v, np.asfortranarray(2.0 * a + b))
    np.testing.assert_allclose(rab, 2.0 * ra + rb, rtol=1e-12, atol=1e-18)
    # the outlet (the node every other node drains to) collects everything
    up_count = np.asarray(cat.uptree_count)
    outlet = int(np.argmax(up_count))
    if up_count[outlet] == cat.nnode - 1:
        np.testing.assert_allclose(ra[outlet].sum(), a.sum(), rtol=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("tree,nriv,nstep", [(False, 3, 700), (True, 5, 1500)])
def test_gpu_router_bit_exact(tree, nriv, nstep):
    from decipher_b200.capi import DeviceCatchment
    cat = _network(tree=tree, nriv=nriv, nstep=nstep, seed=11)
    rng = np.random.default_rng(5)
    n = 37
    v = rng.uniform(150.0, 6000.0, n)
    x = _inflow(cat, n)
    want, st_want = oracle_route_timeseries(cat, v, x)
    dev = DeviceCatchment(cat)
    out = dev.route_timeseries(v, x)
    dev.close()
    assert np.array_equal(out["status"], st_want)
    assert np.array_equal(out["routed"], want), float(np.abs(out["routed"] - want).max())
    assert out["stats"]["n_launches"] >= 4


@pytest.mark.gpu
def test_gpu_router_device_buffers_and_errors():
    import torch
    from decipher_b200.capi import DecipherError, DeviceCatchment
    cat = _network(tree=True, nriv=4, nstep=512, seed=2)
    n = 8
    v = np.linspace(200.0, 3000.0, n)
    x = _inflow(cat, n)
    want, _ = oracle_route_timeseries(cat, v, x)
    dev = DeviceCatchment(cat)
    d_v = torch.from_numpy(v).cuda()
    d_x = torch.from_numpy(np.ascontiguousarray(x.transpose(2, 1, 0))).cuda()      # (n, nstep, nriv) C == ABI layout
    d_out = torch.empty_like(d_x)
    d_st = torch.empty(n, dtype=torch.int32, device="cuda")
    dev.route_timeseries_device(n, d_v.data_ptr(), d_x.data_ptr(), d_out.data_ptr(), d_st.data_ptr())
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().transpose(2, 1, 0)
    assert np.array_equal(got, want)
    assert (d_st.cpu().numpy() == 0).all()
    with pytest.raises(DecipherError):
        dev.route_timeseries(np.array([0.0]), np.asfortranarray(x[:, :, :1]))
    dev.close()
