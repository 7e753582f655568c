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


def _with_sea_outlet(cat, reach_m=4000.0, cell_m=50.0):
    """The network `route_run_standalone` also accepts (DTA/route_run_standalone.f90:139-151): a first node that is an
    UNGAUGED sea outlet (node type 1) below the old outlet.  The flow input then has no column for it and column i feeds
    node i + 1; outputs stay `flow_output_indexes(k) = k` (route_processing_init, dta_route_processing.f90:141-148, counts the
    gauge nodes but stores the COUNTER, not the node index): output k reads node k -- the sea node first, the last gauge
    never.  That is the reference's behaviour and it is reproduced, not corrected."""
    import dataclasses
    nn = cat.nnode
    down_old = np.asarray(cat.meta["down"])
    down = np.concatenate([[0], np.where(down_old > 0, down_old + 1, 1)]).astype(np.int32)      # old outlet drains to the sea node
    _, tree = synth.full_upstream(down)
    nc = int(round(reach_m / cell_m))
    rd = cat.river_data.copy()
    rd[:, 2] += reach_m; rd[:, 6] += reach_m                              # dist, ds_dist of every old cell move up by the new reach
    sea = np.zeros((nc, 8))
    sea[:, 0] = 900.0                                                     # riv_id of the sea node
    sea[:, 1] = 2500.0                                                    # cell area
    sea[:, 6] = np.arange(nc) * cell_m                                    # ds_dist
    sea[:, 2] = sea[:, 6] + cell_m                                        # dist
    sea[:, 3] = np.arange(nc) * cell_m                                    # section_dist
    sea[:, 5] = 0.002
    return dataclasses.replace(
        cat, node_gauge_id=np.concatenate([[900], cat.node_gauge_id]).astype(np.int32),
        node_type=np.concatenate([[1], cat.node_type]).astype(np.int32),
        uptree_count=np.array([len(t) for t in tree], dtype=np.int32),
        uptree_index=np.array([u for t in tree for u in t], dtype=np.int32),
        frac_sub_area=np.concatenate([[1.0], cat.frac_sub_area]),
        river_data=np.asfortranarray(np.vstack([sea, rd])),
        node_to_flow_mapping=(np.arange(cat.nriv) + 2).astype(np.int32),
        meta=dict(down=down))


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


@pytest.mark.parametrize("sea", [False, True])
def test_router_front_end_reads_the_reference_file_set(sea, tmp_path):
    """The C++ host parses what DTA/route_run_standalone.f90 reads -- *_flow_conn.txt (+ the *_flow_point.txt beside it),
    *_river_data.txt and the ASCII-grid flow series -- into the network the generator wrote, maps the series' columns to nodes
    (one column less than nodes = ungauged sea outlet first: column i feeds node i + 1) and writes the routed series like
    write_numeric_list."""
    from decipher_b200 import formats, host
    cat = _network(tree=True, nriv=4, nstep=60, seed=5)
    if sea:
        cat = _with_sea_outlet(cat)
    x = _inflow(cat, 1)[:, :, 0]
    paths = formats.write_router_inputs(cat, x, str(tmp_path))
    r = host.RouterInputs(paths["river_data"], paths["flow_conn"], paths["timeseries"])
    net = r.network()
    assert net["nnode"] == cat.nnode and net["nriv"] == cat.nriv and net["nstep"] == cat.nstep
    for k in ("node_gauge_id", "node_type", "uptree_count", "uptree_index", "node_to_flow_mapping"):
        assert np.array_equal(net[k], getattr(cat, k)), k
    assert np.array_equal(net["river_data"], cat.river_data) and np.array_equal(r.inflow, x)
    out = str(tmp_path / "routed.txt")
    r.write_routed(x * 1000.0, out)
    lines = open(out).read().splitlines()
    assert lines[0].split("\t") == [str(g) for g, t in zip(cat.node_gauge_id, cat.node_type) if t == 2]
    vals = np.array([[float(v) for v in ln.split(" ")] for ln in lines[1:]]).T
    assert vals.shape == x.shape and np.abs(vals - x * 1000.0).max() <= 0.5e-6 + 1e-12
    assert not any(v.startswith("0.") for ln in lines[1:] for v in ln.split(" "))          # gfortran F0.6: ".123456"
    r.close()
    with pytest.raises(Exception):
        host.RouterInputs(paths["river_data"], paths["flow_conn"], paths["river_data"])      # not an ASCII grid


@pytest.mark.gpu
@pytest.mark.parametrize("sea", [False, True])
def test_router_cli_matches_oracle(sea, tmp_path):
    """decipher_route_run (the reference executable's interface) end to end: files -> GPU router -> *_routed.txt, against the
    oracle at the 6 decimals write_numeric_list prints; three velocities in one call."""
    import os
    import subprocess
    from decipher_b200 import formats
    cat = _network(tree=True, nriv=4, nstep=300, seed=9)
    if sea:
        cat = _with_sea_outlet(cat)
    x = _inflow(cat, 1)[:, :, 0] * 1e3
    paths = formats.write_router_inputs(cat, x, str(tmp_path))
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "decipher_b200", "csrc", "build", "decipher_route_run")
    v = [400.0, 1500.0, 5000.0]
    r = subprocess.run([exe, "-river_data", paths["river_data"], "-tree", paths["flow_conn"], "-timeseries_in", paths["timeseries"],
                        "-velocities", ",".join(str(a) for a in v)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    want, _ = oracle_route_timeseries(cat, np.array(v), np.asfortranarray(np.repeat(x[:, :, None], 3, axis=2)))
    for s in range(3):
        lines = open(paths["timeseries"][:-4] + f"_routed_{s + 1}.txt").read().splitlines()
        got = np.array([[float(a) for a in ln.split(" ")] for ln in lines[1:]]).T
        assert got.shape == want[:, :, s].shape and np.abs(got - want[:, :, s]).max() <= 0.5e-6 + 1e-12
    bad = subprocess.run([exe, "-river_data", paths["river_data"]], capture_output=True, text=True, timeout=60)
    assert bad.returncode == 2


@pytest.mark.parametrize("tree,sea", [(False, False), (True, False), (True, True)])
def test_oracle_router_matches_closed_form(tree, sea):
    cat = _network(tree=tree)
    if sea:
        cat = _with_sea_outlet(cat)
        assert cat.nnode == cat.nriv + 1 and cat.node_type[0] == 1
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
@pytest.mark.parametrize("tree,nriv,nstep,sea", [(False, 3, 700, False), (True, 5, 1500, False), (True, 4, 900, True)])
def test_gpu_router_bit_exact(tree, nriv, nstep, sea):
    from decipher_b200.capi import DeviceCatchment
    cat = _network(tree=tree, nriv=nriv, nstep=nstep, seed=11)
    if sea:                                                        # ungauged sea outlet as the first node (route_run_standalone.f90:139-151)
        cat = _with_sea_outlet(cat)
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
