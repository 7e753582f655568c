"""CPU checks of the measurement scripts' host logic (no GPU): the measure comparator of the bench's parity self-check,
the C5 partitioning inputs, the profiler-figure lookup and the reference arm's JSON contract."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bench_configs  # noqa: E402
from decipher_b200.distributed import lpt_assign  # noqa: E402


def test_perf_parity_comparator():
    rng = np.random.default_rng(0)
    p = rng.uniform(-1, 1, (5, 3, 8))
    p[3] *= 50.0
    p[4] = np.abs(p[4]) * 1e-4
    assert bench.perf_parity(p.copy(), p) == 0.0
    q = p.copy(); q[0, 1, 2] = 1.0 - (1.0 - p[0, 1, 2]) * (1 + 1e-10)          # NSE is compared on 1 - NSE
    assert abs(bench.perf_parity(q, p) - 1e-10) < 1e-13
    q = p.copy(); q[2, 0, 0] = np.nan
    assert bench.perf_parity(q, p) == float("inf")                               # NaN patterns must agree


def test_c5_shapes_are_rank_independent_and_balanced():
    shapes = bench_configs.c5_shapes(64)
    assert shapes == bench_configs.c5_shapes(64) and len(shapes) == 64
    nac = np.array([s[0] for s in shapes])
    assert nac.min() >= 100 and nac.max() <= 1000 and (nac > 512).sum() > 10
    owner = lpt_assign([n * 87600 * 10000 for n in nac], 8)
    load = np.bincount(owner, weights=nac.astype(float), minlength=8)
    assert load.max() / load.mean() < 1.02


def test_ncu_figures_lookup():
    f = bench.ncu_figures("k_physics_resident_v2", (4736, 87600, 500))
    assert f["traffic"] > 0 and "profiles/" in f["traffic_source"] and f["flops_executed"] > 100
    g = bench.ncu_figures("k_physics_resident_v2", (1, 2, 3))
    assert "traffic" not in g and "no ncu capture" in g["traffic_source"]
    assert bench.ncu_figures("unknown_kernel", (1, 2, 3)) == {}


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--nstep", "300"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["value"] > 0 and d["unit"] == bench.UNIT and d["metric"] == bench.METRIC and d["higher_is_better"] is True
    assert {"kernel", "math", "l2", "parallelism", "workload", "nac"} <= set(d["config"])
