#!/usr/bin/env python
"""bench.py --config C4 | C5: the two BASELINE.json configurations beside the one the metric is quoted on.

  C4  dense-connectivity catchment: 5 000 HRUs, dynamic_dist weighting matrix with kW in {64, 512, dense} targets per HRU
      (RRMODEL/dyna_dynamic_dist.f90:47-67 is the hot spot), nstep = 8 760, one GPU.  The catchment does not fit on chip,
      so the stepwise path runs it (one launch set per timestep; CSR SpMM or the FP64 GEMM when W is dense enough).
  C5  continental multi-catchment run: N synthetic gauged catchments (nac ~ U{100..1000}, own forcing, own RANMAR
      stream) x M members, catchments dealt to the ranks longest-processing-time first, every rank runs its
      catchments one after the other, ONE gather of the measures at the end.  Reports per-rank busy time and imbalance.

Same JSON contract as bench.py (metric, value, e2e, roofline, cpu_baseline at N = 1); `config.workload` names the case.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "member-HRU-timesteps/sec (MC ensemble)"
UNIT = "member-HRU-timesteps/s"


def _flops_per_unit(cat, hist_len):
    kW = float(cat.ntrans.sum()) / cat.nac
    kR = float((cat.rivers != 0).sum()) / cat.nac
    return 131.0 + 3.0 * kW + 5.0 * kR + 2.0 * hist_len * cat.nnode / cat.nac


def _cpu_sample(cat, mc, cores):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import oracle_run
    t0 = time.perf_counter()
    ref = oracle_run(cat, mc, want_q=False, want_perf=True, want_totals=False, n_threads=cores)
    return ref, time.perf_counter() - t0


def run_c4(args):
    import torch
    from decipher_b200 import capi, synth
    import bench as B
    if int(os.environ.get("WORLD_SIZE", "1")) != 1:
        raise SystemExit("--config C4 is a one-GPU configuration")
    kw = args.kw or "64"
    nac = args.nac or 5000
    nstep = args.nstep or 8760
    M = args.members or 1024
    cfg = dict(synth.CONFIGS["C4"]); cfg.pop("members")
    cfg.update(nac=nac, nstep=nstep)
    if kw == "dense":
        cfg.update(dense_w=True)
    else:
        cfg.update(kW=int(kw))
    cat = synth.synth_catchment(**cfg)
    torch.cuda.set_device(0)
    dev = capi.DeviceCatchment(cat, device=0)
    n_pass = args.warmup + args.steps
    table = synth.sample_parameters(cat.npt, M * n_pass)
    host_mc = [np.ascontiguousarray(table[:, :, k * M:(k + 1) * M].transpose(2, 1, 0)) for k in range(n_pass)]
    d_mc = [torch.from_numpy(h).cuda() for h in host_mc]
    d_work = torch.empty_like(d_mc[0])
    d_perf = torch.empty((M, cat.nriv, capi.NMEASURE), dtype=torch.float64, device="cuda")
    d_status = torch.empty(M, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flags = capi.OPT_FAST_MATH if args.math == "fast" else 0
    kernel = {"auto": capi.KERNEL_AUTO, "streamed": capi.KERNEL_STREAMED, "resident": capi.KERNEL_RESIDENT,
              "stepwise": capi.KERNEL_STEPWISE}[args.kernel]

    def step(k):
        flush.zero_(); d_work.copy_(d_mc[k]); torch.cuda.synchronize()
        return dev.run_device(d_work.data_ptr(), M, perf_ptr=d_perf.data_ptr(), status_ptr=d_status.data_ptr(), flags=flags, kernel=kernel)

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    clocks = B.ClockSampler(0)
    ms_dev = ms_phys = 0.0
    n_phys = launches = 0
    st = None
    for k in range(args.warmup, n_pass):
        st = step(k)
        ms_dev += st["ms_total"]; ms_phys += st["ms_physics"]; n_phys += st["n_physics_launches"]; launches += st["n_launches"]
    torch.cuda.synchronize()
    clk = clocks.stop()
    perf_last = d_perf.cpu().numpy().transpose(2, 1, 0).copy()
    status_last = d_status.cpu().numpy().copy()
    ms_e2e = 0.0
    out = None
    for k in range(args.warmup, n_pass):
        flush.zero_(); torch.cuda.synchronize()
        w = np.asfortranarray(host_mc[k].transpose(2, 1, 0))
        t0 = time.perf_counter()
        out = dev.run(w, flags=flags, kernel=kernel, want_q=False, want_perf=True)
        ms_e2e += (time.perf_counter() - t0) * 1e3
    units_step = M * cat.nac * cat.nstep
    units = units_step * args.steps
    fpu = _flops_per_unit(cat, st["hist_len"])
    fp64_peak, _ = capi.measure_fp64_peak()
    nnz = int(cat.ntrans.sum())
    # dominant work of the dense case: the contraction QIN = W . QBF, 2 nnz(W) flops per member-step
    gemm_tf = 2.0 * nnz * M * cat.nstep * args.steps / (ms_phys * 1e-3) / 1e12
    ach_tf = fpu * units / (ms_phys * 1e-3) / 1e12
    line = {"metric": METRIC, "value": units / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"C4: dense-connectivity catchment, {cat.nac} HRUs, kW = {kw} ({nnz} non-zeros of W, "
                                   f"{100.0 * nnz / cat.nac ** 2:.1f} % full), {cat.nriv} gauged reaches (tree), {cat.nstep} hourly steps",
                       "nac": cat.nac, "kW": kw, "nnz_W": nnz, "nstep_per_step": cat.nstep, "members_per_step_per_gpu": M,
                       "ensemble_of_config": 10000, "kernel": {1: "streamed", 2: "resident", 3: "stepwise"}.get(st["kernel_used"], "?"),
                       "math": args.math, "distinct_members_simulated": M * n_pass, "l2": "256 MiB buffer written between timed steps (L2 flush)"},
            "e2e": {"value": units / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(host_mc[0].nbytes),
                    "d2h_bytes_per_step": int(host_mc[0].nbytes + out["perf"].nbytes + out["status"].nbytes + out["init_diag"].nbytes),
                    "clock": "host wall clock around dcp_run_ensemble (host buffers)"},
            "gpu_launches": int(launches),
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"], "power_w_max": clk.get("power_w_max"),
                       "samples": clk["samples"]},
            "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak, "traffic": None,
                         "flops_per_unit": fpu, "kernel": "k_redistribute_* + k_physics_step + k_river_sum (one launch set per timestep)",
                         "contraction_tflops": gemm_tf, "launch_sets": int(n_phys), "physics_share_of_step": ms_phys / ms_dev,
                         "peak_source": "dcp_measure_fp64_peak (DFMA chains, measured in this run)"}}
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_chk = min(M, cores)
        mc_chk = np.asfortranarray(host_mc[n_pass - 1][:n_chk].transpose(2, 1, 0))
        ref, dt = _cpu_sample(cat, mc_chk, cores)
        line["cpu_baseline"] = {"value": n_chk * cat.nac * cat.nstep / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n_chk} members x {cat.nac} HRUs x {cat.nstep} steps, C++ oracle -O2 strict FP, {dt:.1f} s"}
        line["parity"] = {"max_rel": B.perf_parity(perf_last[:, :, :n_chk], ref["perf"]), "n": int(n_chk), "tol": 1e-9,
                          "status_equal": bool(np.array_equal(status_last[:n_chk], ref["status"]))}
    print(json.dumps(line), flush=True)
    dev.close()


def c5_shapes(n_cat, seed0=20260119 + 5000):
    """C5 catchment shapes: nac ~ U{100..1000}, 2-4 gauged reaches; cheap, identical on every rank."""
    rng = np.random.default_rng(seed0)
    return [(int(rng.integers(100, 1001)), int(rng.integers(2, 5))) for _ in range(n_cat)]


def make_c5_catchment(i, nac, nriv, nstep, seed0=20260119 + 5000):
    """One C5 catchment with its own forcing (only the rank that owns it builds it)."""
    from decipher_b200 import synth
    return synth.synth_catchment(nac=nac, nriv=nriv, npt=1, ngp=4, nge=1, nstep=nstep, kW=8, seed=seed0 + 1 + i, tree=True)


def run_c5(args):
    import torch
    import torch.distributed as dist
    from decipher_b200 import capi, synth
    from decipher_b200.distributed import lpt_assign
    import bench as B
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_cat = args.catchments or (8 * world)
    M = args.members or 2000
    nstep = args.nstep or 87600
    shapes = c5_shapes(n_cat)
    costs = [nac * nstep * M for nac, _ in shapes]
    owner = lpt_assign(costs, world)
    mine = [i for i in range(n_cat) if owner[i] == rank]
    flags = capi.OPT_FAST_MATH if args.math == "fast" else 0
    cats = {i: make_c5_catchment(i, shapes[i][0], shapes[i][1], nstep) for i in mine}
    tables = {i: synth.sample_parameters(1, M, seeds=(5 + i, 2000)) for i in mine}
    handles = {i: capi.DeviceCatchment(cats[i], device=local_rank) for i in mine}        # static data of a catchment lives on its GPU only

    def one_pass(n_mem):
        busy = 0.0; kernels = {}; launches = 0; perf = {}
        t0 = time.perf_counter()
        for i in mine:
            out = handles[i].run(np.asfortranarray(tables[i][:, :, :n_mem]), flags=flags, want_q=False, want_perf=True)
            busy += out["stats"]["ms_total"]; launches += out["stats"]["n_launches"]
            kernels[out["stats"]["kernel_used"]] = kernels.get(out["stats"]["kernel_used"], 0) + 1
            perf[i] = out["perf"]
        return busy, (time.perf_counter() - t0) * 1e3, kernels, launches, perf

    one_pass(max(M // 10, 1))                           # a full pass is minutes of GPU time: the warm-up pass (clocks, allocations,
                                                        # first NCCL call below) runs a tenth of the members
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = B.ClockSampler(local_rank)
    busy = wall = 0.0; launches = 0; kernels = {}; perf = {}
    steps = max(1, min(args.steps, 2))
    for _ in range(steps):
        b, w, kernels, l, perf = one_pass(M)
        busy += b; wall += w; launches += l
    # the only collective: the measures of every catchment to rank 0 (ragged blocks: grouped send / recv)
    g_ms = 0.0
    if world > 1:
        nmax = max(sum(1 for o in owner if o == r) for r in range(world))
        blk = torch.zeros((nmax, capi.NMEASURE, 4, M), dtype=torch.float64, device="cuda")
        for j, i in enumerate(mine):
            blk[j, :, :shapes[i][1], :] = torch.from_numpy(np.ascontiguousarray(perf[i])).cuda()
        outs = [torch.empty_like(blk) for _ in range(world)] if rank == 0 else None
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dist.gather(blk, outs, dst=0); e1.record(); e1.synchronize()
        g_ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    t = torch.tensor([busy + g_ms, wall + g_ms], dtype=torch.float64, device="cuda")
    allb = [torch.zeros(2, dtype=torch.float64, device="cuda") for _ in range(world)]
    if world > 1:
        dist.all_gather(allb, t)
    else:
        allb = [t]
    per_rank = [float(x[0]) / steps for x in allb]
    per_rank_wall = [float(x[1]) / steps for x in allb]
    if rank == 0:
        units = float(sum(costs)) * steps
        ms_max = max(per_rank) * steps
        nac_all = [nac for nac, _ in shapes]
        line = {"metric": METRIC, "value": units / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": 1,
                "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"C5: {n_cat} synthetic gauged catchments (nac ~ U{{100..1000}}: min {min(nac_all)}, median {int(np.median(nac_all))}, "
                                       f"max {max(nac_all)}; {sum(1 for n in nac_all if n > 512)} above 512 HRUs), {M} members each, {nstep} hourly steps, "
                                       f"catchments dealt to {world} rank(s) longest-processing-time first",
                           "catchments": n_cat, "members_per_catchment": M, "nstep_per_step": nstep, "ensemble_of_config": "1000 catchments x 1e4 members",
                           "math": args.math, "kernels_used_rank0": {str(k): v for k, v in kernels.items()},
                           "parallelism": "catchments partitioned over ranks, no data-path collective; one gather of the measures"},
                "e2e": {"value": units / (max(per_rank_wall) * steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(sum(7 * 8 * M for _ in shapes)),
                        "d2h_bytes_per_step": int(sum((7 + capi.NMEASURE * nr + 3) * 8 * M + 4 * M for _, nr in shapes)),
                        "clock": "host wall clock around the rank's dcp_run_ensemble calls (host buffers), max over ranks"},
                "gpu_launches": int(launches),
                "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"], "power_w_max": clk.get("power_w_max"),
                           "samples": clk["samples"]},
                "balance": {"busy_ms_per_rank": per_rank, "imbalance_max_over_mean": max(per_rank) / (sum(per_rank) / len(per_rank)),
                            "cost_model_imbalance": float(max(np.bincount(owner, weights=costs, minlength=world)) / (sum(costs) / world)),
                            "gather_ms": g_ms},
                "roofline": {"bound": "fp64", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None,
                             "note": "mixed kernels over catchments of different size; see the C2 line for the kernel roofline"}}
        print(json.dumps(line), flush=True)
    for h in handles.values():
        h.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def main(args):
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"--impl reference is defined for the headline configuration (C2); "
                                                                  f"{args.config} carries its own cpu_baseline"}), flush=True)
        return
    if args.config == "C4":
        return run_c4(args)
    return run_c5(args)
