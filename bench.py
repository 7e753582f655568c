#!/usr/bin/env python
"""bench.py — member-HRU-timesteps/s of the DECIPHeR Monte-Carlo rainfall-runoff path on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`, one JSON line
from rank 0.  For N>1 it is launched under torchrun (one rank per GPU, NCCL).

Workload = BASELINE.json configs[1] ("C2"): synthetic 500-HRU catchment (kW=8, 3 gauged reaches,
4 rain grids), hourly forcing, parameter sets sampled by RANMAR from the shipped params.dat bounds.
One "step" = one pass of the hot path (init_satzone + build_hist + the Topmod time loop + routing +
NSE/log-NSE) over one batch of `--members` ensemble members for `--nstep` timesteps.  Members are
independent and identically sized, so throughput per unit does not depend on how the 10^5-member,
87 600-step job is cut into such batches; `--nstep 87600` runs full-length simulations.

  value  : device-resident inputs (parameters already in HBM, outputs left in HBM)
  e2e    : the same call through host buffers (pinned), H2D of the parameter sets and D2H of the
           performance measures / status words inside the timed region
  roofline: FP64-pipe roofline of the physics kernel (the path is FP64-pipe bound, not HBM
           or tensor bound): weighted algorithmic flops (App. D) and, beside them, the executed FP64 count of the
           committed ncu capture.
  parity  : the oracle leg that gives `cpu_baseline` runs parameter sets of the last TIMED pass; the GPU's
           measures for those members are compared with it (max relative deviation, tolerance 1e-9).
Every pass (warm-up and timed) simulates DISTINCT parameter sets of one RANMAR table.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "member-HRU-timesteps/sec (MC ensemble)"
UNIT = "member-HRU-timesteps/s"


def algorithmic_flops_per_unit(cat, hist_len):
    """131 + 3*kW + 5*kR + 2*T*nnode/nac weighted FP64 flops per member-HRU-step."""
    kW = float(cat.ntrans.sum()) / cat.nac
    kR = float((cat.rivers != 0).sum()) / cat.nac
    return 131.0 + 3.0 * kW + 5.0 * kR + 2.0 * hist_len * cat.nnode / cat.nac


def build_workload(args):
    from decipher_b200 import synth
    cfg = dict(synth.CONFIGS["C2"])
    cfg.pop("members")
    if getattr(args, "nac", 0) > 0:
        cfg["nac"] = args.nac
    raw = synth.make_raw(**cfg)
    if args.nstep < raw["nstep"]:                     # time window of the same 10-year forcing
        n = args.nstep
        raw["rain"] = np.asfortranarray(raw["rain"][:, :n])
        raw["pet"] = np.asfortranarray(raw["pet"][:, :n])
        raw["qobs"] = np.asfortranarray(raw["qobs"][:, :n])
        raw["nstep"] = n
    return synth.assemble(raw)


def workload_config(args, cat, world=1, extra=None):
    name = "C2" if world == 1 else "C3"
    cfg = {"workload": f"{name}: synthetic {cat.nac}-HRU catchment, 3 gauged reaches, kW=8, hourly forcing (10-year series), "
                       "MC ensemble sampled with RANMAR seeds 5/2000 from RRMODEL_EXAMPLE_FILES/params.dat bounds"
                       + ("" if world == 1 else f", members sharded over {world} GPUs"),
           "nac": cat.nac, "nriv": cat.nriv, "nstep_per_step": cat.nstep, "nstep_full_series": 87600,
           "members_per_step_per_gpu": args.members, "ensemble_of_config": 100000 if world == 1 else 1000000,
           "qobs": cat.meta.get("qobs_source", "linear-reservoir proxy (synth.make_raw)"),
           "units_per_step_per_gpu": int(args.members) * cat.nac * cat.nstep}
    if extra:
        cfg.update(extra)
    return cfg


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); pw.append(float(parts[3]))
            except ValueError:
                continue
            for n, v in zip(names, parts[4:8]):
                if v.lower() == "active":
                    reasons.add(n)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(pw)))
        return out


def cpu_baseline(cat, cores, flags=0, mcpar=None):
    """Times the oracle port on the host cores over a bounded member sample of the same workload (`mcpar`: the parameter
    sets to run, default the first ones of the RANMAR stream).  Returns (info, seconds, units, oracle outputs)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import oracle_run
    from decipher_b200 import synth
    if mcpar is None:
        per_thread = max(1, int(round(4.0e7 / (cat.nac * cat.nstep))))
        mcpar = synth.sample_parameters(cat.npt, cores * per_thread)
    n = mcpar.shape[2]
    t0 = time.perf_counter()
    ref = oracle_run(cat, mcpar, flags=flags, want_q=False, want_perf=True, want_totals=False, n_threads=cores)
    dt = time.perf_counter() - t0
    units = n * cat.nac * cat.nstep
    return {"value": units / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} members x {cat.nac} HRUs x {cat.nstep} steps of the same catchment, "
                      f"C++ oracle -O2 strict FP, {cores} threads over disjoint members, {dt:.1f} s"}, dt, units, ref


def ncu_figures(kname, shape):
    """Figures that only a profiler can give, taken from the committed capture of the SAME launch shape (never measured
    in this run: a number taken under ncu is not a bench value).  profiles/ncu_figures.json maps kernel -> figures."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_figures.json")))
    except Exception:
        return {}
    ent = table.get(kname, {})
    if ent and tuple(ent.get("launch_shape", ())) not in ((), tuple(shape)):
        return {"traffic_source": f"no ncu capture of launch shape {shape} committed (profiles/ncu_figures.json holds {ent.get('launch_shape')})",
                "flops_executed": ent.get("flops_executed"), "executed_source": ent.get("executed_source")}
    return ent


def run_reference(args):
    """--impl reference: the reference's CPU path.  The Fortran sources cannot be compiled in this image
    (no gfortran/flang/ifort), so this times the line-faithful C++ port on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cat = build_workload(args)
    cores = os.cpu_count() or 1
    vals, secs = [], []
    info = None
    for i in range(args.warmup + args.steps):
        info, dt, units, _ = cpu_baseline(cat, cores)
        if i >= args.warmup:
            vals.append(units); secs.append(dt)
    value = sum(vals) / sum(secs)
    info["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # same keys as the GPU arm's config (the driver compares the two lines); the CPU arm times a bounded member sample
            "config": workload_config(args, cat, 1, {
                "kernel": "cpu", "math": "exact", "distinct_members_simulated": int(info["sample"].split()[0]) * (args.warmup + args.steps),
                "l2": "n/a (host cores)", "parallelism": f"{cores} host threads over disjoint members (C++ port of the reference path); "
                                                         "bounded member sample per step, see cpu_baseline.sample"}),
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def perf_parity(p_gpu, p_ref):
    """Largest relative deviation of the GPU measures from the oracle's on the same members: NSE / log-NSE / KGE on
    1 - measure, PBIAS against the 100 % scale, RMSE relatively (the comparator of tests/parity.py); NaN patterns must agree."""
    if not np.array_equal(np.isnan(p_gpu), np.isnan(p_ref)):
        return float("inf")
    worst = 0.0
    for k in range(p_ref.shape[0]):
        ok = np.isfinite(p_ref[k])
        if not ok.any():
            continue
        a, b = p_gpu[k][ok], p_ref[k][ok]
        den = np.maximum(np.abs(b), 100.0) if k == 3 else (np.maximum(np.abs(b), 1e-300) if k == 4 else np.maximum(np.abs(1.0 - b), 1e-300))
        worst = max(worst, float((np.abs(a - b) / den).max()))
    return worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C2", "C4", "C5"],
                    help="C2 (default; C3 under torchrun): the configuration the metric is quoted on. C4 / C5: the dense-W and the "
                         "multi-catchment configurations of BASELINE.json (extra lines, see bench_configs.py)")
    ap.add_argument("--members", type=int, default=0, help="ensemble members per step per GPU (0 = kernel default)")
    ap.add_argument("--nstep", type=int, default=0, help="timesteps simulated per step (0 = kernel default, 87600 = full)")
    ap.add_argument("--nac", type=int, default=0, help="HRUs of the synthetic catchment (0 = the config's 500)")
    ap.add_argument("--kw", default="", help="C4: targets per HRU, 64 | 512 | dense")
    ap.add_argument("--catchments", type=int, default=0, help="C5: number of catchments (0 = default)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "streamed", "resident", "stepwise"])
    ap.add_argument("--math", default="fast", choices=["exact", "fast"],
                    help="fast = reciprocal-hoisted branch-free arithmetic (parity <= 1e-9, tests/test_gpu_parity.py); "
                         "exact = the reference's operation order with IEEE division and library exp/log")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-exact", action="store_true", help="skip the extra exact-math step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3 if args.impl == "ours" else 0)
    if args.config != "C2":
        import bench_configs
        return bench_configs.main(args)

    # defaults per kernel: the resident kernel runs FULL-LENGTH simulations (87 600 hourly steps) for 8 "waves"
    # of members (one wave = 148 SMs x 4 members); the streamed kernel needs ~10^5 members in flight, so its
    # step integrates a 3-month window of the same series instead.
    if args.members <= 0:
        args.members = 100000 if args.kernel == "streamed" else 4736
    if args.nstep <= 0:
        args.nstep = 2190 if args.kernel == "streamed" else 87600
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from decipher_b200 import synth
    from decipher_b200 import capi
    from decipher_b200.distributed import PerfGather, shard_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the DECIPHeR B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cat = build_workload(args)
    dev = capi.DeviceCatchment(cat, device=local_rank)
    M = args.members
    n_pass = args.warmup + args.steps
    # Whole-job parameter table from ONE sequential RANMAR stream: pass k of the bench (warm-up passes first) runs the
    # job's members [k*M*world, (k+1)*M*world), this rank its contiguous block of them -- every pass simulates DISTINCT
    # parameter sets; with the driver's 20 + 3 passes of 4 736 members that is C2's 10^5-member ensemble once.
    table = synth.sample_parameters(cat.npt, M * world * n_pass)
    host_mc, d_mc = [], []
    for k in range(n_pass):
        lo, hi = shard_range(M * world, rank, world)
        mine = np.ascontiguousarray(table[:, :, k * M * world + lo:k * M * world + hi].transpose(2, 1, 0))   # (M, 7, npt) == ABI layout
        h = torch.empty(mine.shape, dtype=torch.float64).pin_memory()
        h.copy_(torch.from_numpy(mine))
        host_mc.append(h)
        d_mc.append(h.cuda())
    del table
    d_work = torch.empty_like(d_mc[0])
    d_perf = torch.empty((M, cat.nriv, capi.NMEASURE), dtype=torch.float64, device="cuda")
    d_status = torch.empty(M, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")               # > 126 MB L2
    kernel = {"auto": capi.KERNEL_AUTO, "streamed": capi.KERNEL_STREAMED, "resident": capi.KERNEL_RESIDENT,
              "stepwise": capi.KERNEL_STEPWISE}[args.kernel]
    flags = capi.OPT_FAST_MATH if args.math == "fast" else 0
    # the only collective: ONE gather of the measures to rank 0 per run, preallocated, timed on its own
    gather = PerfGather(M, (cat.nriv, capi.NMEASURE), world, rank, torch.device("cuda", local_rank)) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(k):
        flush.zero_()                                  # L2 flush between timed iterations
        d_work.copy_(d_mc[k])                          # srinit clip is in/out: the pass's own parameter sets
        torch.cuda.synchronize()
        return dev.run_device(d_work.data_ptr(), M, perf_ptr=d_perf.data_ptr(), status_ptr=d_status.data_ptr(),
                              flags=flags, kernel=kernel)

    for k in range(args.warmup):
        step_device(k)
        if gather is not None:
            gather(d_perf)                             # NCCL communicator set-up happens here, outside the timed region
    barrier()
    clocks = ClockSampler(local_rank)
    wall0 = time.perf_counter()
    ms_dev, ms_phys, n_phys, launches, ms_gather = 0.0, 0.0, 0, 0, 0.0
    st = None
    for k in range(args.warmup, n_pass):
        st = step_device(k)
        ms_dev += st["ms_total"]
        ms_phys += st["ms_physics"]; n_phys += st["n_physics_launches"]; launches += st["n_launches"]
    if gather is not None:                             # one gather per run, as a real C3 job does (members of the last pass)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gather(d_perf); e1.record(); e1.synchronize()
        ms_gather = e0.elapsed_time(e1)
        ms_dev += ms_gather
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clk = clocks.stop()
    perf_last = d_perf.cpu().numpy().transpose(2, 1, 0).copy()       # (NMEASURE, nriv, M) of the last timed pass
    status_last = d_status.cpu().numpy().copy()

    # e2e through the public host-buffer API (pinned parameter sets in, measures/status/diagnostics out), HOST WALL
    # CLOCK around the call: H2D, every kernel, D2H and the library's own host work are all inside
    ms_e2e_wall, ms_e2e_ev = 0.0, 0.0
    work = torch.empty_like(host_mc[0]).pin_memory()
    w_np = work.numpy().transpose(2, 1, 0)
    barrier()
    out = None
    for k in range(args.warmup, n_pass):
        flush.zero_()
        work.copy_(host_mc[k])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = dev.run(w_np, flags=flags, kernel=kernel, want_q=False, want_perf=True, want_totals=False)
        ms_e2e_wall += (time.perf_counter() - t0) * 1e3
        ms_e2e_ev += out["stats"]["ms_total"]
    if gather is not None:
        t_perf = torch.from_numpy(np.ascontiguousarray(out["perf"].transpose(2, 1, 0))).cuda()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gather(t_perf); torch.cuda.synchronize()
        g = (time.perf_counter() - t0) * 1e3
        ms_e2e_wall += g; ms_e2e_ev += g
    barrier()
    h2d = int(w_np.nbytes)
    d2h = int(w_np.nbytes + out["perf"].nbytes + out["status"].nbytes + out["init_diag"].nbytes)
    e2e_same_bits = bool(np.array_equal(out["perf"], perf_last, equal_nan=True))  # host-buffer and device-buffer calls agree

    # max over ranks (device times)
    t = torch.tensor([ms_dev, ms_e2e_wall, wall_ms, ms_e2e_ev], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev_max, ms_e2e_max, wall_max, ms_e2e_ev_max = [float(x) for x in t.tolist()]
    units_step_gpu = M * cat.nac * cat.nstep
    units_total = units_step_gpu * world * args.steps
    value = units_total / (ms_dev_max * 1e-3)
    e2e_value = units_total / (ms_e2e_max * 1e-3)

    if rank == 0:
        fp64_peak, _ = capi.measure_fp64_peak()        # dependency-free DFMA chains, TFLOP/s (measured live)
        fpu = algorithmic_flops_per_unit(cat, st["hist_len"])
        launch_ms = ms_phys / max(n_phys, 1)
        units_per_launch = units_step_gpu * args.steps / max(n_phys, 1)
        ach_tf = fpu * units_per_launch / (launch_ms * 1e-3) / 1e12
        kname = {1: "k_physics_streamed", 2: "k_physics_resident_v2" if args.math == "fast" else "k_physics_resident",
                 3: "k_redistribute_* + k_physics_stepwise"}.get(st["kernel_used"], "?")
        ncu = ncu_figures(kname, (M, cat.nstep, cat.nac))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, cat, world, {
                "kernel": {1: "streamed", 2: "resident", 3: "stepwise"}.get(st["kernel_used"], "?"), "math": args.math,
                "distinct_members_simulated": M * world * n_pass,
                "l2": "256 MiB buffer written between timed steps (L2 flush)", "parallelism": f"members sharded over {world} GPU(s), no data-path collective; ONE NCCL gather of the measures per run"}),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "clock": "host wall clock around dcp_run_ensemble (host buffers)", "value_by_cuda_events": units_total / (ms_e2e_ev_max * 1e-3),
                    "bit_identical_to_device_buffer_call": e2e_same_bits},
            "gpu_launches": int(launches),
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"],
                       "power_w_max": clk.get("power_w_max"), "samples": clk["samples"]},
            "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": ach_tf / fp64_peak, "traffic": None, "traffic_source": ncu.get("traffic_source"),
                         "traffic_ncu_bytes_per_launch": ncu.get("traffic"),
                         "kernel": kname,
                         "flops_per_unit": fpu, "units_per_launch": units_per_launch, "launch_ms": launch_ms,
                         "flops_per_unit_executed": ncu.get("flops_executed"), "executed_source": ncu.get("executed_source"),
                         "achieved_executed": (ncu["flops_executed"] * units_per_launch / (launch_ms * 1e-3) / 1e12) if ncu.get("flops_executed") else None,
                         "peak_source": "dcp_measure_fp64_peak (DFMA chains, measured in this run; MEASURED_PEAKS.json has no FP64 figure)",
                         "physics_share_of_step": ms_phys / max(ms_dev_max - ms_gather, 1e-9)},
            "wall_ms_per_step": wall_max / args.steps, "gather_ms_per_run": ms_gather,
        }
        if world == 1 and not args.no_cpu_baseline:
            # CPU baseline AND parity self-check: the oracle runs the first `cores` parameter sets of the LAST TIMED pass
            cores = os.cpu_count() or 1
            n_chk = min(M, cores * max(1, int(round(4.0e7 / (cat.nac * cat.nstep)))))
            mc_chk = np.asfortranarray(host_mc[n_pass - 1].numpy()[:n_chk].transpose(2, 1, 0))
            info, _, _, ref = cpu_baseline(cat, cores, mcpar=mc_chk)
            line["cpu_baseline"] = info
            line["parity"] = {"max_rel": perf_parity(perf_last[:, :, :n_chk], ref["perf"]), "n": int(n_chk), "tol": 1e-9,
                              "what": "NSE, log-NSE, KGE, PBIAS, RMSE of the first n members of the last timed pass: GPU "
                                      "(timed run) vs the C++ oracle of the cpu_baseline leg, comparator of tests/parity.py",
                              "status_equal": bool(np.array_equal(status_last[:n_chk], ref["status"])),
                              "oracle": "in-repo C++ restatement of the reference (Fortran not buildable here: parity unpinned upstream)"}
        if world == 1 and args.math == "fast" and not args.no_exact:
            # the same step with the EXACT arithmetic policy (bit-faithful to the reference up to exp/log), for reference
            d_work.copy_(d_mc[n_pass - 1]); torch.cuda.synchronize()
            st_e = dev.run_device(d_work.data_ptr(), M, perf_ptr=d_perf.data_ptr(), status_ptr=d_status.data_ptr(), flags=0, kernel=kernel)
            line["exact_math"] = {"value": units_step_gpu / (st_e["ms_total"] * 1e-3), "unit": UNIT, "ms_per_step": st_e["ms_total"], "steps": 1}
        print(json.dumps(line), flush=True)
    dev.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
