#!/usr/bin/env python
"""Benchmark of the thermal hot path: TEqn cell-updates/s (assemble + PCG/DIC + solidification correctors).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, one z-slab per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (CPU oracle)

A "step" is one full pass of additiveFoam's time-loop body (additiveFoam.C:66-95) over the whole block:
properties, time-step control, moving heat source, TEqn assembly, every solidification corrector and every
Krylov iteration.  Workload: BASELINE.json configs[2] -- the synthetic 100 M-cell uniform-hex block (1000 x 400 x 250 cells of 20 um),
single superGaussian track, conduction + solidification, implicit PCG/DIC with the shipped solver controls (tolerance
1e-15, maxIter 20).  It fits one B200 (about 32 GB), so N = 1 runs exactly that case; N GPUs run N such z-slabs
(weak scaling, 1000 x 400 x 250N cells).  `--nz-per-gpu 32` gives the thin-slab variant whose N = 8 run is the
102.4 M-cell case of the BASELINE target (12.8 M cells per GPU).
Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "TEqn cell-updates/s (assemble+PCG)"
UNIT = "cell-updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=1000)
    ap.add_argument("--ny", type=int, default=400)
    ap.add_argument("--nz-per-gpu", type=int, default=250)
    ap.add_argument("--workload", default="track", choices=["track", "rosenthal"])
    ap.add_argument("--explicit", action="store_true", help="as-shipped mode of configs[0]: explicitSolve true, maxDi 1 (diagonal solves only)")
    ap.add_argument("--solver", default="PCG", choices=["PCG", "PBiCGStab"])
    ap.add_argument("--preconditioner", default="DIC", choices=["DIC", "DILU", "diagonal", "none"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-nz", type=int, default=0, help="z layers of the CPU sample (0 = auto)")
    ap.add_argument("--cpu-steps", type=int, default=16)
    return ap.parse_args()


def make_case(args, nz, pre_steps_dt=2e-5):
    from additivefoam_b200 import abi, cases
    solver = dict(PCG=abi.SOLVER_PCG, PBiCGStab=abi.SOLVER_PBICGSTAB)[args.solver]
    precond = dict(DIC=abi.PRECOND_DIC, DILU=abi.PRECOND_DILU, diagonal=abi.PRECOND_DIAGONAL, none=abi.PRECOND_NONE)[args.preconditioner]
    n = (args.nx, args.ny, nz)
    ctl = dict(solver=solver, preconditioner=precond, deltaT=pre_steps_dt, explicitSolve=0, maxDi=10.0)
    if args.explicit:
        ctl.update(explicitSolve=1, maxDi=1.0, deltaT=2e-6)
    if args.workload == "rosenthal":
        return cases.rosenthal(n, **ctl)
    return cases.synthetic_track(n, **ctl)


def workload_name(args, nz):
    kind = "conduction-only Rosenthal track" if args.workload == "rosenthal" else "conduction+solidification track"
    if args.explicit:
        kind += ", explicitSolve true / maxDi 1 (as-shipped mode)"
    return (f"synthetic uniform-hex {args.nx}x{args.ny}x{nz} cells (20 um), single superGaussian beam, {kind}, implicit "
            f"{args.solver}/{args.preconditioner} tol 1e-15 maxIter 20 (BASELINE configs[2]: 100 M cells per GPU slab of "
            f"{args.nz_per_gpu} z-layers; weak scaling over z-slabs)")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        import statistics
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu(args, nz, steps, warmup, ranks):
    """The reference algorithm (CPU oracle, restatement -- not the OpenFOAM binary) on `ranks` host threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    os.environ["OMP_NUM_THREADS"] = str(ranks)
    case = make_case(args, nz)
    o = ol.OracleCase(case, ranks=ranks)
    for _ in range(warmup):
        o.step()
    t0 = time.perf_counter()
    iters = 0
    for _ in range(steps):
        r = o.step()
        iters += r.nSolveIterations
    dt = time.perf_counter() - t0
    n = o.n
    o.close()
    return dict(value=n * steps / dt, seconds=dt, cells=n, iterations=iters, steps=steps)


def ncu_traffic(group, n_local):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the group's kernel from the committed ncu --set full capture of
    this command (profiles/r1_top_kernels_ncu_full_<cells>.csv); None for sizes or groups that were not captured."""
    tag = {"sweep_fwd": "k_wave2<1,", "sweep_bwd": "k_wave2<2,", "factor": "k_wave2<0,", "amul": "k_amul_block<2>"}.get(group)
    path = os.path.join(ROOT, "profiles", f"r1_top_kernels_ncu_full_{n_local}.csv")
    if tag is None or not os.path.exists(path):
        return None
    import csv
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    vals = [float(r[hdr.index("dram_read_bytes")]) + float(r[hdr.index("dram_write_bytes")]) for r in rows[1:] if r and tag in r[0]]
    vals = [v for v in vals if v > 1e6]       # the exact factor sweep behind the fast one returns at once: not a real launch
    return sum(vals) / len(vals) if vals else None


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    N = max(args.gpus, 1)

    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = host_cores()
        nz = args.cpu_sample_nz or max(8, min(cores, args.nz_per_gpu))
        ranks = min(cores, nz)
        r = run_cpu(args, nz, args.steps, args.warmup, ranks)
        sample = f"{args.nx}x{args.ny}x{nz} cells of the per-GPU slab, {args.steps} steps after {args.warmup} warm-up, {ranks} z-slab ranks"
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": N, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args, args.nz_per_gpu * N), "note": "restatement of the reference algorithm "
                           "(CPU oracle, OpenMP over z-slab ranks), not the OpenFOAM binary: OpenFOAM-10 is not installable here"},
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": ranks, "kind": "port", "sample": sample},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    os.environ.setdefault("NCCL_DEBUG", "WARN")       # keep stdout to the one JSON line
    import numpy as np  # noqa: F401
    import torch
    from additivefoam_b200 import abi, lib

    uid = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t = torch.tensor(list(lib.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        uid = bytes(t.cpu().tolist())
    ctx = lib.Context(local, uid, rank, world)
    if world > 1 and os.environ.get("B200_PEER_ALLREDUCE", "1") != "0":
        from additivefoam_b200 import partition
        partition.enable_peer_allreduce(ctx, dist, device="cuda")
    nz = args.nz_per_gpu * world
    case = make_case(args, nz)
    g = lib.Case(ctx, case)
    n_global = g.n_global

    def barrier():
        ctx.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        g.step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.profile_enable(True)
    launches0 = ctx.kernel_launches()
    ctx.mark(0)
    iters = correctors = 0
    last = None
    for _ in range(args.steps):
        last = g.step()
        iters += last.nSolveIterations
        correctors += last.nThermoCorr
    ctx.mark(1)
    ms = ctx.elapsed_ms(0, 1)
    barrier()
    launches = ctx.kernel_launches() - launches0
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    sampler.stop_flag.set()
    sampler.join(2)
    if world > 1:
        tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    value = n_global * args.steps / (ms * 1e-3)

    # ---- end to end through the C ABI with HOST fields: every step T/alpha.solid/alpha.powder go host -> device and
    # T/alpha.solid come back (a host application that owns the fields, e.g. OpenFOAM function objects reading T)
    e2e = None
    if not args.no_e2e:
        nl = g.n_local
        hT = torch.from_numpy(g.field(abi.FIELD_T)).pin_memory().numpy()
        hA = torch.from_numpy(g.field(abi.FIELD_ALPHA_SOLID)).pin_memory().numpy()
        hP = torch.from_numpy(g.field(abi.FIELD_ALPHA_POWDER)).pin_memory().numpy()
        e_steps = max(2, args.steps // 2)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            g.set_field(abi.FIELD_ALPHA_POWDER, hP)
            g.set_field_keep_old(abi.FIELD_T, hT)
            g.set_field_keep_old(abi.FIELD_ALPHA_SOLID, hA)
            g.step()
            g.field_into(abi.FIELD_T, hT)
            g.field_into(abi.FIELD_ALPHA_SOLID, hA)
        barrier()
        et = time.perf_counter() - t0
        if world > 1:
            tm = torch.tensor([et], dtype=torch.float64, device="cuda")
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            et = float(tm.item())
        e2e = {"value": n_global * e_steps / et, "unit": UNIT, "h2d_bytes_per_step": 3 * 8 * nl * world,
               "d2h_bytes_per_step": 2 * 8 * nl * world, "steps": e_steps,
               "what": "b200_case_set_field(T, alpha.solid, alpha.powder) from pinned host + b200_case_step + b200_case_get_field(T, alpha.solid)"}

    if rank != 0:
        g.close(); ctx.close()
        dist.barrier()
        dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel group (CUDA events around every launch, live in this run)
    nl = g.n_local
    d = case["mesh"]["n"]
    nzl = nz // world
    nf = (d[0] - 1) * d[1] * nzl + d[0] * (d[1] - 1) * nzl + d[0] * d[1] * (nzl - 1)
    alg_bytes = {  # SURVEY.md 8(d), per launch
        "amul": 24 * nl + 16 * nf, "sweep_fwd": 24 * nl + 16 * nf, "sweep_bwd": 24 * nl + 16 * nf, "factor": 16 * nl + 16 * nf,
        "vector": None, "normfactor": 24 * nl + 8 * nl + 16 * nf + 24 * nl, "assemble": 48 * nl + 16 * nf, "corrector": 72 * nl,
        "alpha": 48 * nl, "props": 44 * nl, "dinum": 16 * nl, "faces": 8 * nl + 8 * nf}
    total_prof = sum(v[0] for v in prof.values()) or 1.0
    dom = max((k for k in prof if alg_bytes.get(k)), key=lambda k: prof[k][0])
    peak, peak_src = peak_hbm()
    shares = {k: round(v[0] / total_prof, 4) for k, v in prof.items() if v[1]}
    per_launch_ms = {k: v[0] / v[1] for k, v in prof.items() if v[1]}
    achieved = alg_bytes[dom] / (per_launch_ms[dom] * 1e-3) / 1e9
    traffic = ncu_traffic(dom, nl)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes[dom],
                "avg_launch_ms": per_launch_ms[dom], "launches": prof[dom][1], "share_of_kernel_time": shares.get(dom),
                "group_shares": shares,
                "group_GBps": {k: round(alg_bytes[k] / (per_launch_ms[k] * 1e-3) / 1e9, 1) for k in per_launch_ms if alg_bytes.get(k)},
                "note": "achieved = SURVEY.md 8(d) algorithmic bytes (general LDU addressing: 72 B/cell for Amul and for each sweep) / "
                        "measured launch time; the block path moves less (Amul 48, forward sweep 56, backward sweep 64 B/cell: no "
                        "addressing arrays, see DESIGN.md section 5), so compare `traffic` / avg_launch_ms with the peak for the DRAM rate"}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cores = host_cores()
        cnz = args.cpu_sample_nz or max(8, min(cores, args.nz_per_gpu))
        ranks = min(cores, cnz)
        try:
            r = run_cpu(args, cnz, args.cpu_steps, args.warmup, ranks)
            cpu = {"value": r["value"], "unit": UNIT, "cores": ranks, "kind": "port",
                   "sample": f"{args.nx}x{args.ny}x{cnz} cells ({cnz}/{args.nz_per_gpu} of the slab), {args.cpu_steps} steps after {args.warmup} warm-up, "
                             f"{r['seconds']:.1f} s; restatement of the reference algorithm, not the OpenFOAM binary"}
        except Exception as e:     # the oracle is only the reported baseline
            cpu = {"value": None, "unit": UNIT, "cores": ranks, "kind": "port", "sample": f"failed: {e}"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args, nz), "cells": n_global, "krylov_iterations": iters, "correctors": correctors,
                       "cell_iterations_per_s": n_global * iters / (ms * 1e-3), "l2": "inputs larger than L2 (>= 100 MB per field)",
                       "last_step": last.as_dict()},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    g.close(); ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
