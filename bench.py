#!/usr/bin/env python
"""Benchmark of the thermal hot path: TEqn cell-updates/s (assemble + PCG/DIC + solidification correctors).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, one z-slab per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (CPU oracle)

A "step" is one full pass of additiveFoam's time-loop body (additiveFoam.C:66-95) over the whole block:
properties, time-step control, moving heat source, TEqn assembly, every solidification corrector and every
Krylov iteration.  Workload: BASELINE.json configs[2] -- the synthetic 100 M-cell uniform-hex block (1000 x 400 x 250 cells of 20 um),
single superGaussian track, conduction + solidification, implicit PCG/DIC with the shipped solver controls (tolerance
1e-15, maxIter 20), "decomposed 1/2/4/8": the SAME 100 M-cell mesh is cut into N z-slabs, one per GPU (strong scaling; the
reference decomposes one case over its MPI ranks the same way, tutorials/AMB2018-02-B/system/decomposeParDict:18-20).  The
weak-scaled number (N slabs of 100 M cells each) is carried as the sub-record `weak`.  `--impl reference` runs the same global
mesh on the host cores.  With N > 1 a small decomposed case is first checked against the oracle's emulation of the same N
ranks (`parity_check` on the JSON line).
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
    ap.add_argument("--nz", type=int, default=250, help="z layers of the global mesh (strong scaling: cut into --gpus slabs)")
    ap.add_argument("--nz-per-gpu", type=int, default=250, help="z layers per GPU of the weak-scaled run")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaled sub-record (N > 1, --scaling strong)")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--workload", default="track", choices=["track", "rosenthal", "multibeam", "pbf"],
                    help="track: BASELINE configs[2] (the benchmarked one); rosenthal: configs[1]; multibeam: configs[3] (use --nx 1000 --ny 250 --nz 200); "
                         "pbf: configs[4] (--nx 1000 --ny 500 --nz 400, 8 GPUs)")
    ap.add_argument("--explicit", action="store_true", help="as-shipped mode of configs[0]: explicitSolve true, maxDi 1 (diagonal solves only)")
    ap.add_argument("--solver", default="PCG", choices=["PCG", "PBiCGStab"])
    ap.add_argument("--preconditioner", default="DIC", choices=["DIC", "DILU", "diagonal", "none"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-nz", type=int, default=0, help="z layers of the CPU sample (0 = auto)")
    ap.add_argument("--cpu-steps", type=int, default=4)
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
    if args.workload == "multibeam":
        return cases.multi_beam(n, **ctl)
    if args.workload == "pbf":
        ctl.pop("maxDi")                   # configs[4]: maxDi 100 (tutorials/multiLayerPBF/system/controlDict:52), Tmax 3300
        return cases.multi_layer_pbf(n, **ctl)
    return cases.synthetic_track(n, **ctl)


def workload_name(args, nz):
    kind = {"rosenthal": "conduction-only Rosenthal track", "multibeam": "four beams, two of them transient (BASELINE configs[3])",
            "pbf": "powder layer, raster scan, transient modifiedSuperGaussian beam, Tmax 3300 (BASELINE configs[4])"}.get(
        args.workload, "conduction+solidification track")
    if args.explicit:
        kind += ", explicitSolve true / maxDi 1 (as-shipped mode)"
    return (f"synthetic uniform-hex {args.nx}x{args.ny}x{nz} cells (20 um), {"single superGaussian beam, " if args.workload in ("track", "rosenthal") else ""}{kind}, implicit "
            f"{args.solver}/{args.preconditioner} tol 1e-15 maxIter 20" + (" (BASELINE configs[2])" if args.workload == "track" else ""))


def global_nz(args, world):
    return args.nz if args.scaling == "strong" else args.nz_per_gpu * world


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
    return dict(value=n * steps / dt, seconds=dt, cells=n, iterations=iters, steps=steps, cell_iterations_per_s=n * iters / dt)


def source_hash():
    """sha256 over the CUDA sources the library is built from: ties an ncu capture to the build it was taken on."""
    import glob
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "additivefoam_b200", "csrc", "*.cu*"))):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(group, n_local):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the group's kernel, from profiles/r2_traffic.json -- written by
    tools/summarize_ncu.py from an `ncu --set full` capture of this command, together with the hash of the CUDA sources it was
    taken on.  Another build (any kernel source changed since) or a size that was not captured gives None: a stale number is
    worse than none."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(path):
        return None
    try:
        t = json.load(open(path))
    except Exception:
        return None
    if t.get("source_hash") != source_hash():
        return None
    return t.get("bytes_per_launch", {}).get(str(n_local), {}).get(group)


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def available_host_bytes():
    try:
        import psutil
        return int(psutil.virtual_memory().available)
    except Exception:
        return None


ORACLE_BYTES_PER_CELL = 360      # measured: 345 B/cell resident at 6.4 M cells (fields, LDU matrix and addressing, Krylov vectors)


def reference_arm(args, N):
    """The reference algorithm on the host cores, on the SAME global mesh as the GPU arm: every step is a full time step of the
    whole block.  Sub-domains are compact z-slabs, one per core (at least 4 layers thick), each with block-local DIC -- the way an
    `mpirun -np <cores> additiveFoam -parallel` run is decomposed (decomposeParDict:18-20)."""
    cores = host_cores()
    nz = global_nz(args, N)
    ranks = max(1, min(cores, nz // 4 if nz >= 4 else 1))
    note = ""
    avail = available_host_bytes()
    need = ORACLE_BYTES_PER_CELL * args.nx * args.ny * nz
    if args.cpu_sample_nz:
        nz_run = args.cpu_sample_nz
    elif avail is not None and need > 0.8 * avail:
        nz_run = max(4 * ranks, int(0.8 * avail / (ORACLE_BYTES_PER_CELL * args.nx * args.ny)))
        note = f"; host memory ({avail / 2**30:.0f} GiB free) holds only {nz_run} of the {nz} z layers"
    else:
        nz_run = nz
    ranks = max(1, min(ranks, nz_run // 4 if nz_run >= 4 else 1))
    r = run_cpu(args, nz_run, args.steps, args.warmup, ranks)
    full = nz_run == nz
    sample = (f"{'the full mesh' if full else 'a sub-slab'}: {args.nx}x{args.ny}x{nz_run} cells, {args.steps} time steps after {args.warmup} warm-up, "
              f"{ranks} compact z-slab sub-domains of ~{nz_run / ranks:.1f} layers on {ranks} threads, {r['seconds']:.1f} s{note}")
    return {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": N, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, nz), "cells": r["cells"], "krylov_iterations": r["iterations"],
                       "cell_iterations_per_s": r["cell_iterations_per_s"], "same_mesh_as_gpu_arm": full,
                       "note": "restatement of the reference algorithm (CPU oracle, OpenMP over z-slab ranks), not the OpenFOAM "
                               "binary: OpenFOAM-10 is not installable here"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": ranks, "kind": "port", "sample": sample},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def parity_check(ctx, dist, rank, world):
    """A small case decomposed over the SAME ranks, 12 time steps, against the oracle's in-process emulation of `world` ranks
    (rank 0 runs the oracle).  Puts multi-GPU parity on the bench line, where the driver's SCALE run can see it."""
    import numpy as np
    from additivefoam_b200 import abi, cases, lib, partition
    n = (60, 16, max(12, 3 * world))
    case = cases.amb2018_02_b(n=n, explicitSolve=0, maxDi=10.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC)
    g = lib.Case(ctx, case)
    reps = [g.step() for _ in range(12)]
    T = partition.gather_field(g.field(abi.FIELD_T), n, rank, world, dist, device="cuda")
    A = partition.gather_field(g.field(abi.FIELD_ALPHA_SOLID), n, rank, world, dist, device="cuda")
    g.close()
    out = None
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol
        o = ol.OracleCase(case, ranks=world)
        oreps = [o.step() for _ in range(12)]
        To, Ao = o.field(abi.FIELD_T), o.field(abi.FIELD_ALPHA_SOLID)
        eT = float(np.max(np.abs(T - To)) / np.max(np.abs(To)))
        eA = float(np.max(np.abs(A - Ao)))
        its = max(abs(a.lastSolveIterations - b.lastSolveIterations) for a, b in zip(reps, oreps))
        corr = max(abs(a.nThermoCorr - b.nThermoCorr) for a, b in zip(reps, oreps))
        o.close()
        out = {"ranks": world, "case": f"AMB2018-02-B {n[0]}x{n[1]}x{n[2]} implicit PCG/DIC, 12 steps, vs the oracle's {world}-rank emulation",
               "relLinf_T": eT, "Linf_alpha": eA, "iter_diff": its, "corrector_diff": corr, "Tmax": float(To.max()),
               "ok": bool(eT < 1e-9 and eA < 1e-9 and its <= 1 and corr == 0)}
    return out


def timed_run(args, ctx, dist, torch, world, nz, steps, warmup, profile):
    """W warm-up + K timed steps of the case with `nz` global layers; device-timed, max over ranks."""
    from additivefoam_b200 import lib
    case = make_case(args, nz)
    g = lib.Case(ctx, case)

    def barrier():
        ctx.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        g.step()
    barrier()
    if profile:
        ctx.profile_enable(True)
    launches0 = ctx.kernel_launches()
    ctx.mark(0)
    iters = correctors = 0
    last = None
    for _ in range(steps):
        last = g.step()
        iters += last.nSolveIterations
        correctors += last.nThermoCorr
    ctx.mark(1)
    ms = ctx.elapsed_ms(0, 1)
    barrier()
    launches = ctx.kernel_launches() - launches0
    prof = None
    if profile:
        prof = ctx.profile_read()
        ctx.profile_enable(False)
    if world > 1:
        tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    return dict(case=case, g=g, ms=ms, iters=iters, correctors=correctors, last=last, launches=launches, prof=prof, barrier=barrier)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    N = max(args.gpus, 1)

    if args.impl == "reference":
        if rank != 0:
            return 0
        print(json.dumps(reference_arm(args, N)))
        return 0

    os.environ.setdefault("NCCL_DEBUG", "WARN")       # keep stdout to the one JSON line
    import numpy as np  # noqa: F401
    import torch
    from additivefoam_b200 import abi, lib

    uid = None
    dist = None
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
        partition.enable_peer(ctx, dist, device="cuda", halo_doubles=args.nx * args.ny)

    parity = None
    if world > 1 and not args.no_parity_check:
        parity = parity_check(ctx, dist, rank, world)

    nz = global_nz(args, world)
    sampler = ClockSampler(local)
    sampler.start()
    run = timed_run(args, ctx, dist, torch, world, nz, args.steps, args.warmup, profile=True)
    sampler.stop_flag.set()
    sampler.join(2)
    g, ms, iters, correctors, last, prof = run["g"], run["ms"], run["iters"], run["correctors"], run["last"], run["prof"]
    barrier = run["barrier"]
    n_global = g.n_global
    value = n_global * args.steps / (ms * 1e-3)

    # ---- end to end through the C ABI with HOST fields (see e2e_run)
    e2e = None
    if not args.no_e2e:
        e2e = e2e_run(args, ctx, run, abi, torch, dist, world, barrier, n_global)
    nl = g.n_local
    nzl = nl // (args.nx * args.ny)
    g.close()

    # ---- the weak-scaled companion number (N slabs of --nz-per-gpu layers each), fewer steps
    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_weak:
        wsteps, wwarm = min(args.steps, 8), min(args.warmup, 3)
        wr = timed_run(args, ctx, dist, torch, world, args.nz_per_gpu * world, wsteps, wwarm, profile=False)
        wn = wr["g"].n_global
        weak = {"value": wn * wsteps / (wr["ms"] * 1e-3), "unit": UNIT, "cells": wn, "ms_per_step": wr["ms"] / wsteps, "steps": wsteps,
                "warmup": wwarm, "what": f"{world} z-slabs of {args.nx}x{args.ny}x{args.nz_per_gpu} cells (weak scaling)"}
        wr["g"].close()

    if rank != 0:
        ctx.close()
        dist.barrier()
        dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel group (CUDA events around every launch, live in this run)
    d = run["case"]["mesh"]["n"]
    nf = (d[0] - 1) * d[1] * nzl + d[0] * (d[1] - 1) * nzl + d[0] * d[1] * (nzl - 1)
    alg_bytes = {  # SURVEY.md 8(d), per launch
        "amul": 24 * nl + 16 * nf, "sweep_fwd": 24 * nl + 16 * nf, "sweep_bwd": 24 * nl + 16 * nf, "factor": 16 * nl + 16 * nf,
        "vector": None, "normfactor": 24 * nl + 8 * nl + 16 * nf + 24 * nl, "assemble": 48 * nl + 16 * nf, "corrector": 72 * nl,
        "alpha": 48 * nl, "props": 44 * nl, "dinum": 24 * nl + 8 * nf, "faces": 8 * nl + 8 * nf}
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
                "group_ms_per_launch": {k: round(v, 4) for k, v in per_launch_ms.items()},
                "group_GBps": {k: round(alg_bytes[k] / (per_launch_ms[k] * 1e-3) / 1e9, 1) for k in per_launch_ms if alg_bytes.get(k)},
                "note": "achieved = SURVEY.md 8(d) algorithmic bytes (general LDU addressing: 72 B/cell for Amul and for each sweep) / "
                        "measured launch time; the block path moves less (Amul 48, forward sweep 56, backward sweep 64 B/cell: no "
                        "addressing arrays, see DESIGN.md section 5), so compare `traffic` / avg_launch_ms with the peak for the DRAM rate; "
                        "traffic is null unless profiles/r2_traffic.json was captured on this very build"}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cores = host_cores()
        ranks = max(1, min(cores, nz // 4))
        cnz = args.cpu_sample_nz or min(nz, 4 * ranks)
        try:
            r = run_cpu(args, cnz, args.cpu_steps, 2, ranks)
            cpu = {"value": r["value"], "unit": UNIT, "cores": ranks, "kind": "port",
                   "sample": f"{args.nx}x{args.ny}x{cnz} cells ({cnz}/{nz} of the mesh), {args.cpu_steps} steps after 2 warm-up, {ranks} compact z-slab "
                             f"sub-domains, {r['seconds']:.1f} s; restatement of the reference algorithm, not the OpenFOAM binary "
                             "(`--impl reference` times the full mesh)"}
        except Exception as e:     # the oracle is only the reported baseline
            cpu = {"value": None, "unit": UNIT, "cores": ranks, "kind": "port", "sample": f"failed: {e}"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling if world > 1 else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, nz), "cells": n_global, "cells_per_gpu": nl, "decomposition": f"{world} z-slab(s)",
                       "krylov_iterations": iters, "correctors": correctors,
                       "cell_iterations_per_s": n_global * iters / (ms * 1e-3), "l2": "inputs larger than L2 (>= 100 MB per field)",
                       "last_step": last.as_dict()},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": run["launches"], "roofline": roofline, "cpu_baseline": cpu}
    if weak is not None:
        line["weak"] = weak
    if parity is not None:
        line["parity_check"] = parity
    print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def e2e_run(args, ctx, run, abi, torch, dist, world, barrier, n_global):
    """The same metric through the C ABI with HOST buffers: every step's inputs (T, alpha.solid, alpha.powder) come from pinned host
    memory and its results (T, alpha.solid) go back to it, inside the timed region.  The host owns the fields between steps, so a
    case's step n+1 starts from what its step n sent to the host -- the strict chain download -> upload -> step.  Two figures:

    * `value`: TWO independent cases of the benchmarked configuration share the GPU (the timed case and a second one built from the
      same description), stepped alternately.  While one case computes, the other's results cross to its host buffers and come back
      as its next inputs on the two copy streams (b200_case_get_field_async / set_field_async / apply_staged), so PCIe runs under
      the kernels.  Every case does exactly the work of the device-timed run's steps.
    * `serial`: one case, the same bytes with every call blocking (b200_case_set_field / step / get_field)."""
    from additivefoam_b200 import lib
    F_T, F_A, F_P = abi.FIELD_T, abi.FIELD_ALPHA_SOLID, abi.FIELD_ALPHA_POWDER
    dbg = os.environ.get("B200_E2E_DEBUG") == "1"
    e_steps = max(2, args.steps // 2)

    def reduce_max(x):
        if world > 1:
            tm = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            return float(tm.item())
        return x

    def host_fields(g):
        return {f: torch.from_numpy(g.field(f)).pin_memory().numpy() for f in (F_T, F_A, F_P)}

    gA = run["g"]
    hA = host_fields(gA)
    # serial
    s_steps = max(2, e_steps // 2)
    barrier()
    t0 = time.perf_counter()
    for _ in range(s_steps):
        gA.set_field(F_P, hA[F_P])
        gA.set_field_keep_old(F_T, hA[F_T])
        gA.set_field_keep_old(F_A, hA[F_A])
        gA.step()
        gA.field_into(F_T, hA[F_T])
        gA.field_into(F_A, hA[F_A])
    barrier()
    st = reduce_max(time.perf_counter() - t0)

    # two cases, alternating
    gB = lib.Case(ctx, run["case"])
    for _ in range(args.warmup):
        gB.step()
    hB = host_fields(gB)
    cases = [(gA, hA), (gB, hB)]

    def results_out_and_back(g, h):          # queued, returns at once: the results to the host buffers, the host buffers back as inputs
        g.get_field_async(F_T, h[F_T]); g.get_field_async(F_A, h[F_A])
        g.set_field_async(F_P, h[F_P]); g.set_field_async(F_T, h[F_T]); g.set_field_async(F_A, h[F_A])

    def round_(i):
        (gx, _), (gy, hy) = cases[i & 1], cases[1 - (i & 1)]
        ta = time.perf_counter()
        results_out_and_back(gy, hy)         # the case that stepped last: its PCIe traffic runs under this step
        gx.apply_staged()                    # this case's inputs came back during the other case's step
        tb = time.perf_counter()
        rep = gx.step()
        if dbg:
            print(f"e2e round {i}: queueing {1e3 * (tb - ta):.1f} ms, step {1e3 * (time.perf_counter() - tb):.1f} ms, "
                  f"{rep.nSolveIterations} Krylov iterations, {rep.nThermoCorr} correctors", file=sys.stderr, flush=True)

    # untimed: staging buffers get allocated, both cases reach the steady state of the loop (A staged, B stepped last)
    round_(1)
    round_(0)
    round_(1)
    gA.transfers_wait(); gB.transfers_wait()
    barrier()
    t0 = time.perf_counter()
    for i in range(2 * e_steps):
        round_(i)
    gA.transfers_wait(); gB.transfers_wait()
    barrier()
    et = reduce_max(time.perf_counter() - t0)
    gB.close()
    nl = gA.n_local
    return {"value": n_global * 2 * e_steps / et, "unit": UNIT, "h2d_bytes_per_step": 3 * 8 * nl * world,
            "d2h_bytes_per_step": 2 * 8 * nl * world, "steps": 2 * e_steps, "cases": 2,
            "what": "two independent cases of this configuration share the GPU(s) and step alternately; while one computes, the other's "
                    "results (T, alpha.solid) go to pinned host memory and come back with alpha.powder as its next inputs "
                    "(b200_case_get_field_async -> set_field_async -> apply_staged -> b200_case_step); all copies inside the timed region",
            "serial": {"value": n_global * s_steps / st, "steps": s_steps,
                       "what": "one case: b200_case_set_field x3 -> b200_case_step -> b200_case_get_field x2, each blocking"}}


if __name__ == "__main__":
    sys.exit(main())
