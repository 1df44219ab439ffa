#!/usr/bin/env python
"""Where does the streamed e2e figure lose its overlap?  One GPU, the bench's default case.  Per variant: wall time per step, the
device-timed group totals of the step's kernels (do the kernels get slower, or the gaps between them longer?), and the raw PCIe rates
of the same buffers alone and in both directions at once.

    python tools/e2e_probe.py [bench.py arguments]
"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    args = bench.parse()
    import torch
    from additivefoam_b200 import abi, lib
    ctx = lib.Context(0, None, 0, 1)
    case = bench.make_case(args, bench.global_nz(args, 1))
    g = lib.Case(ctx, case)
    for _ in range(3):
        g.step()
    ctx.synchronize()
    nl = g.n_local
    hT = torch.from_numpy(g.field(abi.FIELD_T)).pin_memory().numpy()
    hA = torch.from_numpy(g.field(abi.FIELD_ALPHA_SOLID)).pin_memory().numpy()
    hP = torch.from_numpy(g.field(abi.FIELD_ALPHA_POWDER)).pin_memory().numpy()
    oT = [torch.empty(nl, dtype=torch.float64).pin_memory().numpy() for _ in range(2)]
    oA = [torch.empty(nl, dtype=torch.float64).pin_memory().numpy() for _ in range(2)]

    # raw PCIe with torch's own copies: one direction, then both at once
    dT = torch.empty(nl, dtype=torch.float64, device="cuda")
    dA = torch.empty(nl, dtype=torch.float64, device="cuda")
    tT, tO = torch.from_numpy(hT), torch.from_numpy(oT[0])
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for what in ("h2d", "d2h", "both"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if what in ("h2d", "both"):
            with torch.cuda.stream(s1):
                dT.copy_(tT, non_blocking=True)
        if what in ("d2h", "both"):
            with torch.cuda.stream(s2):
                tO.copy_(dA, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"raw {what}: {8 * nl / 1e9:.2f} GB per direction in {1e3 * dt:.1f} ms = {8 * nl / dt / 1e9:.1f} GB/s per direction", flush=True)

    def run(name, up, down, n=5):
        ctx.synchronize()
        ctx.profile_enable(True)
        ctx.profile_read()
        walls = []
        if up:
            g.set_field_async(abi.FIELD_ALPHA_POWDER, hP); g.set_field_async(abi.FIELD_T, hT); g.set_field_async(abi.FIELD_ALPHA_SOLID, hA)
            g.transfers_wait()
        t00 = time.perf_counter()
        for i in range(n):
            t0 = time.perf_counter()
            if up:
                g.apply_staged()
                g.set_field_async(abi.FIELD_ALPHA_POWDER, hP); g.set_field_async(abi.FIELD_T, hT); g.set_field_async(abi.FIELD_ALPHA_SOLID, hA)
            rep = g.step()
            if down:
                g.get_field_async(abi.FIELD_T, oT[i & 1]); g.get_field_async(abi.FIELD_ALPHA_SOLID, oA[i & 1])
            walls.append((1e3 * (time.perf_counter() - t0), rep.nSolveIterations, rep.nThermoCorr))
        g.transfers_wait()
        total = 1e3 * (time.perf_counter() - t00)
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        ksum = sum(v[0] for v in prof.values())
        print(f"== {name}: total {total:.1f} ms over {n} steps = {total / n:.1f} ms/step; kernels {ksum / n:.1f} ms/step", flush=True)
        print("   per step (wall ms, Krylov iterations, correctors): " + ", ".join(f"({w:.1f}, {it}, {c})" for w, it, c in walls), flush=True)
        print("   groups ms/step: " + ", ".join(f"{k} {v[0] / n:.2f}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]) if v[0] > 0.05 * n),
              flush=True)

    run("steps alone", False, False)
    run("upload under the step", True, False)
    run("download under the step", False, True)
    run("both", True, True)
    run("steps alone again", False, False)


if __name__ == "__main__":
    main()
