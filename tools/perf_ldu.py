"""Micro-benchmark / profiling driver for the block LDU kernels (Amul, DIC sweeps, PCG) on a seeded SPD matrix
(SURVEY.md 8d micro-bench).  Usage: python tools/perf_ldu.py NX NY NZ [reps] [what: precond|amul|solve] [block|general]
`general` hands the same matrix over without the block declaration: the arbitrary-addressing kernels a decomposed OpenFOAM case gets."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402  (addressing generator only)
from additivefoam_b200 import abi, lib  # noqa: E402

nx, ny, nz = (int(v) for v in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
what = sys.argv[5] if len(sys.argv) > 5 else "precond"
addressing = sys.argv[6] if len(sys.argv) > 6 else "block"
O = ol.load()
n, nf, lo, up = ol.block_addressing(O, (nx, ny, nz))
rng = np.random.default_rng(1234)
upper = -rng.uniform(0.5, 1.5, nf)
diag = rng.uniform(0.1, 1.0, n)
np.subtract.at(diag, lo, upper)
np.subtract.at(diag, up, upper)
b = rng.uniform(-1, 1, n)
x = rng.uniform(-1, 1, n)
ctx = lib.Context(0)
s = lib.LduSolver(ctx, lo, up, n, block=(nx, ny, nz) if addressing == "block" else None)
for r in range(reps + 1):
    if r == 1:
        ctx.profile_enable(True)      # the first call allocates; time from the second on
    if what == "precond":
        s.precondition(abi.PRECOND_DIC, b, diag, upper)
    elif what == "amul":
        s.amul(x, diag, upper)
    else:
        s.solve(x, b, diag, upper, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC, tolerance=1e-15, minIter=1, maxIter=20)
prof = ctx.profile_read()
bytes_ = {"amul": 24 * n + 16 * nf, "sweep_fwd": 24 * n + 16 * nf, "sweep_bwd": 24 * n + 16 * nf, "factor": 16 * n + 16 * nf}
for k, (ms, cnt) in prof.items():
    if cnt:
        gb = f"{bytes_[k] / (ms / cnt * 1e-3) / 1e9:8.1f} GB/s(alg)" if k in bytes_ else ""
        print(f"{k:12s} {cnt:5d} launches  {ms / cnt:9.4f} ms/launch  {gb}")
