"""Sweep-only timing at large sizes with cheap host-side setup.  Usage: python tools/perf_wave.py NX NY NZ [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402  (addressing generator only)
from additivefoam_b200 import abi, lib  # noqa: E402

nx, ny, nz = (int(v) for v in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
n, nf, lo, up = ol.block_addressing(ol.load(), (nx, ny, nz))
rng = np.random.default_rng(1234)
upper = -rng.uniform(0.5, 1.5, nf)
diag = np.full(n, 9.5)
b = rng.uniform(-1, 1, n)
ctx = lib.Context(0)
s = lib.LduSolver(ctx, lo, up, n, block=(nx, ny, nz))
bytes_ = {"sweep_fwd": 24 * n + 16 * nf, "sweep_bwd": 24 * n + 16 * nf, "factor": 16 * n + 16 * nf}
variants = [v for v in sys.argv[5:]] or [""]          # each: "KEY=VAL,KEY=VAL"
s.precondition(abi.PRECOND_DIC, b, diag, upper)        # allocates
for var in variants:
    for k in [k for k in os.environ if k.startswith("B200_")]:
        del os.environ[k]
    for kv in filter(None, var.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    s.precondition(abi.PRECOND_DIC, b, diag, upper)
    ctx.profile_enable(True)
    for r in range(reps):
        s.precondition(abi.PRECOND_DIC, b, diag, upper)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    print(f"{var or 'default':40s}", " ".join(f"{k}={ms / cnt:.4f}ms({bytes_[k] / (ms / cnt * 1e-3) / 1e9:.0f}GB/s)" for k, (ms, cnt) in prof.items() if cnt and k in bytes_), flush=True)
