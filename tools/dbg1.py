import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, helpers
from additivefoam_b200 import abi, lib
ctx = lib.Context(0)
for dims in [(1,1,1),(2,1,1),(5,1,1),(1,4,3)]:
    lo, up = helpers.block_addressing_np(dims)
    n = dims[0]*dims[1]*dims[2]
    diag, upper, lower, x0, b = helpers.spd_coefficients(n, lo, up, seed=13)
    for rep in range(3):
        s = lib.LduSolver(ctx, lo, up, n, block=dims)
        got, perf = s.solve(x0, b, diag, upper, None, solver=0, preconditioner=2, tolerance=1e-10, relTol=0.0, minIter=1, maxIter=60)
        A = helpers.dense_from_ldu(n, lo, up, diag, upper)
        print(dims, rep, perf.as_dict(), np.abs(A@got-b).max(), np.abs(got-np.linalg.solve(A,b)).max())
        w = s.precondition(2, b, diag, upper)
        print('   precond', w[:3], (b/diag)[:3] if n==1 else '')
        s.close()
