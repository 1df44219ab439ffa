"""GPU parity of the lduMatrix path against the CPU oracle, through the C ABI (b200_ldu_*)."""
import ctypes as C

import numpy as np
import pytest

import helpers
import oracle_lib as ol
from additivefoam_b200 import abi, lib

pytestmark = pytest.mark.gpu

P = ol.P


def oracle_amul(L, n, lo, up, diag, upper, lower, x):
    out = np.empty(n)
    L.afo_ldu_amul(n, len(lo), P(lo), P(up), P(diag), P(upper), P(lower), P(x), P(out))
    return out


def oracle_precond(L, n, lo, up, diag, upper, lower, kind, r):
    out = np.empty(n)
    L.afo_ldu_precondition(n, len(lo), P(lo), P(up), P(diag), P(upper), P(lower), kind, P(r), P(out), None)
    return out


def oracle_solve(L, n, lo, up, diag, upper, lower, b, x0, solver, precond, tol, relTol, minIter, maxIter):
    x = x0.copy()
    perf = abi.SolverPerf()
    log = np.empty(4096)
    nlog = C.c_int32()
    L.afo_ldu_solve(n, len(lo), P(lo), P(up), P(diag), P(upper), P(lower), P(b), P(x), solver, precond, tol, relTol,
                    minIter, maxIter, C.byref(perf), P(log), 4096, C.byref(nlog))
    return x, perf, log[: nlog.value].copy()


def problems():
    out = []
    for n, extra, seed in [(1, 0, 1), (2, 0, 2), (57, 40, 3), (1000, 2500, 4), (20000, 50000, 5)]:
        lo, up = helpers.random_ldu_addressing(n, extra, seed) if n > 1 else (np.zeros(0, np.int32), np.zeros(0, np.int32))
        out.append(("general", n, lo, up, None))
    for dims in [(1, 1, 1), (5, 1, 1), (1, 4, 3), (7, 5, 3), (33, 17, 9), (150, 25, 15), (300, 7, 2)]:
        lo, up = helpers.block_addressing_np(dims)
        out.append(("block", dims[0] * dims[1] * dims[2], lo, up, dims))
    return out


@pytest.mark.parametrize("kind,n,lo,up,dims", problems(), ids=lambda v: str(v) if isinstance(v, (str, int, tuple)) else "")
@pytest.mark.parametrize("asym", [False, True])
def test_amul_bit_exact(ctx, oracle, kind, n, lo, up, dims, asym):
    diag, upper, lower, x, _ = helpers.spd_coefficients(n, lo, up, seed=11, asym=asym)
    s = lib.LduSolver(ctx, lo, up, n, block=dims)
    got = s.amul(x, diag, upper, lower)
    ref = oracle_amul(oracle, n, lo, up, diag, upper, lower, x)
    s.close()
    # same per-cell summation order, no FMA contraction on either side: bit-exact
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("kind,n,lo,up,dims", problems(), ids=lambda v: str(v) if isinstance(v, (str, int, tuple)) else "")
@pytest.mark.parametrize("precond", [abi.PRECOND_NONE, abi.PRECOND_DIAGONAL, abi.PRECOND_DIC, abi.PRECOND_DILU])
def test_precondition_bit_exact(ctx, oracle, kind, n, lo, up, dims, precond):
    asym = precond == abi.PRECOND_DILU
    diag, upper, lower, x, b = helpers.spd_coefficients(n, lo, up, seed=12, asym=asym)
    s = lib.LduSolver(ctx, lo, up, n, block=dims)
    got = s.precondition(precond, b, diag, upper, lower)
    ref = oracle_precond(oracle, n, lo, up, diag, upper, lower, precond, b)
    s.close()
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("kind,n,lo,up,dims", problems(), ids=lambda v: str(v) if isinstance(v, (str, int, tuple)) else "")
@pytest.mark.parametrize("solver,precond,asym", [
    (abi.SOLVER_PCG, abi.PRECOND_DIC, False), (abi.SOLVER_PCG, abi.PRECOND_DIAGONAL, False),
    (abi.SOLVER_PCG, abi.PRECOND_NONE, False), (abi.SOLVER_PBICGSTAB, abi.PRECOND_DILU, True),
    (abi.SOLVER_PBICGSTAB, abi.PRECOND_DILU, False), (abi.SOLVER_PBICGSTAB, abi.PRECOND_DIAGONAL, True),
    (abi.SOLVER_PBICGSTAB, abi.PRECOND_NONE, True), (abi.SOLVER_PBICGSTAB, abi.PRECOND_NONE, False)])
def test_solve_parity(ctx, oracle, kind, n, lo, up, dims, solver, precond, asym):
    diag, upper, lower, x0, b = helpers.spd_coefficients(n, lo, up, seed=13, asym=asym)
    s = lib.LduSolver(ctx, lo, up, n, block=dims)
    ctl = dict(tol=1e-10, relTol=0.0, minIter=1, maxIter=60)
    got, perf = s.solve(x0, b, diag, upper, lower, solver=solver, preconditioner=precond, tolerance=ctl["tol"],
                        relTol=ctl["relTol"], minIter=ctl["minIter"], maxIter=ctl["maxIter"])
    ref, rperf, _ = oracle_solve(oracle, n, lo, up, diag, upper, lower, b, x0, solver, precond, ctl["tol"], ctl["relTol"],
                                 ctl["minIter"], ctl["maxIter"])
    s.close()
    assert abs(perf.nIterations - rperf.nIterations) <= 1
    assert perf.initialResidual == pytest.approx(rperf.initialResidual, rel=1e-10)
    # north_star tolerance; unpreconditioned BiCGStab amplifies the two sides' differently ordered global sums (its residual is
    # not monotone): same iteration counts and residuals, iterates within 1e-8 after 60 unconverged iterations
    assert helpers.rel_linf(got, ref) < (1e-7 if (solver, precond) == (abi.SOLVER_PBICGSTAB, abi.PRECOND_NONE) else 1e-9)
    # and it actually solves the system
    A = helpers.dense_from_ldu(n, lo, up, diag, upper, lower) if n <= 1000 else None
    if A is not None and perf.converged:
        assert np.linalg.norm(A @ got - b) <= 1e-6 * max(np.linalg.norm(b), 1.0)


def test_as_shipped_controls_hit_max_iter(ctx, oracle):
    """tolerance 1e-15 / maxIter 20 (fvSolution:38-41): both sides run maxIter+1 iterations (post-increment loop test)."""
    dims = (40, 20, 10)
    lo, up = helpers.block_addressing_np(dims)
    n = dims[0] * dims[1] * dims[2]
    diag, upper, lower, x0, b = helpers.spd_coefficients(n, lo, up, seed=14)
    diag = diag * 0.02 + (diag - diag * 0.02) * 0.0 - 0.0   # keep it SPD but poorly conditioned
    diag = -np.bincount(lo, upper, n) - np.bincount(up, upper, n) + 1e-3
    s = lib.LduSolver(ctx, lo, up, n, block=dims)
    got, perf = s.solve(x0, b, diag, upper, None, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC, tolerance=1e-15,
                        relTol=0.0, minIter=1, maxIter=20)
    ref, rperf, _ = oracle_solve(oracle, n, lo, up, diag, upper, None, b, x0, abi.SOLVER_PCG, abi.PRECOND_DIC, 1e-15, 0.0, 1, 20)
    s.close()
    assert rperf.nIterations == 21 and perf.nIterations == 21
    assert helpers.rel_linf(got, ref) < 1e-9


def test_levels_bit_exact(ctx, oracle):
    for kind, n, lo, up, dims in problems():
        s = lib.LduSolver(ctx, lo, up, n, block=dims)
        lv, nl = s.levels()
        ref = np.empty(n, dtype=np.int32)
        rn = C.c_int32()
        oracle.afo_ldu_levels(n, len(lo), P(lo), P(up), P(ref), C.byref(rn))
        s.close()
        assert nl == rn.value and np.array_equal(lv, ref)


def test_set_block_rejects_wrong_addressing(ctx):
    lo, up = helpers.block_addressing_np((6, 5, 4))
    s = lib.LduSolver(ctx, lo, up, 120)
    with pytest.raises(lib.B200Error):
        s.L.b200_ldu_set_block  # noqa: B018
        lib._check(s.L.b200_ldu_set_block(s.h, 5, 6, 4))
    lib._check(s.L.b200_ldu_set_block(s.h, 6, 5, 4))
    s.close()


def test_shared_reciprocal_division_is_the_compilers_division(ctx):
    """The factor sweep forms upper*upper/rD with one reciprocal refinement per rD entry (wave_kernels.cuh: div_refine, div_by).
    That is nvcc's own fast path for a/b split in two, so every quotient must carry the bit pattern of the operator -- checked
    here on 2^32 pseudo-random operand pairs over the exponent range the sweep uses it for, and on a narrow range around 1."""
    assert ctx.selftest_division(1 << 32, seed=12345, exp_range=250) == 0
    assert ctx.selftest_division(1 << 30, seed=777, exp_range=3) == 0
    assert ctx.selftest_division(1 << 28, seed=99, exp_range=0) == 0


@pytest.mark.parametrize("scale_coef,scale_diag", [(2.0 ** 140, 2.0 ** 140), (2.0 ** -140, 2.0 ** -140), (1.0, 2.0 ** -260), (1.0, 2.0 ** 260)])
@pytest.mark.parametrize("precond", [abi.PRECOND_DIC, abi.PRECOND_DILU])
def test_factor_sweep_exact_path_outside_fast_division_range(ctx, oracle, scale_coef, scale_diag, precond):
    """The factor sweep's shared-reciprocal divisions are only used while exponents stay within 2^+-125 (coefficients) and
    2^+-250 (diagonal entries); outside, the same sweep reruns with the plain FP64 operators.  Both ways the result is the
    oracle's bit for bit: scaled coefficients trip the veto at gather time, a scaled diagonal (off-diagonals switched off) the
    in-sweep range report."""
    dims = (37, 19, 9)
    lo, up = helpers.block_addressing_np(dims)
    n = dims[0] * dims[1] * dims[2]
    asym = precond == abi.PRECOND_DILU
    diag, upper, lower, x, b = helpers.spd_coefficients(n, lo, up, seed=21, asym=asym)
    if scale_coef == 1.0:
        upper = upper * 0.0
        lower = None if lower is None else lower * 0.0
        diag = diag * scale_diag
    else:
        diag, upper = diag * scale_diag, upper * scale_coef
        lower = None if lower is None else lower * scale_coef
    s = lib.LduSolver(ctx, lo, up, n, block=dims)
    for _ in range(2):          # the second call reuses the gathered coefficients (and their veto)
        got = s.precondition(precond, b, diag, upper, lower)
    ref = oracle_precond(oracle, n, lo, up, diag, upper, lower, precond, b)
    # and an in-range matrix on the same solver afterwards goes back to the fast path with the right answer
    d2, u2, l2, _, b2 = helpers.spd_coefficients(n, lo, up, seed=22, asym=asym)
    got2 = s.precondition(precond, b2, d2, u2, l2)
    ref2 = oracle_precond(oracle, n, lo, up, d2, u2, l2, precond, b2)
    s.close()
    assert np.array_equal(got, ref)
    assert np.array_equal(got2, ref2)
