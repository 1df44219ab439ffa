"""The CPU oracle against the known answers derived from the reference formulas and against
independent dense / analytic solutions.  The reference ships no golden vectors, so these anchors
are what pins the oracle ("parity unpinned" at the OpenFOAM-10 boundary: see oracle/af_ldu.hpp)."""
import ctypes as C
import math

import numpy as np
import pytest

import helpers
import oracle_lib as ol
from additivefoam_b200 import abi, cases

P = ol.P
PATH = [[1, 0.0, 0.0, 0.0, 0.0, 0.0], [0, 0.002, 0.0, 0.0, 179.2, 0.8]]


def dims3(v):
    a = np.array(v, dtype=np.float64)
    return P(a), a


# ---------------------------------------------------------------- heat source formulas
def test_super_gaussian_V0(oracle):
    """superGaussian.C:73-85, k = 2, dims (85, 85, 30) um."""
    b = cases.make_beam(cases.super_gaussian_beam(PATH))
    p, keep = dims3((85e-6, 85e-6, 30e-6))
    assert oracle.afo_shape_V0(C.byref(b), p) == pytest.approx(2.1335799723345844e-13, rel=1e-14)


def test_modified_super_gaussian_V0(oracle):
    """modifiedSuperGaussian.C:87-103, k = 7.95, m = 2.72, dims (40, 40, 30) um."""
    b = cases.make_beam(cases.modified_super_gaussian_beam(PATH))
    p, keep = dims3((40e-6, 40e-6, 30e-6))
    assert oracle.afo_shape_V0(C.byref(b), p) == pytest.approx(8.921206515601453e-14, rel=1e-13)


def test_projected_gaussian_k_follows_depth(oracle):
    """projectedGaussian.C:96-118: n = clamp(A*log2(max(dz/min(dx0,dy0),1)) + B, 0, 9), k = 2^n, V0 analytic."""
    spec = cases.projected_gaussian_beam(PATH, dims=(50e-6, 50e-6, 30e-6), A=2.0, B=1.0)
    b = cases.make_beam(spec)
    for dz in (30e-6, 100e-6, 400e-6):
        p, keep = dims3((50e-6, 50e-6, dz))
        n = min(max(2.0 * math.log2(max(dz / 50e-6, 1.0)) + 1.0, 0.0), 9.0)
        k = 2.0 ** n
        want = 0.5 * math.pi * 50e-6 * 50e-6 * dz * math.gamma(1.0 / k) / (k * 3.0 ** (1.0 / k))
        assert oracle.afo_shape_V0(C.byref(b), p) == pytest.approx(want, rel=1e-13)


def test_shape_weights_pointwise(oracle):
    b = cases.make_beam(cases.super_gaussian_beam(PATH, k=2.0))
    pd, keepd = dims3((85e-6, 85e-6, 30e-6))
    pv, keepv = dims3((85e-6 / math.sqrt(2.0), 0.0, 0.0))          # one scaled radius: exp(-1)
    assert oracle.afo_shape_weight(C.byref(b), pd, pv) == pytest.approx(math.exp(-1.0), rel=1e-13)
    m = cases.make_beam(cases.modified_super_gaussian_beam(PATH))
    pd, keepd = dims3((40e-6, 40e-6, 30e-6))
    pv, keepv = dims3((0.0, 0.0, 30e-6))                            # at the tip: outside (d.z < s.z fails)
    assert oracle.afo_shape_weight(C.byref(m), pd, pv) == 0.0
    pv, keepv = dims3((0.0, 0.0, 0.0))
    assert oracle.afo_shape_weight(C.byref(m), pd, pv) == 1.0


def test_kelly_absorption(oracle):
    """KellyAbsorption.C:67-99, cone, eta0 = 0.28, etaMin = 0.35; discontinuous at aspect ratio 1."""
    b = cases.make_beam(cases.modified_super_gaussian_beam(PATH))
    assert oracle.afo_absorption_eta(C.byref(b), 0.75) == 0.35
    assert oracle.afo_absorption_eta(C.byref(b), 1.0) == 0.35
    assert oracle.afo_absorption_eta(C.byref(b), 2.0) == pytest.approx(0.6453157893721592, rel=1e-14)
    assert oracle.afo_absorption_eta(C.byref(b), 1.0 + 1e-12) == pytest.approx(0.5054, abs=5e-5)
    spec = cases.modified_super_gaussian_beam(PATH)
    spec["kellyGeometry"] = abi.KELLY_CYLINDER
    c = cases.make_beam(spec)
    a, e0 = 3.0, 0.28
    th = math.atan(1 / a)
    F, G = 0.5 * (1 - math.cos(2 * th)), 0.5 / (1 + a)
    assert oracle.afo_absorption_eta(C.byref(c), a) == pytest.approx(e0 * (1 + (1 - e0) * (G - F)) / (1 - (1 - e0) * (1 - G)), rel=1e-14)


def test_constant_absorption_and_peak_qdot(oracle):
    b = cases.make_beam(cases.super_gaussian_beam(PATH))
    assert oracle.afo_absorption_eta(C.byref(b), 0.35) == 0.33
    assert 0.33 * 179.2 == pytest.approx(59.136)
    assert 59.136 / 2.1335799723345844e-13 == pytest.approx(2.7717e14, rel=1e-4)


# ---------------------------------------------------------------- beam kinematics
def move(oracle, spec, times, start=0.0, end=1.0):
    b = cases.make_beam(spec)
    t = np.array(times, dtype=np.float64)
    pos, pw, idx = np.empty(3 * len(t)), np.empty(len(t)), np.empty(len(t), dtype=np.int32)
    et = C.c_double()
    oracle.afo_beam_move(C.byref(b), start, end, len(t), P(t), P(pos), P(pw), P(idx), C.cast(C.byref(et), abi.c_double_p))
    return pos.reshape(-1, 3), pw, idx, et.value


def test_moving_beam_amb2018(oracle):
    """movingBeam.C:157-225: the shipped path is a zero-dwell point then a 2 mm line at 0.8 m/s (2.5 ms)."""
    pos, pw, idx, endTime = move(oracle, cases.super_gaussian_beam(PATH), [0.0, 5e-11, 1e-9, 1.25e-3, 2.5e-3, 3e-3], end=0.004)
    assert endTime == pytest.approx(2.5e-3)
    assert list(idx) == [2, 2, 2, 2, 2, 2]                      # the zero-dwell point source is skipped (findIndex)
    assert pw[0] == 0.0 and pw[1] == 0.0                          # power switches only after eps = 1e-10
    assert pw[2] == 179.2 and pw[3] == 179.2
    assert pos[3, 0] == pytest.approx(1e-3) and pos[4, 0] == pytest.approx(2e-3)
    assert pos[5, 0] == pytest.approx(2.4e-3)                    # linear extrapolation past the last segment


def test_moving_beam_dwell_and_raster(oracle):
    path = [[1, 0, 0, 0, 0, 0], [0, 1e-3, 0, 0, 100.0, 1.0], [1, 1e-3, 1e-4, 0, 0.0, 5e-4], [0, 0, 1e-4, 0, 100.0, 1.0]]
    pos, pw, idx, endTime = move(oracle, cases.super_gaussian_beam(path), [5e-4, 1.2e-3, 1.5e-3, 2.0e-3, 2.5e-3], end=1.0)
    assert endTime == pytest.approx(2.5e-3)
    assert pos[0, 0] == pytest.approx(5e-4) and pw[0] == 100.0
    assert list(idx[1:3]) == [3, 3] and pw[1] == 0.0 and pos[1, 1] == pytest.approx(1e-4)      # dwelling at the turn
    assert idx[3] == 4 and pos[3, 0] == pytest.approx(5e-4) and pw[3] == 100.0


def test_adjust_deltaT_hits_path_end(oracle):
    """movingBeam.C:228-256: the step is shrunk to land on the segment boundary, dilated at most 1 %."""
    b = cases.make_beam(cases.super_gaussian_beam(PATH))
    dt = oracle.afo_beam_adjust_deltaT(C.byref(b), 0.0, 0.004, 2.49e-3, 2.49e-3, 7e-6)
    assert dt == pytest.approx(1e-5 / 2)                          # 10 us left, 7 us steps -> two steps of 5 us
    dt = oracle.afo_beam_adjust_deltaT(C.byref(b), 0.0, 0.004, 2.49e-3, 2.49e-3, 9.95e-6)
    assert dt == pytest.approx(9.95e-6)                           # within the 1 % dilation: one step, unchanged


def test_scan_path_parse(oracle):
    text = "Mode X Y Z Power Param\n1 0.000 0.000 0 0 0\n\n0 0.002 0.000 0 179.2 0.8\n"
    out = np.empty(6 * 8)
    n = oracle.afo_scan_path_parse(text.encode(), P(out), 8)
    assert n == 2 and list(out[6:12]) == [0, 0.002, 0, 0, 179.2, 0.8]


def test_interpolateXY_descending_table(oracle):
    """interpolateXY.C:33-109 tolerates descending abscissae; Tliq/Tsol come from swapping the columns (createFields.H:59-71)."""
    T, a = np.array([1410.0, 1620.0]), np.array([1.0, 0.0])
    f = lambda x, X, Y: oracle.afo_interpolateXY(x, len(X), P(X), P(Y))      # noqa: E731
    assert f(0.0, a, T) == 1620.0 and f(1.0, a, T) == 1410.0
    assert f(1515.0, T, a) == pytest.approx(0.5)
    assert f(300.0, T, a) == 1.0 and f(3000.0, T, a) == 0.0


# ---------------------------------------------------------------- mesh / addressing
@pytest.mark.parametrize("n,cells,faces", [((150, 25, 15), 56250, 162375), ((200, 100, 1), 20000, 39700)])
def test_mesh_counts(oracle, n, cells, faces):
    nc, nf, lo, up = ol.block_addressing(oracle, n)
    assert (nc, nf) == (cells, faces)
    assert np.all(lo < up) and np.all(np.diff(lo) >= 0)           # upper-triangular order
    same = lo[1:] == lo[:-1]
    assert np.all(up[1:][same] > up[:-1][same])


def test_addressing_matches_numpy(oracle):
    for n in [(4, 3, 2), (1, 5, 1), (7, 1, 3)]:
        nc, nf, lo, up = ol.block_addressing(oracle, n)
        lo2, up2 = helpers.block_addressing_np(n)
        assert np.array_equal(lo, lo2) and np.array_equal(up, up2)


def test_slab_decomposition_addressing(oracle):
    """Rank r of R owns global k in [nz*r/R, nz*(r+1)/R); local order = global order restricted to the rank."""
    n, R = (5, 4, 7), 3
    total = 0
    for r in range(R):
        nc, nf, lo, up = ol.block_addressing(oracle, n, r, R)
        nzl = 7 * (r + 1) // R - 7 * r // R
        lo2, up2 = helpers.block_addressing_np((5, 4, nzl))
        assert nc == 20 * nzl and np.array_equal(lo, lo2) and np.array_equal(up, up2)
        total += nc
    assert total == 140


def test_levels_are_hyperplanes(oracle):
    n = (6, 5, 4)
    nc, nf, lo, up = ol.block_addressing(oracle, n)
    lv = np.empty(nc, dtype=np.int32)
    nl = C.c_int32()
    oracle.afo_ldu_levels(nc, nf, P(lo), P(up), P(lv), C.byref(nl))
    i, j, k = np.meshgrid(range(6), range(5), range(4), indexing="ij")
    want = (i + j + k).transpose(2, 1, 0).ravel()
    assert nl.value == 6 + 5 + 4 - 2 and np.array_equal(lv, want)


# ---------------------------------------------------------------- LDU kernels against dense algebra
def test_amul_and_preconditioners_against_dense(oracle):
    n = 60
    lo, up = helpers.random_ldu_addressing(n, 80, 7)
    for asym in (False, True):
        diag, upper, lower, x, b = helpers.spd_coefficients(n, lo, up, seed=8, asym=asym)
        A = helpers.dense_from_ldu(n, lo, up, diag, upper, lower)
        y = np.empty(n)
        oracle.afo_ldu_amul(n, len(lo), P(lo), P(up), P(diag), P(upper), P(lower), P(x), P(y))
        assert np.allclose(y, A @ x, rtol=1e-13, atol=1e-13)
        # DIC / DILU = (D' + L) D'^-1 (D' + U) with D' from the incomplete factorisation restricted to the diagonal
        kind = abi.PRECOND_DILU if asym else abi.PRECOND_DIC
        w, rD = np.empty(n), np.empty(n)
        oracle.afo_ldu_precondition(n, len(lo), P(lo), P(up), P(diag), P(upper), P(lower), kind, P(b), P(w), P(rD))
        Lm, Um = np.tril(A, -1), np.triu(A, 1)
        Dp = np.diag(1.0 / rD)
        M = (Dp + Lm) @ np.diag(rD) @ (Dp + Um)
        assert np.allclose(M @ w, b, rtol=1e-11, atol=1e-12)
        # and D' satisfies its defining recurrence
        d = diag.copy()
        lw = upper if lower is None else lower
        for f in range(len(lo)):
            d[up[f]] -= upper[f] * lw[f] / d[lo[f]]
        assert np.allclose(1.0 / d, rD, rtol=1e-13)


@pytest.mark.parametrize("solver,precond,asym", [(abi.SOLVER_PCG, abi.PRECOND_DIC, False), (abi.SOLVER_PCG, abi.PRECOND_DIAGONAL, False),
                                                 (abi.SOLVER_PBICGSTAB, abi.PRECOND_DILU, True), (abi.SOLVER_PBICGSTAB, abi.PRECOND_NONE, True)])
def test_krylov_solvers_against_dense(oracle, solver, precond, asym):
    n = 120
    lo, up = helpers.random_ldu_addressing(n, 200, 9)
    diag, upper, lower, x0, b = helpers.spd_coefficients(n, lo, up, seed=10, asym=asym)
    A = helpers.dense_from_ldu(n, lo, up, diag, upper, lower)
    x = x0.copy()
    perf = abi.SolverPerf()
    log, nlog = np.empty(512), C.c_int32()
    oracle.afo_ldu_solve(n, len(lo), P(lo), P(up), P(diag), P(upper), P(lower), P(b), P(x), solver, precond, 1e-12, 0.0, 0, 400,
                         C.byref(perf), P(log), 512, C.byref(nlog))
    assert perf.converged and np.allclose(x, np.linalg.solve(A, b), rtol=1e-8, atol=1e-10)
    # residual normalisation: [UPSTREAM] lduMatrix::solver::normFactor
    xbar = x0.mean()
    sumA = A.sum(axis=1)
    nf = np.abs(A @ x0 - xbar * sumA).sum() + np.abs(b - xbar * sumA).sum() + 1e-20
    assert perf.initialResidual == pytest.approx(np.abs(b - A @ x0).sum() / nf, rel=1e-12)
    assert log[0] == perf.initialResidual and log[nlog.value - 1] == perf.finalResidual


def test_max_iter_is_post_increment(oracle):
    """do {...} while (nIterations()++ < maxIter_ && !converged): an unconverged solve reports maxIter + 1 iterations
    (the familiar 'No Iterations 1001' of OpenFOAM logs); minIter forces at least that many."""
    n = 200
    lo, up = helpers.random_ldu_addressing(n, 300, 11)
    diag, upper, lower, x0, b = helpers.spd_coefficients(n, lo, up, seed=12)
    diag = -np.bincount(lo, upper, n) - np.bincount(up, upper, n) + 1e-6
    for maxIter, want in [(3, 4), (20, 21)]:
        x = x0.copy()
        perf = abi.SolverPerf()
        oracle.afo_ldu_solve(n, len(lo), P(lo), P(up), P(diag), P(upper), None, P(b), P(x), abi.SOLVER_PCG, abi.PRECOND_DIC, 1e-15, 0.0,
                             1, maxIter, C.byref(perf), None, 0, None)
        assert perf.nIterations == want and not perf.converged
    x = b / diag
    perf = abi.SolverPerf()
    oracle.afo_ldu_solve(n, 0, P(lo[:0].copy()), P(up[:0].copy()), P(diag), P(upper[:0].copy()), None, P(b), P(x), abi.SOLVER_PCG,
                         abi.PRECOND_DIAGONAL, 1e-6, 0.0, 3, 20, C.byref(perf), None, 0, None)
    assert perf.nIterations == 3 or perf.singular              # minIter honoured unless the residual is exactly zero


# ---------------------------------------------------------------- the case as a whole
def test_material_anchors(oracle):
    c = ol.OracleCase(cases.amb2018_02_b(n=(6, 5, 4)), lib=oracle)
    assert oracle.afo_case_Tliq(c.h) == 1620.0 and oracle.afo_case_Tsol(c.h) == 1410.0
    assert np.all(c.field(abi.FIELD_ALPHA_SOLID) == 1.0)
    kap, cp = c.field(abi.FIELD_KAPPA), c.field(abi.FIELD_CP)
    assert kap[0] == pytest.approx(8.275 + 0.01472 * 300.0, rel=1e-15) and cp[0] == 579.28
    hot = np.full(c.n, 1700.0)
    c.set_field(abi.FIELD_T, hot)
    c.set_field(abi.FIELD_ALPHA_SOLID, np.zeros(c.n))
    c.step()
    # liquid conductivity is evaluated at clamp(T, Tliq, 2 Tliq); kappa_solid(1410) = 29.0302, kappa_liq(1620) = 28.77266
    assert 8.275 + 0.01472 * 1410.0 == pytest.approx(29.0302)
    assert 4.889 + 0.014743 * 1620.0 == pytest.approx(28.77266)
    c.close()


def test_power_is_conserved_every_step(oracle):
    """Sum(V*qDot) = eta*P while the Riemann sum stays within 5 % of V0 (heatSourceModel.C:359-367)."""
    c = ol.OracleCase(cases.amb2018_02_b(n=(75, 13, 8)), lib=oracle)
    for _ in range(25):
        r = c.step()
        assert r.totalPower == pytest.approx(0.33 * 179.2, rel=1e-12)
    c.close()


def test_explicit_step_limit(oracle):
    """alpha_th = kappa/(rho Cp) ~ 6.6e-6 m^2/s: Di = 1 at dt ~ 1.0e-5 s on 20 um cells; steps below it are explicit."""
    c = ol.OracleCase(cases.amb2018_02_b(), lib=oracle)
    reps = [c.step() for _ in range(45)]
    assert all(r.explicitStep == 1 for r in reps)
    assert max(r.DiNum for r in reps) <= 1.0 + 0.21            # growth capped at 1.2x per step
    assert 5e-6 < reps[-1].deltaT < 1.05e-5
    c.close()


def test_conduction_only_keeps_energy(oracle):
    """Lf = 0, constant properties, adiabatic walls: rho*Cp*V*sum(T - T0) == deposited energy (implicit PCG/DIC)."""
    case = cases.rosenthal(n=(40, 16, 8), tolerance=1e-14, maxIter=200)
    case["controls"]["deltaT"] = 2e-6
    c = ol.OracleCase(case, lib=oracle)
    V = (20e-6) ** 3
    E = 0.0
    for _ in range(20):
        r = c.step()
        E += r.totalPower * r.deltaT
        assert r.explicitStep == 0
    T = c.field(abi.FIELD_T)
    stored = 7569.92 * 600.0 * V * np.sum(T - 300.0)
    assert stored == pytest.approx(E, rel=1e-9)
    c.close()


def test_serial_vs_slab_decomposed(oracle):
    """The z-slab decomposition (block-local DIC, processor interfaces) reproduces the serial fields."""
    kw = dict(explicitSolve=0, maxDi=10.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC)
    a = ol.OracleCase(cases.amb2018_02_b(n=(40, 10, 9), **kw), ranks=1, lib=oracle)
    b = ol.OracleCase(cases.amb2018_02_b(n=(40, 10, 9), **kw), ranks=3, lib=oracle)
    for _ in range(40):
        ra, rb = a.step(), b.step()
        assert ra.deltaT == pytest.approx(rb.deltaT, rel=1e-10)
    assert helpers.rel_linf(b.field(abi.FIELD_T), a.field(abi.FIELD_T)) < 1e-9
    assert np.max(np.abs(b.field(abi.FIELD_ALPHA_SOLID) - a.field(abi.FIELD_ALPHA_SOLID))) < 1e-9
    a.close(); b.close()


def test_melt_pool_forms_and_solidifies(oracle):
    """After the beam has passed, cells that melted are solid again and alpha.solid is back in [0, 1] everywhere."""
    c = ol.OracleCase(cases.amb2018_02_b(n=(75, 13, 8)), lib=oracle)
    minA = 1.0
    for _ in range(150):
        c.step()
        minA = min(minA, c.field(abi.FIELD_ALPHA_SOLID).min())
    a = c.field(abi.FIELD_ALPHA_SOLID)
    assert minA < 0.5 and a.min() >= 0.0 and a.max() <= 1.0
    c.close()


def test_rosenthal_analytic(oracle):
    """An oracle-independent pin of the conduction path: the moving-source field against Rosenthal's solution."""
    import rosenthal
    case = rosenthal.case()
    c = ol.OracleCase(case, ranks=5, lib=oracle)
    while c.running():
        r = c.step()
    rosenthal.check(rosenthal.compare(c.field(abi.FIELD_T), r.time, case["mesh"]))
    c.close()


# ---------------------------------------------------------------- meltPoolDimensions function object
def test_melt_pool_dimensions_ellipsoid_and_rotation(oracle):
    """meltPoolDimensions.C:123-294: an ellipsoidal Gaussian hot spot inside the block.  The isotherm is the ellipsoid
    with semi-axes s_q*sqrt(ln(dT/(iso-T0))); linear interpolation between cell centres locates it to O(h^2).  Rotating
    the frame by 90 degrees swaps the x and y spans."""
    n = (64, 48, 40)
    case = cases.amb2018_02_b(n=n)
    m = case["mesh"]
    h = [(m["max"][q] - m["min"][q]) / n[q] for q in range(3)]
    o = ol.OracleCase(case, lib=oracle)
    x, y, z = [m["min"][q] + (np.arange(n[q]) + 0.5) * h[q] for q in range(3)]
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    ctr = [m["min"][q] + 0.5 * (m["max"][q] - m["min"][q]) for q in range(3)]
    s = [0.25 * (m["max"][0] - m["min"][0]), 0.2 * (m["max"][1] - m["min"][1]), 0.15 * (m["max"][2] - m["min"][2])]
    T = 300.0 + 2000.0 * np.exp(-((X - ctr[0]) / s[0]) ** 2 - ((Y - ctr[1]) / s[1]) ** 2 - ((Z - ctr[2]) / s[2]) ** 2)
    o.set_field(abi.FIELD_T, T.ravel())
    iso = 1000.0
    want = [2.0 * sq * np.sqrt(np.log(2000.0 / (iso - 300.0))) for sq in s]
    d0 = o.melt_pool_dimensions([iso, 5000.0], 0.0)
    for q in range(3):
        assert d0[0, q] == pytest.approx(want[q], abs=0.6 * h[q])
    assert np.all(d0[1] == 0.0)                                  # an isovalue nothing reaches: max(span, 0) = 0
    d90 = o.melt_pool_dimensions([iso], 90.0)
    assert d90[0, 0] == pytest.approx(d0[0, 1], rel=1e-12) and d90[0, 1] == pytest.approx(d0[0, 0], rel=1e-12)
    assert d90[0, 2] == d0[0, 2]
    o.close()


def test_melt_pool_dimensions_boundary_faces(oracle):
    """Physical boundary faces contribute their face centre when the cell (or patch) value reaches the isovalue
    (meltPoolDimensions.C:226-262): for T linear in x every boundary cell beyond the crossing plane counts, so the box runs
    from the crossing plane to xmax and over the whole y and z extent of the block."""
    n = (40, 10, 8)
    case = cases.amb2018_02_b(n=n)
    m = case["mesh"]
    o = ol.OracleCase(case, lib=oracle)
    hx = (m["max"][0] - m["min"][0]) / n[0]
    x = m["min"][0] + (np.arange(n[0]) + 0.5) * hx
    T = np.broadcast_to(300.0 + 1.0e6 * (x - m["min"][0]), (n[2], n[1], n[0])).copy()
    o.set_field(abi.FIELD_T, T.ravel())
    xstar = m["min"][0] + 0.3712 * (m["max"][0] - m["min"][0])
    iso = 300.0 + 1.0e6 * (xstar - m["min"][0])
    d = o.melt_pool_dimensions([iso], 0.0)[0]
    assert d[0] == pytest.approx(m["min"][0] + (m["max"][0] - m["min"][0]) - xstar, rel=1e-9)
    assert d[1] == pytest.approx((m["max"][1] - m["min"][1]), rel=1e-12) and d[2] == pytest.approx((m["max"][2] - m["min"][2]), rel=1e-12)
    o.close()


# ---------------------------------------------------------------- solidificationData function object
def test_solidification_events_against_numpy(oracle):
    """solidificationData.C:140-187 re-derived from field snapshots in numpy: a cell is recorded in the step where it cools
    through the isovalue; G is the Gauss-linear gradient (central differences inside the block), and the cooling rate is the
    one stored by the PREVIOUS execute, (T^{n-1} - T^{n-2})/dt^{n-1} -- not the current step's."""
    n = (75, 13, 20)
    case = cases.amb2018_02_b(n=n)
    m = case["mesh"]
    h = [(m["max"][q] - m["min"][q]) / n[q] for q in range(3)]
    o = ol.OracleCase(case, lib=oracle)
    iso = 1620.0
    o.solidification_enable([-1, -1, -1], [1, 1, 1], iso)
    shape = (n[2], n[1], n[0])
    snaps = [o.field(abi.FIELD_T).reshape(shape).copy()]
    want = []
    Rprev = np.zeros(shape)
    for step in range(260):
        rep = o.step()
        T = o.field(abi.FIELD_T).reshape(shape).copy()
        T0 = snaps[-1]
        ev = np.argwhere((T0 > iso) & (T <= iso))                    # C order = cell order (k, j, i)
        for k, j, i in ev:
            if not (0 < i < n[0] - 1 and 0 < j < n[1] - 1 and 0 < k < n[2] - 1):
                continue                                              # boundary cells are compared on the GPU side only
            g = np.array([(T[k, j, i + 1] - T[k, j, i - 1]) / (2 * h[0]), (T[k, j + 1, i] - T[k, j - 1, i]) / (2 * h[1]),
                          (T[k + 1, j, i] - T[k - 1, j, i]) / (2 * h[2])])
            G = max(np.linalg.norm(g), 1e-15)
            R = abs(Rprev[k, j, i])
            want.append([m["min"][0] + (i + 0.5) * h[0], m["min"][1] + (j + 0.5) * h[1], m["min"][2] + (k + 0.5) * h[2],
                         rep.time, R, G, R / G])
        Rprev = (T - T0) / rep.deltaT
        snaps = [T]
    got = o.solidification_events()
    assert len(got) > 50, len(got)
    interior = np.array([e for e in got if m["min"][0] + h[0] < e[0] < m["max"][0] - h[0] and m["min"][1] + h[1] < e[1] < m["max"][1] - h[1]
                         and m["min"][2] + h[2] < e[2] < m["max"][2] - h[2]])
    want = np.array(want)
    assert interior.shape == want.shape and len(want) > 20
    assert np.allclose(interior[:, :4], want[:, :4], rtol=1e-12, atol=0)
    assert np.allclose(interior[:, 4:], want[:, 4:], rtol=1e-9)
    assert np.all(got[:, 4] > 0) and np.all(np.diff(got[:, 3]) >= 0)    # cooling rates known by then; append order is by time
    o.close()


def test_solidification_box_and_decomposition(oracle):
    """setOverlapCells (solidificationData.C:99-131): only cells whose bounding box touches the box record events; the union
    of the per-rank event lists of a decomposed run is the serial list."""
    case = cases.amb2018_02_b(n=(75, 13, 8))
    box = ([0.0004, -1, -1], [0.0012, 1, 1])
    a, b = ol.OracleCase(case, lib=oracle), ol.OracleCase(case, ranks=2, lib=oracle)
    for o in (a, b):
        o.solidification_enable(box[0], box[1], 1620.0)
    for _ in range(220):
        a.step(); b.step()
    ea, eb = a.solidification_events(), b.solidification_events()
    dx = 0.003 / 75
    assert len(ea) > 10 and ea[:, 0].min() >= 0.0004 - dx and ea[:, 0].max() <= 0.0012 + dx
    assert len(ea) == len(eb)
    key = lambda e: np.lexsort((e[:, 0], e[:, 1], e[:, 2], e[:, 3]))
    assert np.allclose(ea[key(ea)], eb[key(eb)], rtol=1e-9)
    a.close(); b.close()


def test_ldu_decomposition_helper_reassembles_the_global_product():
    """partition.decompose_ldu (local matrix + processor interfaces, what the decomposed solver seam receives):
    the ranks' local products minus interfaceBouCoeffs * neighbour values add up to the global A x."""
    import helpers
    from additivefoam_b200 import partition
    n = 500
    lo, up = helpers.random_ldu_addressing(n, 1200, 5)
    for asym in (False, True):
        diag, upper, lower, x, _ = helpers.spd_coefficients(n, lo, up, seed=3, asym=asym)
        want = helpers.dense_from_ldu(n, lo, up, diag, upper, lower) @ x
        for R in (2, 3, 5):
            parts = [partition.decompose_ldu(n, lo, up, diag, upper, lower, r, R) for r in range(R)]
            got = np.zeros(n)
            for r, d in enumerate(parts):
                assert np.all(d["lower"] < d["upper"]) and np.all(np.diff(d["lower"]) >= 0)      # still upper-triangular order
                y = helpers.dense_from_ldu(d["n"], d["lower"], d["upper"], d["diag"], d["upper_coeffs"], d["lower_coeffs"]) @ x[d["c0"]:d["c1"]]
                for (other, fc), bc in zip(d["interfaces"], d["bou_coeffs"]):
                    o = parts[other]
                    nfc = next(f for rr, f in o["interfaces"] if rr == r)
                    assert len(nfc) == len(fc)
                    np.subtract.at(y, fc, bc * x[o["c0"] + nfc])
                got[d["c0"]:d["c1"]] = y
            assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))
