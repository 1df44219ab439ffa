"""Size-independent properties at BASELINE.json's sizes (the oracle is too slow there):
energy conservation of the conduction-only 10 M-cell case (configs[1]), exact power conservation, bounded alpha.solid, and
the Krylov solve actually reducing the residual it reports, checked with the library's own Amul."""
import numpy as np
import pytest

from additivefoam_b200 import abi, cases, lib

pytestmark = pytest.mark.gpu


def test_rosenthal_10m_energy_and_power(ctx):
    """configs[1]: 500x200x100 conduction-only moving Gaussian, implicit PCG/DIC.  Lf = 0, constant kappa/Cp, adiabatic
    walls => rho*Cp*V*sum(T - T0) equals the deposited energy, and sum(V*qDot) = eta*P every step."""
    case = cases.rosenthal(n=(500, 200, 100), tolerance=1e-13, maxIter=200)
    case["controls"]["deltaT"] = 5e-6
    g = lib.Case(ctx, case)
    V = (20e-6) ** 3
    E = 0.0
    for _ in range(6):
        r = g.step()
        assert r.explicitStep == 0
        assert r.totalPower == pytest.approx(0.33 * 179.2, rel=1e-11)
        E += r.totalPower * r.deltaT
    T = g.field(abi.FIELD_T)
    stored = 7569.92 * 600.0 * V * float(np.sum(T - 300.0, dtype=np.float64))
    assert stored == pytest.approx(E, rel=1e-8)
    assert T.min() > 299.999 and np.isfinite(T).all()
    g.close()


def test_track_4m_alpha_bounds_and_monotone_time(ctx):
    """configs[2] family at 4 M cells: alpha.solid stays in [0, 1], time advances by the reported deltaT, the melt pool forms."""
    case = cases.synthetic_track((400, 100, 100), deltaT=2e-5)
    g = lib.Case(ctx, case)
    t = 0.0
    for _ in range(25):
        r = g.step()
        t += r.deltaT
        assert r.time == pytest.approx(t, rel=1e-12)
        assert 1 <= r.nThermoCorr <= 20
    a = g.field(abi.FIELD_ALPHA_SOLID)
    T = g.field(abi.FIELD_T)
    assert a.min() >= 0.0 and a.max() <= 1.0 and a.min() < 0.5 and T.max() > 1620.0
    # liquid fraction is consistent with temperature where the cell is fully one phase
    assert np.all(a[T > 1620.0 + 1e-6] <= 1e-6) and np.all(a[T < 1410.0 - 1e-6] >= 1.0 - 1e-6)
    g.close()


def test_solve_residual_property_3m(ctx, oracle):
    """3 M-cell block: the PCG/DIC solution reduces ||b - A x||_1 / normFactor to the reported final residual."""
    import oracle_lib as ol
    dims = (200, 150, 100)
    n, nf, lo, up = ol.block_addressing(oracle, dims)
    rng = np.random.default_rng(1234)
    upper = -rng.uniform(0.5, 1.5, nf)
    diag = rng.uniform(0.1, 1.0, n)
    np.subtract.at(diag, lo, upper)
    np.subtract.at(diag, up, upper)
    b = rng.uniform(-1, 1, n)
    x0 = np.zeros(n)
    s = lib.LduSolver(ctx, lo, up, n, block=dims)
    x, perf = s.solve(x0, b, diag, upper, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC, tolerance=1e-9, maxIter=200)
    assert perf.converged and perf.nIterations < 200
    Ax = s.amul(x, diag, upper)
    sumA = s.amul(np.ones(n), diag, upper)
    xbar = x0.mean()
    nfac = np.abs(s.amul(x0, diag, upper) - xbar * sumA).sum() + np.abs(b - xbar * sumA).sum() + 1e-20
    assert np.abs(b - Ax).sum() / nfac == pytest.approx(perf.finalResidual, rel=1e-6)
    assert perf.finalResidual < 1e-9
    s.close()


def test_rosenthal_analytic_gpu(ctx):
    """The same analytic pin through the C ABI (b200_case_*)."""
    import rosenthal
    case = rosenthal.case()
    g = lib.Case(ctx, case)
    while g.running():
        r = g.step()
    rosenthal.check(rosenthal.compare(g.field(abi.FIELD_T), r.time, case["mesh"]))
    g.close()


def test_multibeam_50m_power_and_bounds(ctx):
    """configs[3]: tutorials/multiBeam's two beam types on four simultaneous tracks, 1000x250x200 = 50 M cells, implicit PCG/DIC.
    sum(V*qDot) = sum_i eta_i(aspectRatio_i) P_i every step: constant absorption for the superGaussian pair, Kelly (cone) for
    the transient modifiedSuperGaussian pair, whose aspect ratio is the static one until a 1620 K isotherm exists and can only
    grow after that (heatSourceModel.C:249-257, :214-218)."""
    case = cases.multi_beam(n=(1000, 250, 200), deltaT=1e-5)
    g = lib.Case(ctx, case)
    kelly = case["beams"][1]
    eta_static = lib.beam_absorption_eta(kelly, 30e-6 / 40e-6)
    assert eta_static == 0.35                                      # aspect ratio <= 1: etaMin (KellyAbsorption.C:67-71)
    t = 0.0
    for s in range(15):
        r = g.step()
        t += r.deltaT
        assert r.time == pytest.approx(t, rel=1e-12)
        lo = 2 * 0.33 * 179.2 + 2 * eta_static * 179.2
        hi = 2 * 0.33 * 179.2 + 2 * 1.0 * 179.2
        if s == 0:
            assert r.totalPower == pytest.approx(lo, rel=1e-11)      # T = 300 K everywhere: static depth
        assert lo * (1 - 1e-11) <= r.totalPower <= hi
        assert 1 <= r.nThermoCorr <= 20
    a, T = g.field(abi.FIELD_ALPHA_SOLID), g.field(abi.FIELD_T)
    g.close()
    assert np.isfinite(T).all() and T.min() > 299.999 and T.max() > 1620.0
    assert a.min() >= 0.0 and a.max() <= 1.0 and a.min() < 0.5
    # all four tracks are heated; the block's y edges, 2.3 mm away, are untouched
    nx, ny, nz = 1000, 250, 200
    top = a.reshape(nz, ny, nx)[-1]
    Ttop = T.reshape(nz, ny, nx)[-1]
    for y in (50e-6, -50e-6, 150e-6, -150e-6):
        j = int((y + 0.5 * ny * 20e-6) / 20e-6)
        assert Ttop[j].max() > 1000.0                                # every track is being heated
    assert top[0].min() == 1.0 and top[-1].min() == 1.0


def test_multilayer_25m_slab_powder_and_walls(ctx):
    """configs[4] at the per-GPU share of the 8-GPU run (200 M / 8 = 25 M cells: 1000x500x50): powder layer above z = -40 um,
    fixedValue 300 K bottom and sides, transient Kelly modifiedSuperGaussian on a raster path.  Properties: absorbed power within
    the Kelly bounds and exactly etaMin*P on the cold first step, powder only ever removed, and removed wherever the cell was
    not fully solid at the start of a step (updateProperties.H:7-10), temperatures finite and not below the wall value (to the
    solver tolerance)."""
    case = cases.multi_layer_pbf(n=(1000, 500, 50), deltaT=1e-5)
    g = lib.Case(ctx, case)
    p0 = g.field(abi.FIELD_ALPHA_POWDER)
    assert set(np.unique(p0)) == {0.0, 1.0}
    for s in range(15):
        if s == 14:
            a4 = g.field(abi.FIELD_ALPHA_SOLID)                      # what the last step's updateProperties sees
        r = g.step()
        if s == 0:
            assert r.totalPower == pytest.approx(0.35 * 195.0, rel=1e-11)
        assert 0.35 * 195.0 * (1 - 1e-11) <= r.totalPower <= 195.0
    p, a, T = g.field(abi.FIELD_ALPHA_POWDER), g.field(abi.FIELD_ALPHA_SOLID), g.field(abi.FIELD_T)
    g.close()
    assert np.isfinite(T).all() and T.min() >= 300.0 - 1e-6 and T.max() > 1620.0
    assert np.all(p <= p0) and set(np.unique(p)) <= {0.0, 1.0}
    molten = a4 < 1.0
    assert molten.any() and not p[molten].any()
    assert (p0 - p).sum() >= (molten & (p0 == 1.0)).sum() > 0
    assert a.min() >= 0.0 and a.max() <= 1.0


def test_track_100m_full_size_properties(ctx):
    """configs[2] at its full size (1000x400x250 = 100 M cells, the workload bench.py times): absorbed power exact, the clock
    advances by the reported deltaT, every solve ends below its initial residual, alpha.solid stays in [0, 1] and is consistent
    with T outside the mushy zone, nothing but the track region is heated."""
    nx, ny, nz = 1000, 400, 250
    g = lib.Case(ctx, cases.synthetic_track((nx, ny, nz), deltaT=1e-5))
    t = 0.0
    for _ in range(10):
        r = g.step()
        t += r.deltaT
        assert r.time == pytest.approx(t, rel=1e-12)
        assert r.totalPower == pytest.approx(0.33 * 179.2, rel=1e-11)
        assert r.explicitStep == 0 and 1 <= r.nThermoCorr <= 20
        assert r.lastFinalResidual <= r.lastInitialResidual and np.isfinite(r.lastFinalResidual)
    a = np.empty(nx * ny * nz)
    T = np.empty(nx * ny * nz)
    g.field_into(abi.FIELD_ALPHA_SOLID, a)
    g.field_into(abi.FIELD_T, T)
    g.close()
    assert np.isfinite(T).all() and T.min() > 299.999 and T.max() > 1620.0
    assert a.min() >= 0.0 and a.max() <= 1.0 and a.min() < 0.5
    top = T.reshape(nz, ny, nx)[-1]
    assert top[ny // 2 - 1:ny // 2 + 1].max() == T.max()                          # hottest on the track line y = 0
    assert np.all(T.reshape(nz, ny, nx)[: nz // 2] < 300.001)                       # 2.5 mm below the surface: untouched after 100 us
    hot, cold = T > 1620.0 + 1e-6, T < 1410.0 - 1e-6
    assert np.all(a[hot] <= 1e-6) and np.all(a[cold] >= 1.0 - 1e-6)
