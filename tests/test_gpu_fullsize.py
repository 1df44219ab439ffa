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
