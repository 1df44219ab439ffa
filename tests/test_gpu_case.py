"""GPU parity of the resident thermal step (b200_case_*) against the CPU oracle, step by step."""
import numpy as np
import pytest

import helpers
import oracle_lib as ol
from additivefoam_b200 import abi, cases, lib

pytestmark = pytest.mark.gpu

TOL = 1e-9     # BASELINE.json north_star: T and alpha.solid within 1e-9 relative Linf per time step


def run_pair(ctx, oracle, case, steps, check_every=1):
    g = lib.Case(ctx, case)
    o = ol.OracleCase(case, lib=oracle)
    worst_T = worst_a = 0.0
    for s in range(steps):
        rg, ro = g.step(), o.step()
        assert rg.explicitStep == ro.explicitStep, f"step {s}"
        assert rg.deltaT == pytest.approx(ro.deltaT, rel=1e-11), f"step {s}"
        assert rg.nThermoCorr == ro.nThermoCorr, f"step {s}"
        assert abs(rg.lastSolveIterations - ro.lastSolveIterations) <= 1, f"step {s}"
        assert rg.totalPower == pytest.approx(ro.totalPower, rel=1e-9, abs=1e-12)
        if s % check_every == 0 or s == steps - 1:
            Tg, To = g.field(abi.FIELD_T), o.field(abi.FIELD_T)
            ag, ao = g.field(abi.FIELD_ALPHA_SOLID), o.field(abi.FIELD_ALPHA_SOLID)
            worst_T = max(worst_T, helpers.rel_linf(Tg, To))
            worst_a = max(worst_a, float(np.max(np.abs(ag - ao))))
            assert worst_T < TOL and worst_a < TOL, f"step {s}: T {worst_T:.3e} alpha {worst_a:.3e}"
    out = (g.field(abi.FIELD_T), g.field(abi.FIELD_ALPHA_SOLID), rg, ro)
    g.close(); o.close()
    return out


def test_amb2018_as_shipped_explicit(ctx, oracle):
    """configs[0]: tutorials/AMB2018-02-B as shipped (explicit steps, solidification correctors)."""
    T, a, rg, ro = run_pair(ctx, oracle, cases.amb2018_02_b(), 80, check_every=4)
    assert rg.explicitStep == 1
    assert T.max() > 1620.0 and a.min() < 0.5      # a melt pool exists


@pytest.mark.parametrize("solver,precond", [(abi.SOLVER_PCG, abi.PRECOND_DIC), (abi.SOLVER_PBICGSTAB, abi.PRECOND_DILU),
                                            (abi.SOLVER_PCG, abi.PRECOND_DIAGONAL)])
def test_implicit_small(ctx, oracle, solver, precond):
    case = cases.amb2018_02_b(n=(60, 12, 8), explicitSolve=0, maxDi=10.0, solver=solver, preconditioner=precond)
    T, a, rg, ro = run_pair(ctx, oracle, case, 50, check_every=5)
    assert rg.explicitStep == 0 and T.max() > 1620.0


def test_fixed_value_powder_tmax(ctx, oracle):
    """multiLayerPBF ingredients: fixedValue walls, powder layer, Tmax limiter, transient modifiedSuperGaussian + Kelly."""
    case = cases.multi_layer_pbf(n=(60, 24, 12), nHatch=2, endTime=0.002)
    case["controls"]["deltaT"] = 1e-6
    run_pair(ctx, oracle, case, 40, check_every=5)


def test_multibeam_small(ctx, oracle):
    case = cases.multi_beam(n=(80, 30, 10), endTime=0.002)
    case["controls"]["deltaT"] = 1e-6
    run_pair(ctx, oracle, case, 30, check_every=5)


def test_projected_gaussian_subcycled(ctx, oracle):
    case = cases.amb2018_02_b(n=(60, 12, 8), explicitSolve=0, maxDi=10.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC)
    path = case["beams"][0]["path"]
    b = cases.projected_gaussian_beam(path)
    b["deltaT"] = 2e-6          # beam sub-cycling (movingHeatSourceModel.C:128-147)
    case["beams"] = [b]
    case["controls"]["deltaT"] = 1e-6
    run_pair(ctx, oracle, case, 30, check_every=5)


def test_rosenthal_conduction_only(ctx, oracle):
    case = cases.rosenthal(n=(60, 20, 10))
    case["controls"]["deltaT"] = 1e-6
    T, a, rg, ro = run_pair(ctx, oracle, case, 30, check_every=5)
    assert rg.nThermoCorr >= 2     # beam on => at least two passes (TEqn.H:58)


# ---------------------------------------------------------------- meltPoolDimensions on the resident field
@pytest.mark.parametrize("which", ["amb_explicit", "pbf_fixed_value"])
def test_melt_pool_dimensions_parity(ctx, oracle, which):
    """functionObjects/meltPoolDimensions/meltPoolDimensions.C:123-294 on the device-resident T against the oracle.  The
    crossing points are computed with the same operations (no FMA contraction) and min/max do not depend on order, so on
    identical T the spans agree to the last bit; on each side's own T after the same steps they inherit the 1e-9 of T."""
    if which == "amb_explicit":
        case, steps = cases.amb2018_02_b(n=(75, 13, 8)), 80
    else:
        case, steps = cases.multi_layer_pbf(n=(60, 24, 12), nHatch=2, endTime=0.002), 40
        case["controls"]["deltaT"] = 1e-6
    g = lib.Case(ctx, case)
    o = ol.OracleCase(case, lib=oracle)
    for _ in range(steps):
        g.step(); o.step()
    iso = [1620.0, 1410.0, 800.0, 1.0e5]
    for angle in (0.0, 67.0):
        dg, do = g.melt_pool_dimensions(iso, angle), o.melt_pool_dimensions(iso, angle)
        assert do[0].min() > 0.0 and np.all(do[3] == 0.0)
        assert np.max(np.abs(dg - do)) <= 1e-9 * np.max(do)
    # identical field on both sides -> identical bits
    T = o.field(abi.FIELD_T)
    g.set_field(abi.FIELD_T, T)
    assert np.array_equal(g.melt_pool_dimensions(iso, 67.0), o.melt_pool_dimensions(iso, 67.0))
    g.close(); o.close()


# ---------------------------------------------------------------- solidificationData on the device
@pytest.mark.parametrize("which", ["amb_explicit", "amb_implicit_box", "pbf_fixed_value"])
def test_solidification_events_parity(ctx, oracle, which):
    """functionObjects/solidificationData/solidificationData.C:140-187 executed on the device at the end of every step against
    the oracle: same events in the same order; positions bit-equal, times/rates/gradients to the accuracy T carries."""
    box = ([-1, -1, -1], [1, 1, 1])
    if which == "amb_explicit":
        case, steps = cases.amb2018_02_b(n=(75, 13, 16)), 200
    elif which == "amb_implicit_box":
        case, steps = cases.amb2018_02_b(n=(60, 12, 16), explicitSolve=0, maxDi=10.0), 90
        box = ([0.0002, -0.00005, -1], [0.0009, 1, 1])
    else:
        case, steps = cases.multi_layer_pbf(n=(60, 24, 12), nHatch=2, endTime=0.002), 220
        case["controls"]["deltaT"] = 1e-6
    g = lib.Case(ctx, case)
    o = ol.OracleCase(case, lib=oracle)
    g.solidification_enable(box[0], box[1], 1620.0)
    o.solidification_enable(box[0], box[1], 1620.0)
    for _ in range(steps):
        g.step(); o.step()
    eg, eo = g.solidification_events(), o.solidification_events()
    assert len(eo) > 10 and eg.shape == eo.shape
    assert np.array_equal(eg[:, :3], eo[:, :3])
    assert np.allclose(eg[:, 3], eo[:, 3], rtol=1e-11, atol=0)
    assert np.allclose(eg[:, 4:], eo[:, 4:], rtol=1e-6)
    assert helpers.rel_linf(g.field(abi.FIELD_T), o.field(abi.FIELD_T)) < TOL      # the hook does not disturb the step
    g.close(); o.close()


def test_solidification_overflow_is_reported(ctx):
    case = cases.amb2018_02_b(n=(75, 13, 16))
    g = lib.Case(ctx, case)
    g.solidification_enable([-1, -1, -1], [1, 1, 1], 1620.0, max_events=4)
    for _ in range(200):
        g.step()
    with pytest.raises(RuntimeError, match="exceed the capacity"):
        g.solidification_events()
    with pytest.raises(RuntimeError, match="already enabled"):
        g.solidification_enable([-1, -1, -1], [1, 1, 1], 1620.0)
    g.close()
