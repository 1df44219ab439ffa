"""Edge cases of the resident step through the C ABI, against the oracle: no beam, beam outside the block, degenerate
blocks (one cell thick in each direction, a single cell), a curved thermoPath table, field upload / restart, beam switch-off."""
import numpy as np
import pytest

import helpers
import oracle_lib as ol
from additivefoam_b200 import abi, cases, lib

pytestmark = pytest.mark.gpu
IMPLICIT = dict(explicitSolve=0, maxDi=10.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC)


def compare(ctx, oracle, case, steps, prepare=None):
    g, o = lib.Case(ctx, case), ol.OracleCase(case, lib=oracle)
    if prepare:
        prepare(g, o)
    for s in range(steps):
        rg, ro = g.step(), o.step()
        assert rg.explicitStep == ro.explicitStep and rg.nThermoCorr == ro.nThermoCorr, s
        assert rg.deltaT == pytest.approx(ro.deltaT, rel=1e-11)
        assert abs(rg.lastSolveIterations - ro.lastSolveIterations) <= 1
    Tg, To = g.field(abi.FIELD_T), o.field(abi.FIELD_T)
    assert helpers.rel_linf(Tg, To) < 1e-9
    assert np.max(np.abs(g.field(abi.FIELD_ALPHA_SOLID) - o.field(abi.FIELD_ALPHA_SOLID))) < 1e-9
    out = (Tg, rg)
    g.close(); o.close()
    return out


def test_no_beams_stays_uniform(ctx, oracle):
    case = cases.amb2018_02_b(n=(12, 6, 5), **IMPLICIT)
    case["beams"] = []
    T, r = compare(ctx, oracle, case, 6)
    assert r.nThermoCorr == 1 and r.totalPower == 0.0          # neither beam nor liquid: one corrector (solutionControls.H:25-28)
    assert np.allclose(T, 300.0, rtol=0, atol=1e-9)


def test_beam_outside_the_block(ctx, oracle):
    case = cases.amb2018_02_b(n=(12, 6, 5), **IMPLICIT)
    case["beams"][0]["path"] = [[1, 0.1, 0.1, 0.0, 0.0, 0.0], [0, 0.102, 0.1, 0.0, 179.2, 0.8]]
    T, r = compare(ctx, oracle, case, 5)
    assert r.totalPower == 0.0


@pytest.mark.parametrize("n", [(1, 1, 1), (9, 1, 1), (1, 7, 1), (1, 1, 6), (8, 5, 1), (3, 1, 4)])
@pytest.mark.parametrize("mode", ["implicit", "explicit"])
def test_degenerate_blocks(ctx, oracle, n, mode):
    kw = IMPLICIT if mode == "implicit" else dict()
    case = cases.amb2018_02_b(n=n, **kw)
    # keep the 20 um cell size so that the beam still covers the block
    case["mesh"]["min"] = [-10e-6 * n[0], -10e-6 * n[1], -20e-6 * n[2]]
    case["mesh"]["max"] = [10e-6 * n[0], 10e-6 * n[1], 0.0]
    case["beams"][0]["path"] = [[1, 0.0, 0.0, 0.0, 0.0, 0.0], [1, 0.0, 0.0, 0.0, 40.0, 1.0]]     # dwell at the centre
    compare(ctx, oracle, case, 12)


def test_curved_thermo_path(ctx, oracle):
    """A thermoPath with several rows, listed in descending temperature (the reference's scans tolerate it)."""
    case = cases.amb2018_02_b(n=(40, 10, 8), **IMPLICIT)
    case["material"]["thermoT"] = [1620.0, 1590.0, 1520.0, 1450.0, 1410.0]
    case["material"]["thermoAlpha"] = [0.0, 0.15, 0.55, 0.9, 1.0]
    T, r = compare(ctx, oracle, case, 45)
    assert T.max() > 1620.0


def test_upload_fields_and_continue(ctx, oracle):
    """b200_case_set_field: start from a hot, partly molten state."""
    case = cases.amb2018_02_b(n=(20, 8, 6), **IMPLICIT)
    rng = np.random.default_rng(3)

    def prepare(g, o):
        T0 = 300.0 + 1500.0 * rng.random(g.n_local)
        a0 = np.clip((1620.0 - T0) / 210.0, 0.0, 1.0)
        for c in (g, o):
            c.set_field(abi.FIELD_T, T0)
            c.set_field(abi.FIELD_ALPHA_SOLID, a0)
    compare(ctx, oracle, case, 12, prepare)


def test_beam_switches_off_and_pool_solidifies(ctx, oracle):
    case = cases.amb2018_02_b(n=(40, 10, 8), **IMPLICIT)
    case["beams"][0]["path"] = [[1, 0.0, 0.0, 0.0, 0.0, 0.0], [0, 0.0003, 0.0, 0.0, 179.2, 0.8]]      # 0.375 ms of beam
    case["controls"]["endTime"] = 0.002
    T, r = compare(ctx, oracle, case, 110)
    assert r.totalPower == 0.0 and r.time > 0.000375


@pytest.mark.parametrize("mode", ["implicit", "explicit"])
@pytest.mark.parametrize("walls", ["base", "sides", "all"])
def test_wall_temperature_differs_from_initial_temperature(ctx, oracle, mode, walls):
    """fixedValue walls at 290 K around a 300 K block: the first step already sees the wall value (a fixedValue patch is
    constructed from its own `value` entry).  All shipped cases have wall = initial temperature, which hid an oracle bug."""
    kw = IMPLICIT if mode == "implicit" else dict()
    case = cases.amb2018_02_b(n=(20, 8, 6), **kw)
    if walls in ("base", "all"):
        case["patches"][0] = cases._patch(abi.BC_FIXED_VALUE, [abi.ZMIN], value=290.0)
    if walls in ("sides", "all"):
        case["patches"][2] = cases._patch(abi.BC_FIXED_VALUE, [abi.XMIN, abi.XMAX, abi.YMIN, abi.YMAX], value=290.0)
    T, r = compare(ctx, oracle, case, 15)
    assert T.min() < 299.9        # the walls do cool the block from the first step on
