"""The beam shapes (weight, V0) and the table interpolation against values computed by the REFERENCE's own code
(tests/golden/ref_models.json, made by tests/golden/make_ref_models_golden.py from oracle/_ref/refModels = the reference's
superGaussian.C, modifiedSuperGaussian.C, projectedGaussian.C and interpolateXY.C compiled where they lie).  Checked: the
oracle's restatement and the library's host functions (what the device kernel calls, b200_beam_shape_*).  Host code only."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from additivefoam_b200 import cases, lib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_models.json")))
PATH = [[1, 0, 0, 0, 100.0, 0.0], [0, 1e-3, 0, 0, 100.0, 0.8]]


def beam_of(s):
    if s["kind"] == "super":
        return cases.super_gaussian_beam(PATH, dims=s["dims"], k=s["k"])
    if s["kind"] == "modified":
        return cases.modified_super_gaussian_beam(PATH, dims=s["dims"], k=s["k"], m=s["m"])
    return cases.projected_gaussian_beam(PATH, dims=s["dims"], A=s["A"], B=s["B"])


def ulps(a, b):
    if a == b:
        return 0.0
    return abs(a - b) / np.spacing(max(abs(a), abs(b)))


@pytest.mark.parametrize("i", range(len(GOLD["shapes"])))
def test_shapes_match_the_reference_code(oracle, i):
    s = GOLD["shapes"][i]
    b = beam_of(s)
    ob = cases.make_beam(b)
    dims = np.array([s["dims"][0], s["dims"][1], s["depth"]])
    want_v0 = float.fromhex(s["V0"])
    # V0: products of the same factors; a different association of the product may cost an ulp or two
    assert ulps(oracle.afo_shape_V0(C.byref(ob), ol.P(dims)), want_v0) <= 2
    assert ulps(lib.beam_shape_V0(b, dims)[0], want_v0) <= 2
    worst_o = worst_l = 0.0
    for p, w in zip(s["points"], s["weights"]):
        want = float.fromhex(w)
        d = np.array(p)
        got_o = oracle.afo_shape_weight(C.byref(ob), ol.P(dims), ol.P(d))
        got_l = lib.beam_shape_weight(b, dims, d)
        if want == 0.0:
            assert got_o == 0.0 and got_l == 0.0
        else:
            worst_o, worst_l = max(worst_o, abs(got_o - want) / want), max(worst_l, abs(got_l - want) / want)
    # exp(-x): an ulp of difference in x (operand order of the quotients) becomes x ulps of the result
    assert worst_o < 1e-13 and worst_l < 1e-13, (worst_o, worst_l)


@pytest.mark.parametrize("i", range(len(GOLD["tables"])))
def test_interpolateXY_matches_the_reference_code(oracle, i):
    t = GOLD["tables"][i]
    X, Y = np.array(t["x"]), np.array(t["y"])
    for q, v in zip(t["queries"], t["values"]):
        assert oracle.afo_interpolateXY(q, len(X), ol.P(X), ol.P(Y)) == float.fromhex(v), q       # same expression: bit for bit


@pytest.mark.skipif(not os.path.exists("/root/reference/applications/solvers/additiveFoam/interpolateXY/interpolateXY.C"),
                    reason="the reference tree is only present in the build container")
def test_golden_is_what_the_reference_code_computes(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_ref_models_golden as mk
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    s = mk.shape_requests()[3]
    lines = [f"shape {s['kind']} {s['k']!r} {s['m']!r} {s['A']!r} {s['B']!r} {s['dims'][0]!r} {s['dims'][1]!r} {s['dims'][2]!r} {s['depth']!r} {len(s['points'])}"]
    lines += [f"{p[0]!r} {p[1]!r} {p[2]!r}" for p in s["points"]]
    ans = mk.ask(lines)
    assert ans[0][3:] == GOLD["shapes"][3]["V0"] and ans[1:] == GOLD["shapes"][3]["weights"]


# ---------------------------------------------------------------------------------------------------------------------
# The per-cell fragments of the time loop -- updateProperties.H (kappa, Cp, powder reset) and thermo/thermoSource.H (dFdT, T0) --
# as the REFERENCE's own text computes them (included verbatim into oracle/_ref/refModels), against the oracle and the CUDA
# kernels (k_props, k_corrector) through the case API: upload T / alpha.solid / alpha.powder, take one step with a single
# corrector, read kappa, Cp, alpha.powder, dFdT, T0 back.
def cell_case(rec, tol, explicit):
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_ref_models_golden as mk
    from additivefoam_b200 import abi
    n = len(rec["T"])
    case = cases.amb2018_02_b(n=(n, 1, 1), explicitSolve=1 if explicit else 0, maxDi=1e30, solver=abi.SOLVER_PCG,
                              preconditioner=abi.PRECOND_DIAGONAL, nThermoCorrectors=1, thermoTolerance=tol, deltaT=1e-9, adjustTimeStep=0)
    case["beams"] = []
    m = mk.MATERIALS[rec["material"]]
    x, y = mk.TABLES[rec["material"]]
    case["material"].update(kappa=m["kappa"], Cp=m["Cp"], thermoT=list(x), thermoAlpha=list(y))
    return case


def check_cells(make, rec, tol):
    from additivefoam_b200 import abi
    for explicit in (False, True):
        c = make(cell_case(rec, tol, explicit))
        c.set_field(abi.FIELD_T, np.array(rec["T"]))
        c.set_field(abi.FIELD_ALPHA_SOLID, np.array(rec["alpha1"]))
        c.set_field(abi.FIELD_ALPHA_POWDER, np.array(rec["alpha3"]))
        rep = c.step()
        assert rep.nThermoCorr == 1 and rep.explicitStep == int(explicit)
        want = lambda key: np.array([float.fromhex(v) for v in rec[key]])      # noqa: E731
        # kappa, Cp: sums of products in the fragment's order, no transcendental function: bit for bit
        assert np.array_equal(c.field(abi.FIELD_KAPPA), want("kappa"))
        assert np.array_equal(c.field(abi.FIELD_CP), want("Cp"))
        assert np.array_equal(c.field(abi.FIELD_ALPHA_POWDER), want("alpha3_after"))
        if not explicit:        # the explicit pass fuses the solve and the alpha update; dFdT and T0 are not stored there
            # a temperature exactly on an interior row of the table gives lo == hi and dFdT = 0/0 in the reference's text;
            # the restatements reproduce that NaN rather than repairing it
            assert np.array_equal(c.field(abi.FIELD_DFDT), want(f"dFdT_{tol:g}"), equal_nan=True)
            assert np.array_equal(c.field(abi.FIELD_T0), want(f"T0_{tol:g}"), equal_nan=True)
        c.close()


@pytest.mark.parametrize("i", range(len(GOLD["cells"])))
@pytest.mark.parametrize("tol", [1e-8, 1e-6])
def test_oracle_cell_fragments_match_the_reference_text(oracle, i, tol):
    rec = GOLD["cells"][i]
    assert float.fromhex(rec["Tliq"]) == max(rec["T"][:2]) and float.fromhex(rec["Tsol"]) == min(rec["T"][:2])
    check_cells(lambda case: ol.OracleCase(case, lib=oracle), rec, tol)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(GOLD["cells"])))
@pytest.mark.parametrize("tol", [1e-8, 1e-6])
def test_gpu_cell_fragments_match_the_reference_text(ctx, i, tol):
    check_cells(lambda case: lib.Case(ctx, case), GOLD["cells"][i], tol)


@pytest.mark.parametrize("i", range(len(GOLD["kelly"])))
def test_kelly_absorption_matches_the_reference_code(oracle, i):
    """absorptionModels/Kelly/KellyAbsorption.C (compiled into oracle/_ref/refKelly): cone and cylinder, the discontinuity at
    aspect ratio 1, against the oracle's restatement and the library's host function."""
    from additivefoam_b200 import abi
    k = GOLD["kelly"][i]
    b = cases.modified_super_gaussian_beam(PATH, eta0=k["eta0"], etaMin=k["etaMin"])
    b["kellyGeometry"] = abi.KELLY_CONE if k["geometry"] == "cone" else abi.KELLY_CYLINDER
    ob = cases.make_beam(b)
    for a, e in zip(k["ratios"], k["eta"]):
        want = float.fromhex(e)
        assert oracle.afo_absorption_eta(C.byref(ob), a) == want, a          # same expression: bit for bit
        assert lib.beam_absorption_eta(b, a) == want, a
