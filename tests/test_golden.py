"""Committed vectors of tests/golden/ (see tests/golden/make_golden.py for what they are and how they were made).

CPU (-m "not gpu"): the hand-derived known answers hold for the oracle, and the oracle still reproduces its frozen outputs.
GPU (-m gpu): the CUDA path, through the C ABI, against the same frozen outputs -- no oracle call involved."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402

import helpers  # noqa: E402
from additivefoam_b200 import abi, cases  # noqa: E402

GOLD = os.path.join(HERE, "golden")
TOL = 1e-9            # BASELINE.json north_star: fields within 1e-9 relative Linf, iteration counts within +-1


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def known():
    with open(os.path.join(GOLD, "known_answers.json")) as f:
        return json.load(f)


# ------------------------------------------------------------------------------------------------ CPU: pins of the oracle
def test_known_answers_file_is_what_the_formulas_give():
    assert known() == json.loads(json.dumps(mg.known_answers()))


def test_oracle_meets_known_answers(oracle):
    import ctypes as C
    import oracle_lib as ol
    ka = known()
    path = [[1, 0, 0, 0, 100.0, 0.0], [0, 1e-3, 0, 0, 100.0, 0.8]]
    d = np.array([85e-6, 85e-6, 30e-6])
    assert oracle.afo_shape_V0(C.byref(cases.make_beam(cases.super_gaussian_beam(path))), ol.P(d)) == pytest.approx(
        ka["superGaussian_V0_k2_85_85_30um"], rel=1e-14)
    d = np.array([40e-6, 40e-6, 30e-6])
    assert oracle.afo_shape_V0(C.byref(cases.make_beam(cases.modified_super_gaussian_beam(path))), ol.P(d)) == pytest.approx(
        ka["modifiedSuperGaussian_V0_k7.95_m2.72_40_40_30um"], rel=1e-13)
    kelly = cases.make_beam(cases.modified_super_gaussian_beam(path))
    assert oracle.afo_absorption_eta(C.byref(kelly), 0.75) == ka["kelly_cone_eta_aspect_0.75"]
    assert oracle.afo_absorption_eta(C.byref(kelly), 2.0) == pytest.approx(ka["kelly_cone_eta_aspect_2"], rel=1e-14)
    o = ol.OracleCase(cases.amb2018_02_b(), lib=oracle)
    assert o.n == ka["amb2018_cells"]
    assert oracle.afo_case_Tliq(o.h) == ka["Tliq"] and oracle.afo_case_Tsol(o.h) == ka["Tsol"]
    rep = o.step()
    assert rep.totalPower == pytest.approx(ka["amb2018_absorbed_power_W"], rel=1e-12)
    o.close()
    lo, up = helpers.block_addressing_np((150, 25, 15))
    assert len(lo) == ka["amb2018_internal_faces"]


@pytest.mark.parametrize("which", mg.LDU_PROBLEMS)
def test_oracle_reproduces_frozen_ldu(oracle, which):
    now, then = mg.oracle_ldu(oracle, which), gold("ldu_" + which)
    assert sorted(now) == sorted(then.files)
    for k in now:
        assert np.array_equal(now[k], then[k]), k      # same compiler flags, no contraction: bit for bit


@pytest.mark.parametrize("which", mg.CASES)
def test_oracle_reproduces_frozen_case(oracle, which):
    now, then = mg.oracle_case(oracle, which), gold("case_" + which)
    # exp/pow/tgamma come from libm: allow a few ulp should the C library differ from the one that froze the vectors
    np.testing.assert_allclose(now["reports"], then["reports"], rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(now["T"], then["T"], rtol=1e-11)
    np.testing.assert_allclose(now["alpha_solid"], then["alpha_solid"], rtol=0, atol=1e-11)


def test_oracle_reproduces_frozen_beams(oracle):
    now, then = mg.oracle_beams(oracle), gold("beams")
    for k in now:
        np.testing.assert_allclose(now[k], then[k], rtol=1e-12, atol=0)
    # power conservation of the Riemann sum (heatSourceModel.C:359-367): sum(V * qDot) = absorbed power
    V = np.prod((np.array(mg.BEAM_MESH["max"]) - np.array(mg.BEAM_MESH["min"])) / np.array(mg.BEAM_MESH["n"]))
    for name, *_ in mg.beam_inputs():
        assert float(then[name + "_qdot"].sum() * V) == pytest.approx(float(then[name + "_scalars"][2]), rel=1e-12)


# ------------------------------------------------------------------------------------------------ GPU against the frozen vectors
@pytest.mark.gpu
@pytest.mark.parametrize("which", mg.LDU_PROBLEMS)
def test_gpu_ldu_against_frozen(ctx, which):
    from additivefoam_b200 import lib
    dims, n, lo, up, diag, upper, lower, x, b = mg.ldu_problem(which)
    g = gold("ldu_" + which)
    s = lib.LduSolver(ctx, lo, up, n, block=dims)
    assert np.array_equal(s.amul(x, diag, upper, lower), g["amul"])
    for name, kind in (("dic", abi.PRECOND_DIC), ("dilu", abi.PRECOND_DILU), ("diagonal", abi.PRECOND_DIAGONAL)):
        if "precond_" + name in g.files:
            assert np.array_equal(s.precondition(kind, b, diag, upper, lower), g["precond_" + name]), name
    solver, precond = (abi.SOLVER_PCG, abi.PRECOND_DIC) if lower is None else (abi.SOLVER_PBICGSTAB, abi.PRECOND_DILU)
    psi, perf = s.solve(x, b, diag, upper, lower, solver=solver, preconditioner=precond, tolerance=1e-10, relTol=0.0,
                        minIter=1, maxIter=60)
    s.close()
    init, final, nit, conv, sing = g["solve_perf"]
    assert abs(perf.nIterations - int(nit)) <= 1 and perf.converged == int(conv) and perf.singular == int(sing)
    assert perf.initialResidual == pytest.approx(init, rel=1e-10)
    assert helpers.rel_linf(psi, g["solve_psi"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("which", mg.CASES)
def test_gpu_case_against_frozen(ctx, which):
    from additivefoam_b200 import lib
    case, steps = mg.case_inputs(which)
    g = gold("case_" + which)
    c = lib.Case(ctx, case)
    F = mg.REPORT_FIELDS
    for s in range(steps):
        r = c.step()
        ref = dict(zip(F, g["reports"][s]))
        assert r.explicitStep == int(ref["explicitStep"]) and r.nThermoCorr == int(ref["nThermoCorr"]), f"step {s}"
        assert abs(r.lastSolveIterations - int(ref["lastSolveIterations"])) <= 1, f"step {s}"
        assert r.deltaT == pytest.approx(ref["deltaT"], rel=1e-11), f"step {s}"
        assert r.time == pytest.approx(ref["time"], rel=1e-11), f"step {s}"
        assert r.totalPower == pytest.approx(ref["totalPower"], rel=1e-9, abs=1e-12), f"step {s}"
        assert r.DiNum == pytest.approx(ref["DiNum"], rel=1e-9), f"step {s}"
    T, a = c.field(abi.FIELD_T), c.field(abi.FIELD_ALPHA_SOLID)
    c.close()
    assert helpers.rel_linf(T, g["T"]) < TOL
    assert float(np.max(np.abs(a - g["alpha_solid"]))) < TOL


@pytest.mark.gpu
def test_gpu_beams_against_frozen(ctx):
    from additivefoam_b200 import lib
    g = gold("beams")
    for name, b, pos, dims in mg.beam_inputs():
        q, sumw, vol, absorbed = lib.beam_qdot(ctx, mg.BEAM_MESH, b, pos, dims, mg.BEAM_POWER)
        ref = g[name + "_qdot"]
        # exp/pow of the device differ from glibc's by <= 1 ulp
        assert float(np.max(np.abs(q - ref))) <= 1e-11 * float(np.max(np.abs(ref))), name
        np.testing.assert_allclose([sumw, vol, absorbed], g[name + "_scalars"], rtol=1e-11)
