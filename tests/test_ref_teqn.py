"""The temperature equation and its corrector loop -- thermo/TEqn.H with thermo/thermoScheme.H and thermo/thermoSource.H -- against
values computed by the REFERENCE's own text (tests/golden/ref_teqn.json, made by tests/golden/make_ref_teqn_golden.py from
oracle/_ref/refTEqn = the three fragments included verbatim where they lie against an fvScalarMatrix stand-in).  The fragment was fed
the oracle's state at the start of every step of five small cases (as-shipped explicit; implicit PCG/DIC; implicit with the as-shipped
PBiCGStab/DILU; the Tmax penalty active; powder layer with fixedValue walls; three of them again with ddtSchemes backward and again with CrankNicolson (psi 0.9, 1, 0.5), where
thermoScheme.H scales the rDeltaT of the Sp terms by coefft and TEqn.H takes fvc::ddt(alpha1) from the scheme), so re-running the cases must give, step by step, the
reference text's T, alpha.solid, face values and value fractions of every patch, and the Krylov iterations of every corrector pass.
Checked: the oracle (CPU, bit for bit) and the GPU stepper through b200_case_step (1e-9; iterations +-1).

What the fixture found when it was first made: TEqn.H:44 calls T.correctBoundaryConditions() AFTER solve() has already evaluated the
patches, and the second evaluate recomputes the mixedTemperature value fraction from the new face values; the oracle and the GPU
stepper evaluated once.  Same internal field in that step, different face values / value fractions into the next step's matrix
(up to 6.5 % in the value fraction of the powder case).  Both now evaluate twice."""
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib as ol
from additivefoam_b200 import abi, lib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_ref_teqn_golden as mk  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "ref_teqn.json")))
NAMES = sorted(GOLD)
H = float.fromhex


def exact(a, d, what):
    a = np.asarray(a, dtype=np.float64)
    assert len(a) == d["n"], what
    assert [float(a[i]).hex() for i in d["idx"]] == d["val"], what
    s = 0.0
    for v in a:
        s += float(v)
    assert float(s).hex() == d["sum"], what


def close(a, d, what, rel, scale=None):
    a = np.asarray(a, dtype=np.float64)
    ref = np.array([H(v) for v in d["val"]])
    sc = scale if scale is not None else max(np.max(np.abs(ref)), 1e-300)
    assert np.max(np.abs(a[d["idx"]] - ref)) <= rel * sc, what
    assert abs(float(np.sum(a)) - H(d["sum"])) <= rel * max(abs(H(d["sum"])), sc) * 4, what


@pytest.mark.parametrize("name", NAMES)
def test_oracle_equation_matches_the_reference_text(oracle, name):
    case, steps = mk.configs()[name]
    o = ol.OracleCase(case, lib=oracle)
    for s, want in enumerate(GOLD[name]):
        rep = o.step()
        its = want["iters"]
        assert rep.nThermoCorr == len(its) and rep.nSolveIterations == sum(its) and rep.lastSolveIterations == its[-1], f"step {s}"
        exact(o.field(abi.FIELD_T), want["T"], f"T, step {s}")
        exact(o.field(abi.FIELD_ALPHA_SOLID), want["alpha1"], f"alpha.solid, step {s}")
        for p, (Tb, vf) in enumerate(o.patch_state()):
            exact(Tb, want["Tb"][p], f"patch {p} values, step {s}")
            exact(vf, want["valueFraction"][p], f"patch {p} value fraction, step {s}")
    assert s == steps - 1


def test_fixture_covers_the_branches_of_the_equation():
    ex = {name: mk.configs()[name][0]["controls"]["explicitSolve"] for name in NAMES}
    assert ex["amb_explicit"] == 1 and ex["amb_implicit_pcg"] == 0
    assert any(len(r["iters"]) >= 3 for r in GOLD["track_tmax"])                      # the corrector loop iterates
    assert all(it == 0 for r in GOLD["amb_explicit"] for it in r["iters"])            # diagonal solves
    assert any(it > 1 for r in GOLD["amb_implicit_bicg"] for it in r["iters"])
    case = mk.configs()["track_tmax"][0]
    assert max(H(v) for r in GOLD["track_tmax"] for v in r["T"]["val"]) <= case["controls"]["Tmax"] * (1 + 1e-6)   # the penalty holds T at Tmax
    # the backward cases differ from their Euler twins (same cases otherwise) from the second step on, and the first corrector of the
    # first step is where alpha.solid has no old-time level yet
    assert GOLD["amb_backward_pcg"][0]["T"]["sum"] != GOLD["amb_implicit_pcg"][0]["T"]["sum"]
    assert GOLD["track_backward_tmax"][-1]["T"]["val"] != GOLD["track_tmax"][-1]["T"]["val"]
    # the CrankNicolson cases: different from Euler and from backward; the as-shipped one takes explicit (Euler) steps while ddt0(T)
    # lives on in setDeltaT.H
    assert GOLD["amb_cn09_pcg"][-1]["T"]["sum"] not in (GOLD["amb_implicit_pcg"][-1]["T"]["sum"], GOLD["amb_backward_pcg"][-1]["T"]["sum"])
    assert GOLD["pbf_cn05"][-1]["T"]["val"] != GOLD["pbf"][-1]["T"]["val"]
    assert all(it == 0 for r in GOLD["amb_cn09_as_shipped_explicit"] for it in r["iters"])
    assert max(H(v) for r in GOLD["track_tmax"] for v in r["T"]["val"]) > 1620.0 and max(H(v) for r in GOLD["amb_explicit"] for v in r["T"]["val"]) > 1410.0   # melting / mushy zone


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_equation_matches_the_reference_text(ctx, name):
    case, _ = mk.configs()[name]
    g = lib.Case(ctx, case)
    for s, want in enumerate(GOLD[name]):
        rep = g.step()
        its = want["iters"]
        assert rep.nThermoCorr == len(its) and abs(rep.lastSolveIterations - its[-1]) <= 1, f"step {s}"
        close(g.field(abi.FIELD_T), want["T"], f"T, step {s}", 1e-9)
        close(g.field(abi.FIELD_ALPHA_SOLID), want["alpha1"], f"alpha.solid, step {s}", 1e-9, scale=1.0)
    g.close()


def test_crank_nicolson_with_zero_off_centring_is_euler(oracle):
    """psi = 0 switches every off-centre term off and leaves coef_ = 1: the run must be the Euler run bit for bit, while the scheme's
    bookkeeping (ddt0 fields, second old-time levels) runs all the same.  With psi > 0 the first TEqn.H already uses coef_ = 1 + psi,
    because setDeltaT.H created ddt0(T) one time index earlier; ddt0(alpha.solid) is created inside the first corrector loop."""
    kw = dict(n=(40, 10, 6), writeInterval=0.0, explicitSolve=0, maxDi=10.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC)
    from additivefoam_b200 import cases
    e = ol.OracleCase(cases.amb2018_02_b(**kw), lib=oracle)
    z = ol.OracleCase(cases.amb2018_02_b(ddtScheme=abi.DDT_CRANK_NICOLSON, ocCoeff=0.0, **kw), lib=oracle)
    c = ol.OracleCase(cases.amb2018_02_b(ddtScheme=abi.DDT_CRANK_NICOLSON, ocCoeff=0.9, **kw), lib=oracle)
    for s in range(20):
        re_, rz, rc = e.step(), z.step(), c.step()
        assert np.array_equal(e.field(abi.FIELD_T), z.field(abi.FIELD_T)) and rz.deltaT == re_.deltaT, s
        assert np.array_equal(e.field(abi.FIELD_ALPHA_SOLID), z.field(abi.FIELD_ALPHA_SOLID)), s
        st = c.cn_snapshot()[0]
        assert st["timeIndex"] == s + 1 and st["existsT"] == 1 and st["startT"] == 0 and st["tiT"] == s      # brought to index s by setDeltaT.H
        assert (st["existsA"], st["startA"]) == ((0, 0) if s == 0 else (1, 1))
        assert st["TnOld"] == (1 if s == 0 else 2)
    assert not np.array_equal(c.field(abi.FIELD_T), e.field(abi.FIELD_T))
    assert np.abs(z.field(abi.FIELD_DDT0_T)).max() > 0.0          # ddt0 = (T.old - T.oldOld)/deltaT0 is kept even though psi = 0 never uses it


@pytest.mark.gpu
def test_gpu_crank_nicolson_with_zero_off_centring_is_euler(ctx):
    """The same identity on the device: CnCoeffs with psi = 0 takes the CrankNicolson branches of k_props, k_assemble_implicit and
    k_corrector and must reproduce the Euler branches bit for bit."""
    from additivefoam_b200 import cases
    kw = dict(n=(60, 12, 8), writeInterval=0.0, explicitSolve=0, maxDi=10.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC)
    e = lib.Case(ctx, cases.amb2018_02_b(**kw))
    z = lib.Case(ctx, cases.amb2018_02_b(ddtScheme=abi.DDT_CRANK_NICOLSON, ocCoeff=0.0, **kw))
    for s in range(30):
        re_, rz = e.step(), z.step()
        assert rz.deltaT == re_.deltaT and rz.lastSolveIterations == re_.lastSolveIterations, s
    assert np.array_equal(e.field(abi.FIELD_T), z.field(abi.FIELD_T)) and e.field(abi.FIELD_T).max() > 1000.0
    assert np.array_equal(e.field(abi.FIELD_ALPHA_SOLID), z.field(abi.FIELD_ALPHA_SOLID))
    assert np.abs(z.field(abi.FIELD_DDT0_T)).max() > 0.0
    e.close(); z.close()
