"""setDeltaT.H and solutionControls.H -- the two fragments of the reference's time loop that set the time step and the per-step
solution controls -- against values computed by the REFERENCE's own text (tests/golden/ref_ctl.json, made by
tests/golden/make_ref_ctl_golden.py from oracle/_ref/refCtl = both fragments included verbatim where they lie against stand-in
types).  The fragments were fed the oracle's state at the start of every step of five small cases (as-shipped explicit; explicit
requested but the diffusion number growing past 1, then the beam running out; implicit with non-default controls; beam off;
powder layer with Tmax), so re-running the cases must give, step by step, the reference's thermo number, diffusion number and
new deltaT bit for bit, its absorbed power, its explicit/implicit decision and its corrector cap.  Checked: the oracle (CPU,
exact) and the GPU stepper through b200_case_step (1e-9: its T differs from the oracle's in the last digits)."""
import json
import os
import sys

import pytest

import oracle_lib as ol
from additivefoam_b200 import lib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_ref_ctl_golden as mk  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "ref_ctl.json")))
NAMES = sorted(GOLD)
H = float.fromhex


def rows(name):
    for r in GOLD[name]:
        yield dict(alphaCoNum=H(r[0]), DiNum=H(r[1]), deltaT=H(r[2]), nThermoCorr=H(r[3]), thermoTol=H(r[4]), Tmax=H(r[5]),
                   totalPower=H(r[6]), beamOn=int(r[7]), fluidInDomain=int(r[8]), explicitSolve=int(r[9]))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_controls_match_the_reference_text(oracle, name):
    case, steps = mk.configs()[name]
    o = ol.OracleCase(case, lib=oracle)
    ct = case["controls"]
    for s, want in enumerate(rows(name)):
        rep = o.step()
        assert rep.alphaCoNum == want["alphaCoNum"] and rep.DiNum == want["DiNum"] and rep.deltaT == want["deltaT"], f"step {s}"
        assert rep.totalPower == pytest.approx(want["totalPower"], rel=1e-14, abs=1e-300), f"step {s}"     # gSum over ranks vs one sum
        assert rep.explicitStep == want["explicitSolve"], f"step {s}"
        assert rep.nThermoCorr <= want["nThermoCorr"] and (want["nThermoCorr"] > 1 or rep.nThermoCorr == 1), f"step {s}"
        assert want["beamOn"] == int(rep.totalPower > 1e-15)
        # the defaults / entries the fragment read are the ones the case description carries
        assert want["thermoTol"] == ct["thermoTolerance"] and want["Tmax"] == (ct["Tmax"] if ct["Tmax"] > 0 else 1e300)
    assert s == steps - 1


def test_fixture_covers_both_decisions():
    ex = {name: {r["explicitSolve"] for r in rows(name)} for name in NAMES}
    assert ex["amb_switches"] == {0, 1} and ex["amb_explicit"] == {1} and ex["track_implicit"] == {0}
    assert any(r["nThermoCorr"] == 1.0 for r in rows("beam_off")) and any(r["fluidInDomain"] == 1 for r in rows("pbf"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_controls_match_the_reference_text(ctx, name):
    case, _ = mk.configs()[name]
    g = lib.Case(ctx, case)
    for s, want in enumerate(rows(name)):
        rep = g.step()
        assert rep.deltaT == pytest.approx(want["deltaT"], rel=1e-9), f"step {s}"
        assert rep.alphaCoNum == pytest.approx(want["alphaCoNum"], rel=1e-8, abs=1e-12), f"step {s}"
        assert rep.DiNum == pytest.approx(want["DiNum"], rel=1e-9), f"step {s}"
        assert rep.totalPower == pytest.approx(want["totalPower"], rel=1e-9, abs=1e-12), f"step {s}"
        assert rep.explicitStep == want["explicitSolve"], f"step {s}"
        assert rep.nThermoCorr <= want["nThermoCorr"] and (want["nThermoCorr"] > 1 or rep.nThermoCorr == 1), f"step {s}"
    g.close()
