"""solidificationData (functionObjects/solidificationData/solidificationData.C: setOverlapCells :99-131, execute :142-183, end
:189-222) against rows written by the REFERENCE's own code (tests/golden/ref_solid.json, made by
tests/golden/make_ref_solid_golden.py from oracle/_ref/refFO = the reference's solidificationData.C with its real header,
compiled where it lies, fed the temperature fields of 400 consecutive steps of a small single-track case): which cells the box
selects (incl. box faces lying on cell faces, where the 1e-10 growth of the cell boxes decides), which cells cooled through the
isotherm in which step and in which order, the cooling rate kept from the PREVIOUS execute, |grad T| (Gauss linear, boundary
faces with the patch value), their ratio.  Checked: the oracle (CPU, bit for bit) and k_solidification through
b200_case_solidification_* (GPU: positions exact, times 1e-11, rates and gradients to the accuracy T carries)."""
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib as ol
from additivefoam_b200 import lib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_ref_solid_golden as mk  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "ref_solid.json")))
NAMES = sorted(GOLD)


def want(name):
    return np.array([[float.fromhex(v) for v in r] for r in GOLD[name]]).reshape(-1, 7)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_events_match_the_reference_code(oracle, name):
    case, steps, lo, hi, iso, _ = mk.configs()[name]
    o = ol.OracleCase(case, lib=oracle)
    o.solidification_enable(lo, hi, iso)
    for _ in range(steps):
        o.step()
    got, ref = o.solidification_events(), want(name)
    assert len(ref) > 20 and got.shape == ref.shape
    assert np.array_equal(got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_events_match_the_reference_code(ctx, name):
    case, steps, lo, hi, iso, _ = mk.configs()[name]
    g = lib.Case(ctx, case)
    g.solidification_enable(lo, hi, iso)
    for _ in range(steps):
        g.step()
    got, ref = g.solidification_events(), want(name)
    g.close()
    assert got.shape == ref.shape
    assert np.array_equal(got[:, :3], ref[:, :3])
    assert np.allclose(got[:, 3], ref[:, 3], rtol=1e-11, atol=0)
    assert np.allclose(got[:, 4:], ref[:, 4:], rtol=1e-6)
