"""meltPoolDimensions::execute (functionObjects/meltPoolDimensions/meltPoolDimensions.C:123-286) against values computed by the
REFERENCE's own code (tests/golden/ref_fo.json, made by tests/golden/make_ref_fo_golden.py from oracle/_ref/refFO = the
reference's meltPoolDimensions.C with its real header, compiled where it lies against a stand-in uniform hex block with six
physical patches): isotherm crossings on internal faces, boundary faces whose cell or patch value reaches the isotherm (hot
fixedValue wall, cold fixedValue top, zeroGradient), rotation by the scan-path angle, the empty box.  Checked: the oracle (CPU)
and k_melt_pool through b200_case_melt_pool_dimensions (GPU).  No exp/pow on this path: both sides use the same operations,
so the comparison is to a few ulp of the coordinates (1e-12 of the block size)."""
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib as ol
from additivefoam_b200 import abi, cases, lib

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_fo.json")))
SIDES = [abi.XMIN, abi.XMAX, abi.YMIN, abi.YMAX, abi.ZMIN, abi.ZMAX]


def case_and_field(g):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_ref_qdot_golden as mk
    m = g["mesh"]
    case = cases.amb2018_02_b(n=tuple(m["n"]))
    case["mesh"]["min"], case["mesh"]["max"] = m["min"], m["max"]
    case["patches"] = []
    for side, p in zip(SIDES, g["patches"]):
        if p == "z":
            case["patches"].append(cases._patch(abi.BC_ZERO_GRADIENT, [side]))
        else:
            case["patches"].append(cases._patch(abi.BC_FIXED_VALUE, [side], value=float(p.split()[1])))
    (c, r, peak) = g["spot"]
    return case, np.array(mk.hot_spot(m["n"], m["min"], m["max"], (tuple(c), tuple(r), peak)))


def check(got, g):
    want = np.array([[float.fromhex(v) for v in d] for d in g["dimensions"]])
    size = np.max(np.array(g["mesh"]["max"]) - np.array(g["mesh"]["min"]))
    assert np.array_equal(got == 0.0, want == 0.0)
    assert np.max(np.abs(got - want)) <= 1e-12 * size, np.max(np.abs(got - want)) / size


@pytest.mark.parametrize("i", range(len(GOLD)))
def test_oracle_melt_pool_dimensions_match_the_reference_code(oracle, i):
    g = GOLD[i]
    case, T = case_and_field(g)
    o = ol.OracleCase(case, lib=oracle)
    o.set_field(abi.FIELD_T, T)
    check(o.melt_pool_dimensions(g["iso"], g["angle"]), g)
    o.close() if hasattr(o, "close") else None


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(GOLD)))
def test_gpu_melt_pool_dimensions_match_the_reference_code(ctx, i):
    g = GOLD[i]
    case, T = case_and_field(g)
    c = lib.Case(ctx, case)
    c.set_field(abi.FIELD_T, T)
    check(c.melt_pool_dimensions(g["iso"], g["angle"]), g)
    c.close()
