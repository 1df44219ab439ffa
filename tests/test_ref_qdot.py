"""heatSourceModel::qDot() against values computed by the REFERENCE's own code (tests/golden/ref_qdot.json, made by
tests/golden/make_ref_qdot_golden.py from oracle/_ref/refQDot = the reference's heatSourceModel.C with its real header and the
three shape sources, compiled where they lie against a stand-in uniform hex block): the beam bounding box, the per-cell
sample-point count, the Riemann sum of the shape and the 5 % normalisation rule -- including the cases where the rule does
not apply (coarse sampling, beam mostly outside the block).  Checked: the oracle (CPU) and b200_beam_qdot (GPU)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from additivefoam_b200 import abi, cases, lib

HERE = os.path.dirname(os.path.abspath(__file__))
_ALL = json.load(open(os.path.join(HERE, "golden", "ref_qdot.json")))
GOLD, DEPTH = _ALL["qdot"], _ALL["depth"]
PATH = [[1, 0, 0, 0, 100.0, 0.0], [0, 1e-3, 0, 0, 100.0, 0.8]]


def beam_of(g):
    if g["kind"] == "super":
        b = cases.super_gaussian_beam(PATH, dims=g["dims"], nPoints=g["nPoints"], k=g["k"], eta=g["eta"])
    elif g["kind"] == "modified":
        b = cases.modified_super_gaussian_beam(PATH, dims=g["dims"], nPoints=g["nPoints"], k=g["k"], m=g["m"], transient=0)
        b["absorptionModel"], b["eta"] = abi.ABSORPTION_CONSTANT, g["eta"]
    else:
        b = cases.projected_gaussian_beam(PATH, dims=g["dims"], nPoints=g["nPoints"], A=g["A"], B=g["B"], transient=0, eta=g["eta"])
    return b


def dense(g):
    q = np.zeros(g["n"][0] * g["n"][1] * g["n"][2])
    q[g["cells"]] = [float.fromhex(v) for v in g["qdot"]]
    return q


@pytest.mark.parametrize("i", range(len(GOLD)))
def test_oracle_qdot_matches_the_reference_code(oracle, i):
    g = GOLD[i]
    bm = abi.BlockMesh()
    bm.nx, bm.ny, bm.nz = g["n"]
    for q in range(3):
        bm.min[q], bm.max[q] = g["min"][q], g["max"][q]
    beam = cases.make_beam(beam_of(g))
    pos, dims = np.array(g["position"]), np.array(g["dims"])
    got = np.zeros(bm.nx * bm.ny * bm.nz)
    sumw, vol, absorbed = np.zeros(1), np.zeros(1), np.zeros(1)
    oracle.afo_beam_qdot(C.byref(bm), C.byref(beam), ol.P(pos), ol.P(dims), g["power"], ol.P(got), ol.P(sumw), ol.P(vol), ol.P(absorbed))
    want = dense(g)
    assert np.array_equal(got != 0.0, want != 0.0)                     # exactly the same cells are touched
    scale = max(np.max(np.abs(want)), 1e-300)
    assert np.max(np.abs(got - want)) <= 1e-13 * scale, np.max(np.abs(got - want)) / scale
    V = np.prod((np.array(g["max"]) - np.array(g["min"])) / np.array(g["n"]))
    assert float(got.sum() * V) == pytest.approx(float.fromhex(g["power_deposited"]), rel=1e-12, abs=1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(GOLD)))
def test_gpu_qdot_matches_the_reference_code(ctx, i):
    g = GOLD[i]
    mesh = dict(n=g["n"], min=g["min"], max=g["max"])
    got, sumw, vol, absorbed = lib.beam_qdot(ctx, mesh, beam_of(g), g["position"], g["dims"], g["power"])
    want = dense(g)
    assert np.array_equal(got != 0.0, want != 0.0)
    scale = max(np.max(np.abs(want)), 1e-300)
    # exp/pow of the device differ from glibc's by <= 1 ulp
    assert np.max(np.abs(got - want)) <= 1e-11 * scale
    assert absorbed == pytest.approx(g["eta"] * g["power"], rel=1e-14)


def depth_beam(g):
    if g["kind"] == "modified":
        return cases.modified_super_gaussian_beam(PATH, dims=g["dims"], k=g["k"], m=g["m"], transient=1, isoValue=g["isoValue"])
    return cases.projected_gaussian_beam(PATH, dims=g["dims"], A=g["A"], B=g["B"], transient=1, isoValue=g["isoValue"])


def depth_case(g, b):
    case = cases.amb2018_02_b(n=tuple(g["n"]))
    case["mesh"]["min"], case["mesh"]["max"], case["beams"] = g["min"], g["max"], [b]
    return case


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(DEPTH)))
def test_gpu_isotherm_depth_matches_the_reference_code(ctx, i):
    """k_isotherm_depth through b200_case_isotherm_depth against the reference's updateDimensions()."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_ref_qdot_golden as mk
    g = DEPTH[i]
    c = lib.Case(ctx, depth_case(g, depth_beam(g)))
    c.set_field(abi.FIELD_T, np.array(mk.hot_spot(g["n"], g["min"], g["max"], (g["spot"][0], g["spot"][1], g["spot"][2]))))
    got = c.isotherm_depth(0, g["position"])
    c.close()
    want = [float.fromhex(v) for v in g["dimensions"]]
    # cell centres are origin + (i + 0.5) d on both sides; the interpolation is the same expression
    assert got[0] == want[0] and got[1] == want[1] and got[2] == pytest.approx(want[2], rel=1e-14)


@pytest.mark.parametrize("i", range(len(DEPTH)))
def test_oracle_isotherm_depth_matches_the_reference_code(oracle, i):
    """heatSourceModel::updateDimensions (heatSourceModel.C:123-219) as the reference's own code runs it on a hot-spot temperature
    field: depth beyond, equal to and capped by the static one, isotherm outside the search radius, non-cubic cells.  The CUDA
    kernel (k_isotherm_depth) is held to the oracle by the transient-beam cases of test_gpu_case.py / test_gpu_fuzz.py."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_ref_qdot_golden as mk
    g = DEPTH[i]
    if g["kind"] == "modified":
        b = cases.modified_super_gaussian_beam(PATH, dims=g["dims"], k=g["k"], m=g["m"], transient=1, isoValue=g["isoValue"])
    else:
        b = cases.projected_gaussian_beam(PATH, dims=g["dims"], A=g["A"], B=g["B"], transient=1, isoValue=g["isoValue"])
    case = depth_case(g, b)
    o = ol.OracleCase(case, lib=oracle)
    o.set_field(abi.FIELD_T, np.array(mk.hot_spot(g["n"], g["min"], g["max"], (g["spot"][0], g["spot"][1], g["spot"][2]))))
    pos, dims = np.array(g["position"]), np.zeros(3)
    assert oracle.afo_case_isotherm_depth(o.h, 0, ol.P(pos), ol.P(dims)) == 0
    o.close()
    assert list(dims) == [float.fromhex(v) for v in g["dimensions"]]        # same interpolation expression: bit for bit
