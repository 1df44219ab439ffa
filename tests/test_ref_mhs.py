"""movingHeatSourceModel::adjustDeltaT / ::update (movingHeatSourceModel.C:92-150) against values computed by the REFERENCE's own
code (tests/golden/ref_mhs.json, made by tests/golden/make_ref_mhs_golden.py from oracle/_ref/refMHS = the reference's
movingHeatSourceModel.C with its real header on top of its heatSourceModel.C, shapes, movingBeam.C and segment.C, compiled where
they lie): every beam sub-cycled over the time step with its own deltaT (ragged last sub-step), dt-weighted mean, sum over
beams, the time step cut at path intervals, beams that dwell unpowered or have run out.  Checked: the oracle's restatement
(Case::updateSources + movingBeam::adjustDeltaT) and the library (b200_case_sources_update: host movingBeam, k_qdot_weights,
k_qdot_apply, the sub-cycle accumulation on the device)."""
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from additivefoam_b200 import abi, cases, lib

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_mhs.json")))


def beam_of(b):
    if b["kind"] == "super":
        d = cases.super_gaussian_beam(b["path"], dims=b["dims"], nPoints=b["nPoints"], k=b["k"], eta=b["eta"])
    elif b["kind"] == "modified":
        d = cases.modified_super_gaussian_beam(b["path"], dims=b["dims"], nPoints=b["nPoints"], k=b["k"], m=b["m"], transient=0)
        d["absorptionModel"], d["eta"] = abi.ABSORPTION_CONSTANT, b["eta"]
    else:
        d = cases.projected_gaussian_beam(b["path"], dims=b["dims"], nPoints=b["nPoints"], A=b["A"], B=b["B"], transient=0, eta=b["eta"])
    d["deltaT"], d["hitPathIntervals"] = b["deltaT"], b["hit"]
    if b.get("kelly"):
        geometry, eta0, etaMin = b["kelly"]
        d["absorptionModel"], d["eta0"], d["etaMin"] = abi.ABSORPTION_KELLY, eta0, etaMin
        d["kellyGeometry"] = abi.KELLY_CONE if geometry == "cone" else abi.KELLY_CYLINDER
    return d


def case_of(g):
    case = cases.amb2018_02_b(n=tuple(g["mesh"]["n"]), startTime=g["start"], endTime=g["end"])
    case["mesh"]["min"], case["mesh"]["max"] = g["mesh"]["min"], g["mesh"]["max"]
    case["beams"] = [beam_of(b) for b in g["beams"]]
    return case


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(GOLD)))
def test_gpu_sources_update_matches_the_reference_code(ctx, i):
    g = GOLD[i]
    c = lib.Case(ctx, case_of(g))
    n = int(np.prod(g["mesh"]["n"]))
    for s, (proposed, want) in enumerate(zip(g["proposed"], g["steps"])):
        dt = c.sources_update(proposed)
        assert dt == float.fromhex(want["dt"]), f"step {s}: adjusted deltaT"        # host arithmetic: bit for bit
        got = c.field(abi.FIELD_QDOT)
        ref = np.zeros(n)
        ref[want["cells"]] = [float.fromhex(v) for v in want["qDot"]]
        assert np.array_equal(got != 0.0, ref != 0.0), f"step {s}: cells touched"
        scale = max(np.max(np.abs(ref)), 1e-300)
        # exp/pow of the device differ from glibc's by <= 1 ulp
        assert np.max(np.abs(got - ref)) <= 1e-11 * scale, f"step {s}: {np.max(np.abs(got - ref)) / scale}"
    c.close()


@pytest.mark.parametrize("i", range(len(GOLD)))
def test_oracle_sources_update_matches_the_reference_code(oracle, i):
    g = GOLD[i]
    o = ol.OracleCase(case_of(g), lib=oracle)
    n = int(np.prod(g["mesh"]["n"]))
    for s, (proposed, want) in enumerate(zip(g["proposed"], g["steps"])):
        dt = o.sources_probe(proposed)
        assert dt == float.fromhex(want["dt"]), f"step {s}: adjusted deltaT"        # bit for bit
        got = o.field(abi.FIELD_QDOT)
        ref = np.zeros(n)
        ref[want["cells"]] = [float.fromhex(v) for v in want["qDot"]]
        assert np.array_equal(got != 0.0, ref != 0.0), f"step {s}: cells touched"
        scale = max(np.max(np.abs(ref)), 1e-300)
        # exp/pow: the oracle is built -O3 -march=native (libmvec paths), the reference -O2; 1e-13 as for qDot()
        assert np.max(np.abs(got - ref)) <= 1e-13 * scale, f"step {s}: {np.max(np.abs(got - ref)) / scale}"


def test_golden_covers_the_cases_it_claims():
    """sub-steps, a cut time step, an unpowered dwell and a finished path are all present in the fixture."""
    dts = [[float.fromhex(s["dt"]) for s in g["steps"]] for g in GOLD]
    assert any(abs(d - p) > 1e-3 * p for g, ds in zip(GOLD, dts) for d, p in zip(ds, g["proposed"]))      # adjustDeltaT acted
    assert any(len(s["cells"]) == 0 for g in GOLD for s in g["steps"])                                    # nothing deposited
    assert any(len(g["beams"]) == 2 for g in GOLD)
    assert any(b["deltaT"] > 0 and b["deltaT"] < min(g["proposed"]) for g in GOLD for b in g["beams"])    # sub-cycling
