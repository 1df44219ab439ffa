"""Beam kinematics against values computed by the REFERENCE's own code (tests/golden/ref_beam.json, made by
tests/golden/make_ref_beam_golden.py from oracle/_ref/refBeam = the reference's movingBeam.C and segment.C compiled where they
lie): scan-path reading, segment times, findIndex with its carried index, move (position, power with the 1e-10 switch),
activePath, and the Delta-t snapping of adjustDeltaT.  Checked: the oracle's restatement (CPU) and, on the GPU, the library's
host classes through a case whose reported time steps and deposited power they drive."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from additivefoam_b200 import abi, cases, lib

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_beam.json")))


def beam_of(cfg):
    path = lib.scan_path_parse(GOLD["paths"][cfg["path"]]).tolist()
    b = cases.super_gaussian_beam(path)
    b["deltaT"] = cfg["deltaT"] if cfg["deltaT"] > 0 else 0.0
    b["hitPathIntervals"] = cfg["hitPathIntervals"]
    return b


@pytest.mark.parametrize("i", range(len(GOLD["configs"])))
def test_oracle_move_matches_the_reference_code(oracle, i):
    cfg = GOLD["configs"][i]
    b = cases.make_beam(beam_of(cfg))
    t = np.array(cfg["times"])
    pos, pw, idx, et = np.empty(3 * len(t)), np.empty(len(t)), np.empty(len(t), dtype=np.int32), C.c_double()
    oracle.afo_beam_move(C.byref(b), cfg["start"], cfg["end"], len(t), ol.P(t), ol.P(pos), ol.P(pw), ol.P(idx),
                         C.cast(C.byref(et), abi.c_double_p))
    for q, m in enumerate(cfg["moves"]):
        want = [float.fromhex(v) for v in m[:4]]
        assert list(pos[3 * q: 3 * q + 3]) == want[:3], (q, t[q])          # same expression order: bit for bit
        assert pw[q] == want[3], (q, t[q])
        assert int((et.value - t[q]) > 1e-10) == int(m[4]), (q, t[q])       # activePath (movingBeam.C:151-154)


@pytest.mark.parametrize("i", range(len(GOLD["configs"])))
def test_oracle_adjustDeltaT_matches_the_reference_code(oracle, i):
    cfg = GOLD["configs"][i]
    b = cases.make_beam(beam_of(cfg))
    for a in cfg["adjust"]:
        got = oracle.afo_beam_adjust_deltaT(C.byref(b), cfg["start"], cfg["end"], a["moveTo"], a["now"], a["dt"])
        assert got == float.fromhex(a["result"]), a


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["raster", "awkward"])
def test_gpu_time_stepping_follows_the_pinned_oracle(ctx, oracle, name):
    """The library's movingBeam/segment host classes drive Delta t (adjustDeltaT) and the deposited power; with the oracle's
    kinematics pinned on the reference code above, equal time steps and power over a whole path pin the library's too."""
    path = lib.scan_path_parse(GOLD["paths"][name]).tolist()
    case = cases.amb2018_02_b(n=(40, 16, 6), explicitSolve=0, maxDi=50.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC,
                              endTime=0.02)
    case["mesh"]["min"], case["mesh"]["max"] = [-2e-4, -3e-4, -1.2e-4], [2.2e-3, 7e-4, 0.0]
    case["beams"] = [cases.super_gaussian_beam(path)]
    case["controls"]["deltaT"] = 1e-6
    g, o = lib.Case(ctx, case), ol.OracleCase(case, lib=oracle)
    powers = set()
    for s in range(400):
        if not o.running():
            break
        rg, ro = g.step(), o.step()
        assert rg.deltaT == pytest.approx(ro.deltaT, rel=1e-11), s
        assert rg.time == pytest.approx(ro.time, rel=1e-12), s
        assert rg.totalPower == pytest.approx(ro.totalPower, rel=1e-9, abs=1e-12), s
        powers.add(round(ro.totalPower, 3))
    assert g.running() == o.running()
    assert len(powers) >= 2          # the beam did switch between powered and unpowered segments
    g.close(); o.close()
