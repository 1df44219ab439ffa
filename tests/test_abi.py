"""The C ABI: the library loads, exports every symbol the header declares, parses scan paths, and refuses to compute
without a CUDA device (no CPU fallback).  No GPU needed."""
import os
import re

import numpy as np
import pytest

from additivefoam_b200 import abi, cases, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "additivefoam_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = lib.load()
    declared = header_symbols()
    assert declared, "no declarations parsed"
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(lib.SYMBOLS) == declared          # the python binding names exactly the header's surface
    assert L.b200_abi_version() == 3          # 2: b200_controls.ddtScheme; 3: ocCoeff (CrankNicolson)


def test_struct_sizes_match_header_layout():
    # layouts are plain C: spot-check sizes that would change if a field were dropped or reordered
    import ctypes as C
    assert C.sizeof(abi.BlockMesh) == 3 * 4 + 4 + 6 * 8
    assert C.sizeof(abi.SolverPerf) == 2 * 8 + 3 * 4 + 4
    assert C.sizeof(abi.Patch) == 8 * 4 + 4 * 8
    assert abi.StepReport().as_dict().keys() >= {"time", "deltaT", "nThermoCorr", "lastSolveIterations"}


def test_scan_path_parse_shipped_format():
    text = "Mode    X       Y       Z   Power   Param\n1       0.000   0.000   0   0       0\n\n0       0.002   0.000   0   179.2   0.8\n"
    p = lib.scan_path_parse(text)
    assert p.shape == (2, 6) and list(p[1]) == [0, 0.002, 0, 0, 179.2, 0.8]
    assert lib.scan_path_parse("header only\n").shape == (0, 6)
    with pytest.raises(lib.B200Error):
        lib.scan_path_parse(text, max_segments=1)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(lib.B200Error, match="no CUDA device"):
        lib.Context(0)


def test_case_descriptions_build():
    for case in (cases.amb2018_02_b(), cases.synthetic_track((100, 40, 25)), cases.rosenthal((50, 20, 10)),
                 cases.multi_beam((100, 25, 20)), cases.multi_layer_pbf((100, 50, 40))):
        b = cases.BuiltCase(case)
        assert b.desc.mesh.nx * b.desc.mesh.ny * b.desc.mesh.nz == b.n_cells
        assert b.desc.nPatches == 3 and b.desc.nBeams == len(case["beams"])
        sides = sorted(s for p in case["patches"] for s in p["sides"])
        assert sides == list(range(6))
    pbf = cases.multi_layer_pbf((100, 50, 40))
    assert pbf["controls"]["Tmax"] == 3300.0 and pbf["powderZMin"] == -40e-6
    assert len(cases.multi_beam((100, 25, 20))["beams"]) == 4


def test_host_shape_functions_match_oracle_and_known_answers(oracle):
    """b200_beam_shape_v0 / _weight / b200_beam_absorption_eta (what the OpenFOAM heat-source shims forward their pure virtuals
    to) are host code: checked here against the oracle's restatement and the hand-derived values of tests/golden."""
    import ctypes as C
    import json
    import oracle_lib as ol
    ka = json.load(open(os.path.join(ROOT, "tests", "golden", "known_answers.json")))
    path = [[1, 0, 0, 0, 100.0, 0.0], [0, 1e-3, 0, 0, 100.0, 0.8]]
    beams = [cases.super_gaussian_beam(path), cases.modified_super_gaussian_beam(path), cases.projected_gaussian_beam(path)]
    v0, _ = lib.beam_shape_V0(beams[0], (85e-6, 85e-6, 30e-6))
    assert v0 == pytest.approx(ka["superGaussian_V0_k2_85_85_30um"], rel=1e-14)
    v0, _ = lib.beam_shape_V0(beams[1], (40e-6, 40e-6, 30e-6))
    assert v0 == pytest.approx(ka["modifiedSuperGaussian_V0_k7.95_m2.72_40_40_30um"], rel=1e-13)
    assert lib.beam_absorption_eta(beams[1], 0.75) == ka["kelly_cone_eta_aspect_0.75"]
    assert lib.beam_absorption_eta(beams[1], 2.0) == pytest.approx(ka["kelly_cone_eta_aspect_2"], rel=1e-14)
    rng = np.random.default_rng(5)
    for b in beams:
        ob = cases.make_beam(b)
        for depth in (30e-6, 55e-6, 170e-6):          # transient beams change dimensions.z (heatSourceModel.C:207-213)
            dims = np.array([b["dimensions"][0], b["dimensions"][1], depth])
            assert lib.beam_shape_V0(b, dims)[0] == pytest.approx(oracle.afo_shape_V0(C.byref(ob), ol.P(dims)), rel=1e-14)
            for _ in range(40):
                d = np.abs(rng.normal(0, 1, 3)) * dims * 0.8
                got, ref = lib.beam_shape_weight(b, dims, d), oracle.afo_shape_weight(C.byref(ob), ol.P(dims), ol.P(d))
                assert got == pytest.approx(ref, rel=1e-13, abs=1e-300)
        for aspect in (0.3, 1.0, 1.0000001, 2.0, 7.5):
            assert lib.beam_absorption_eta(b, aspect) == pytest.approx(oracle.afo_absorption_eta(C.byref(ob), aspect), rel=1e-14)
