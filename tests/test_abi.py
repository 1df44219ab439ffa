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
    assert L.b200_abi_version() == 1


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
