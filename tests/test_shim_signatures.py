"""The OpenFOAM-side shims (openfoam/bPCG: lduMatrix::solver; openfoam/b200HeatSource: heatSourceModel) must at least parse against declarations that mirror the
upstream interface (OpenFOAM-10 is not installable here) and call only functions the C header declares."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bpcg_shim_parses_against_stub_headers():
    cmd = ["g++", "-std=c++14", "-fsyntax-only", "-I" + os.path.join(ROOT, "openfoam", "stubs"),
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "openfoam", "bPCG", "bPCG.C")]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]


def test_heat_source_shim_parses_against_stub_headers():
    cmd = ["g++", "-std=c++14", "-fsyntax-only", "-I" + os.path.join(ROOT, "openfoam", "stubs"),
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "openfoam", "b200HeatSource", "b200HeatSources.C")]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]


def test_resident_solver_application_parses_against_stub_headers():
    """openfoam/b200AdditiveFoam: additiveFoam's time loop around b200_case_step."""
    cmd = ["g++", "-std=c++14", "-fsyntax-only", "-I" + os.path.join(ROOT, "openfoam", "stubs"),
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "openfoam", "b200AdditiveFoam", "b200AdditiveFoam.C")]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    src = open(os.path.join(ROOT, "openfoam", "b200AdditiveFoam", "b200AdditiveFoam.C")).read()
    for call in ("b200_case_create", "b200_case_step", "b200_case_set_field", "b200_case_reset_alpha_solid", "b200_case_get_field", "runTime.write()"):
        assert call in src


def test_heat_source_shim_registers_the_three_shapes():
    src = open(os.path.join(ROOT, "openfoam", "b200HeatSource", "b200HeatSources.C")).read()
    for cls, shape in (("b200SuperGaussian", "B200_SUPER_GAUSSIAN"), ("b200ModifiedSuperGaussian", "B200_MODIFIED_SUPER_GAUSSIAN"),
                       ("b200ProjectedGaussian", "B200_PROJECTED_GAUSSIAN")):
        assert f"addToRunTimeSelectionTable(heatSourceModel, {cls}, dictionary)" in src
        assert f"B200_DEFINE_HEAT_SOURCE({cls}, {shape})" in src


def test_shims_call_only_declared_entry_points():
    hdr = open(os.path.join(ROOT, "include", "additivefoam_b200.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z_0-9]+)\s*\(", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)))
    for rel in ("bPCG/bPCG.C", "b200HeatSource/b200HeatSources.C", "b200AdditiveFoam/b200AdditiveFoam.C"):
        src = open(os.path.join(ROOT, "openfoam", rel)).read()
        used = set(re.findall(r"\b(b200_[a-z_0-9]+)\s*\(", src))
        assert used and used <= declared, used - declared


def test_bpcg_registers_in_both_tables():
    src = open(os.path.join(ROOT, "openfoam", "bPCG", "bPCG.C")).read()
    for cls in ("bPCG", "bPBiCGStab"):
        assert f"addsymMatrixConstructorToTable<{cls}>" in src and f"addasymMatrixConstructorToTable<{cls}>" in src
