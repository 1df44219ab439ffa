"""foamcase.read_case: an AdditiveFOAM case directory -> case description.

Checked (CPU) on dictionaries written here, and -- where the reference tree is present -- on the reference's own shipped
tutorials, which at the same time pins the hand-restated descriptions of additivefoam_b200/cases.py on the shipped files.
The GPU test runs a case read from a directory against the oracle."""
import os
import textwrap

import numpy as np
import pytest

from additivefoam_b200 import abi, cases, foamcase

TUT = "/root/reference/tutorials"
needs_reference = pytest.mark.skipif(not os.path.isdir(TUT), reason="the reference tree is only present in the build container")

HEADER = "FoamFile\n{\n    version 2.0;\n    format ascii;\n    class dictionary;\n    object x;\n}\n"


def write_case(root, explicit="false", solver="PCG", precond="DIC", powder=False, second_beam=False):
    """A small case in the tutorials' layout with numbers of its own (20 x 8 x 6 cells of 25 um)."""
    files = {
        "system/blockMeshDict": HEADER + textwrap.dedent("""
            convertToMeters 1e-3;   // vertices in mm
            x0 -0.1; x1 0.4; y0 -0.1; y1 0.1; z0 -0.15; z1 0;
            vertices ( ($x0 $y0 $z0) ($x1 $y0 $z0) ($x1 $y1 $z0) ($x0 $y1 $z0) ($x0 $y0 $z1) ($x1 $y0 $z1) ($x1 $y1 $z1) ($x0 $y1 $z1) );
            blocks ( hex (0 1 2 3 4 5 6 7) (20 8 6) simpleGrading (1 1 1) );
            edges ( );
            boundary
            (
                base  { type wall; faces ( (0 3 2 1) ); }
                lid   { type wall; faces ( (4 5 6 7) ); }
                walls { type wall; faces ( (0 4 7 3) (2 6 5 1) (1 5 4 0) (3 7 6 2) ); }
            );
            """),
        "0/T": HEADER + textwrap.dedent("""
            dimensions [0 0 0 1 0 0 0];
            internalField uniform 320;
            boundaryField
            {
                base    { type fixedValue; value uniform 310; }
                lid     { type mixedTemperature; h 25.0; emissivity 0.3; Tinf uniform 295; value uniform 320; }
                "wal.*" { type zeroGradient; /* regular-expression key */ }
            }
            """),
        "constant/transportProperties": HEADER + textwrap.dedent("""
            solid  { kappa (9.0 0.015 0.0);  Cp (580.0 0.01 0.0); }
            liquid { kappa (5.0 0.014 0.0);  Cp (750.0 0.0 0.0); }
            powder { kappa (0.2 0.001 0.0);  Cp (320.0 0.0 0.0); }
            rho [1 -3 0 0 0 0 0] 7600;
            mu  [1 -1 -1 0 0 0 0] 0.003;
            Lf  [0 2 -2 0 0 0 0] 2.1e5;
            """),
        "constant/thermoPath": "(\n1400.0 1.0\n1500.0 0.6\n1600.0 0.0\n)\n",
        "constant/heatSourceDict": HEADER + ("sources (laser spot);\n" if second_beam else "sources (laser);\n") + textwrap.dedent("""
            laser
            {
                pathName track;
                absorptionModel constant;
                constantCoeffs { eta 0.4; }
                heatSourceModel superGaussian;
                superGaussianCoeffs { k 2.5; dimensions (60e-6 60e-6 40e-6); nPoints (4 4 4); }
            }
            spot
            {
                pathName track;
                deltaT 2e-6;
                absorptionModel Kelly;
                KellyCoeffs { geometry cylinder; eta0 0.3; etaMin 0.32; }
                heatSourceModel projectedGaussian;
                projectedGaussianCoeffs { A 1.5; B 0.5; dimensions (50e-6 50e-6 45e-6); transient true; isoValue 1600; }
            }
            """),
        "constant/track": "Mode X Y Z Power Param\n1 0.0 0.0 0 0 0\n\n0 3e-4 0.0 0 150 0.9\n",
        "system/controlDict": HEADER + textwrap.dedent("""
            application additiveFoam;
            startTime 0; endTime 2e-4; deltaT 5e-7;
            writeControl adjustableRunTime; writeInterval 1e-4;
            adjustTimeStep yes; maxCo 0.5; maxDi 8; maxAlphaCo 1;
            functions
            {
                #ifeq ${SOMETHING} true
                probe { type none; }
                #endif
            }
            """),
        "system/fvSchemes": HEADER + textwrap.dedent("""
            ddtSchemes { default Euler; }
            gradSchemes { default Gauss linear; }
            divSchemes { default Gauss upwind; }
            laplacianSchemes { default Gauss linear corrected; laplacian(kappa,T) Gauss harmonic corrected; }
            interpolationSchemes { default linear; }
            snGradSchemes { default corrected; }
            """),
        "system/fvSolution": HEADER + textwrap.dedent(f"""
            solvers
            {{
                p_rgh {{ solver PCG; preconditioner DIC; tolerance 1e-8; relTol 0.05; }}
                p_rghFinal {{ $p_rgh; relTol 0; }}
                "T.*" {{ solver {solver}; preconditioner {precond}; tolerance 1e-14; relTol 0; minIter 1; maxIter 30; }}
            }}
            PIMPLE {{ nOuterCorrectors 0; nThermoCorrectors 15; thermoTolerance 1e-7; explicitSolve {explicit}; Tmax 3200; }}
            """),
    }
    if powder:
        files["system/setFieldsDict"] = HEADER + textwrap.dedent("""
            defaultFieldValues ( volScalarFieldValue alpha.powder 0 );
            regions ( boxToCell { box (-1 -1 -5e-5) (1 1 1); fieldValues ( volScalarFieldValue alpha.powder 1 ); } );
            """)
    for rel, text in files.items():
        path = os.path.join(root, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(text)
    return str(root)


def test_dictionary_grammar():
    d = foamcase.parse_dictionary(HEADER + textwrap.dedent("""
        a 1; b 2.5e-3;      // trailing comment
        /* block
           comment */
        name  word;  flag yes;
        v (1 2 3);  nested ( (0 1) (2 3) );  dims [0 1 -2 0 0 0 0] 9.81;
        sub { c $a; deeper { e $b; } "re.*" { f 1; } }
        copy { $sub; g 2; }
        #include "somewhere"
        """))
    assert d["a"] == 1 and d["b"] == 2.5e-3 and d["name"] == "word" and d["v"] == [1, 2, 3] and d["nested"] == [[0, 1], [2, 3]]
    assert d["dims"] == [("dimensions", [0, 1, -2, 0, 0, 0, 0]), 9.81]
    assert d["sub"]["c"] == 1 and d["sub"]["deeper"]["e"] == 2.5e-3 and d["sub"]["re.*"]["f"] == 1
    assert d["copy"]["$include"] == ["sub"] and d["copy"]["g"] == 2
    with pytest.raises(foamcase.FoamCaseError):
        foamcase.parse_dictionary("a $missing;")
    with pytest.raises(foamcase.FoamCaseError):
        foamcase.parse_dictionary("a { b 1; ")


def test_read_case_from_written_dictionaries(tmp_path):
    c = foamcase.read_case(write_case(tmp_path, powder=True, second_beam=True))
    assert c["mesh"]["n"] == [20, 8, 6]
    assert np.allclose(c["mesh"]["min"], [-1e-4, -1e-4, -1.5e-4], rtol=1e-15) and np.allclose(c["mesh"]["max"], [4e-4, 1e-4, 0.0], rtol=1e-15)
    assert [(p["type"], p["sides"]) for p in c["patches"]] == [
        (abi.BC_FIXED_VALUE, [abi.ZMIN]), (abi.BC_MIXED_TEMPERATURE, [abi.ZMAX]), (abi.BC_ZERO_GRADIENT, [abi.XMIN, abi.XMAX, abi.YMIN, abi.YMAX])]
    assert c["patches"][0]["value"] == 310.0 and (c["patches"][1]["h"], c["patches"][1]["emissivity"], c["patches"][1]["Tinf"]) == (25.0, 0.3, 295.0)
    assert c["Tinitial"] == 320.0 and c["powderZMin"] == -5e-5
    m = c["material"]
    assert m["kappa"][2] == [0.2, 0.001, 0.0] and m["Cp"][0] == [580.0, 0.01, 0.0] and (m["rho"], m["Lf"]) == (7600.0, 2.1e5)
    assert m["thermoT"] == [1400.0, 1500.0, 1600.0] and m["thermoAlpha"] == [1.0, 0.6, 0.0]
    laser, spot = c["beams"]
    assert laser["heatSourceModel"] == abi.SUPER_GAUSSIAN and laser["k"] == 2.5 and laser["nPoints"] == [4, 4, 4] and laser["eta"] == 0.4
    assert laser["transient"] == 0 and laser["deltaT"] == 0.0 and laser["path"] == [[1, 0, 0, 0, 0, 0], [0, 3e-4, 0, 0, 150, 0.9]]
    assert spot["heatSourceModel"] == abi.PROJECTED_GAUSSIAN and (spot["A"], spot["B"]) == (1.5, 0.5) and spot["nPoints"] == [1, 1, 1]
    assert spot["absorptionModel"] == abi.ABSORPTION_KELLY and spot["kellyGeometry"] == abi.KELLY_CYLINDER
    assert (spot["eta0"], spot["etaMin"], spot["transient"], spot["isoValue"], spot["deltaT"]) == (0.3, 0.32, 1, 1600.0, 2e-6)
    k = c["controls"]
    assert (k["endTime"], k["deltaT"], k["adjustTimeStep"], k["maxDi"], k["writeInterval"]) == (2e-4, 5e-7, 1, 8.0, 1e-4)
    assert (k["solver"], k["preconditioner"], k["tolerance"], k["minIter"], k["maxIter"]) == (abi.SOLVER_PCG, abi.PRECOND_DIC, 1e-14, 1, 30)
    assert (k["nThermoCorrectors"], k["thermoTolerance"], k["explicitSolve"], k["Tmax"], k["maxDeltaT"]) == (15, 1e-7, 0, 3200.0, 1e15)
    assert cases.BuiltCase(c).n_cells == 20 * 8 * 6


def test_what_cannot_be_represented_is_refused(tmp_path):
    root = write_case(tmp_path)
    path = os.path.join(root, "system", "blockMeshDict")
    text = open(path).read()
    open(path, "w").write(text.replace("simpleGrading (1 1 1)", "simpleGrading (1 2 1)"))
    with pytest.raises(foamcase.FoamCaseError, match="graded"):
        foamcase.read_case(root)
    open(path, "w").write(text)
    tpath = os.path.join(root, "0", "T")
    ttext = open(tpath).read()
    open(tpath, "w").write(ttext.replace("type zeroGradient;", "type codedFixedValue;"))
    with pytest.raises(foamcase.FoamCaseError, match="codedFixedValue"):
        foamcase.read_case(root)


@pytest.mark.parametrize("rel,old,new,match", [
    ("system/fvSolution", "nOuterCorrectors 0;", "nOuterCorrectors 1;", "nOuterCorrectors"),
    ("system/fvSchemes", "default Euler;", "default localEuler;", "ddtSchemes"),
    ("system/fvSchemes", "default Euler;", "default CrankNicolson 1.2;", "ddtSchemes"),
    ("system/fvSchemes", "default Euler;", "default Euler; T CrankNicolson 0.9;", "default"),      # thermoScheme.H:53-64 reads psi from `default`
    ("system/fvSchemes", "laplacian(kappa,T) Gauss harmonic corrected;", "", "laplacian"),
    ("system/fvSchemes", "interpolationSchemes { default linear; }", "interpolationSchemes { default cubic; }", "interpolationSchemes"),
])
def test_schemes_and_flow_coupling_the_kernels_do_not_implement_are_refused(tmp_path, rel, old, new, match):
    """a marangoni-style case (nOuterCorrectors 1, tutorials/verificationAndValidation/marangoni/system/fvSolution:48) or
    a non-Euler / non-harmonic selection must raise instead of being run as Euler + harmonic pure conduction."""
    root = write_case(tmp_path)
    path = os.path.join(root, rel)
    text = open(path).read()
    assert old in text
    open(path, "w").write(text.replace(old, new))
    with pytest.raises(foamcase.FoamCaseError, match=match):
        foamcase.read_case(root)


@needs_reference
def test_shipped_marangoni_case_is_refused():
    root = os.path.join(TUT, "verificationAndValidation", "marangoni")
    with pytest.raises(foamcase.FoamCaseError):           # (its one-cell-thick `empty` patch is refused before the controls are read)
        foamcase.read_case(root)
    with pytest.raises(foamcase.FoamCaseError, match="nOuterCorrectors"):
        foamcase._read_controls(root)


def _normalised(case):
    c = {k: v for k, v in case.items()}
    c["beams"] = [dict(b, isoValue=b["isoValue"] if b["transient"] else 0.0) for b in case["beams"]]   # unused unless transient
    return c


@needs_reference
def test_shipped_amb2018_equals_the_hand_restated_case():
    """tutorials/AMB2018-02-B as shipped == cases.amb2018_02_b(): mesh, patches, material, beam, scan path, every control."""
    assert _normalised(foamcase.read_case(os.path.join(TUT, "AMB2018-02-B"))) == _normalised(cases.amb2018_02_b())


@needs_reference
def test_shipped_multibeam_and_multilayer_agree_with_the_restated_pieces():
    mb = foamcase.read_case(os.path.join(TUT, "multiBeam"))
    ref = cases.multi_beam((150, 25, 15))
    shape = lambda b: {k: v for k, v in b.items() if k != "path"}                                      # noqa: E731
    for got, want in zip(_normalised(mb)["beams"], _normalised(ref)["beams"][:2]):   # beam1 / beam2; the tracks are synthetic in cases.py
        assert shape(got) == shape(want)
    assert mb["beams"][0]["path"][1][4:] == [179.2, 0.8] and mb["controls"]["maxDi"] == 10.0
    ml = foamcase.read_case(os.path.join(TUT, "multiLayerPBF"))
    want = cases.multi_layer_pbf((150, 25, 30))
    assert ml["patches"] == want["patches"] and ml["material"] == want["material"] and ml["powderZMin"] == want["powderZMin"]
    assert ml["controls"]["Tmax"] == 3300.0 and ml["controls"]["maxDi"] == want["controls"]["maxDi"]
    assert {k: v for k, v in ml["beams"][0].items() if k != "path"} == {k: v for k, v in want["beams"][0].items() if k != "path"}
    # the scan path comes from constant/createScanPathDict through b200_create_scan_path: three bidirectional 2 mm lines
    assert [r[0] for r in ml["beams"][0]["path"]] == [1, 0, 1, 0, 1, 0] and ml["beams"][0]["path"][1][4:] == [195.0, 0.8]


@pytest.mark.gpu
def test_case_read_from_a_directory_runs_like_the_oracle(ctx, oracle, tmp_path):
    import helpers
    import oracle_lib as ol
    from additivefoam_b200 import lib
    root = write_case(tmp_path, powder=True, second_beam=True)
    g = lib.Case.from_case_dir(ctx, root)
    o = ol.OracleCase(foamcase.read_case(root), lib=oracle)
    for s in range(40):
        rg, ro = g.step(), o.step()
        assert rg.nThermoCorr == ro.nThermoCorr and abs(rg.lastSolveIterations - ro.lastSolveIterations) <= 1, f"step {s}"
        assert rg.deltaT == pytest.approx(ro.deltaT, rel=1e-11)
    assert helpers.rel_linf(g.field(abi.FIELD_T), o.field(abi.FIELD_T)) < 1e-9
    assert float(np.max(np.abs(g.field(abi.FIELD_ALPHA_SOLID) - o.field(abi.FIELD_ALPHA_SOLID)))) < 1e-9
    assert g.field(abi.FIELD_T).max() > 1000.0
    g.close(); o.close()


# ---------------------------------------------------------------------------------------------------------------------------
# restart I/O: the time directory additiveFoam writes and reads back -- T and alpha.powder as [UPSTREAM]
# volScalarField files, ascii or binary (createFields.H:2-29; controlDict writeFormat); alpha.solid is rebuilt from T (:31-44,74-78)
def test_field_files_round_trip_in_both_write_formats(tmp_path):
    from additivefoam_b200 import foamcase
    case = cases.amb2018_02_b(n=(7, 5, 3))
    rng = np.random.default_rng(5)
    T = rng.uniform(300.0, 3200.0, 105)
    powder = (rng.uniform(0, 1, 105) > 0.6).astype(np.float64)
    for binary in (True, False):
        pT, pA = foamcase.write_time_directory(str(tmp_path), "0.0005", case, T, powder, binary=binary, precision=17)
        head = open(pT, "rb").read(700).decode(errors="replace")
        assert "class       volScalarField;" in head and ("format      binary;" if binary else "format      ascii;") in head
        gotT, patches = foamcase.read_field(pT, 105)
        gotA, _ = foamcase.read_field(pA, 105)
        assert np.array_equal(gotT, T) and np.array_equal(gotA, powder)
        assert [p["type"] for p in patches.values()] == ["zeroGradient", "mixedTemperature", "zeroGradient"]
        assert patches["patch1"]["Tinf"] == 300.0 and float(patches["patch1"]["h"]) == 10.0
    # default ascii precision is the case's writePrecision: values come back to 8 significant digits
    pT, _ = foamcase.write_time_directory(str(tmp_path), "0.001", case, T, powder, binary=False, precision=8)
    assert np.allclose(foamcase.read_field(pT, 105)[0], T, rtol=1e-7, atol=0)
    # a uniform field is written as `uniform`, and read back at the mesh size
    pT, _ = foamcase.write_time_directory(str(tmp_path), "0", case, np.full(105, 300.0), np.zeros(105))
    assert b"internalField   uniform 300;" in open(pT, "rb").read()
    assert np.array_equal(foamcase.read_field(pT, 105)[0], np.full(105, 300.0))
    with pytest.raises(foamcase.FoamCaseError):
        foamcase.read_field(_truncate(tmp_path), 24)


def _truncate(tmp_path):
    """A binary field file cut short: must be refused, not zero-filled."""
    from additivefoam_b200 import foamcase
    case = cases.amb2018_02_b(n=(4, 3, 2))
    p, _ = foamcase.write_time_directory(str(tmp_path), "9", case, np.arange(24.0) + 1.5, np.zeros(24), binary=True)
    data = open(p, "rb").read()
    i = data.index(b"List<scalar>")
    open(p, "wb").write(data[:i + 60])
    return p


def test_restart_rebuilds_alpha_solid_from_temperature(oracle):
    """Oracle side of the restart: stop after 12 steps, write the time directory, start a fresh case from it (alpha.solid from T through
    the thermoPath table, createFields.H:74-78) and continue -- the continued run must equal the uninterrupted one in T; alpha.solid
    differs from the run's own only where the solidification history made it differ from the equilibrium table (it is not a file)."""
    import tempfile
    import oracle_lib as ol
    case = cases.amb2018_02_b(n=(40, 10, 6), explicitSolve=0, maxDi=10.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC, writeInterval=0.0)
    a = ol.OracleCase(case, lib=oracle)
    for _ in range(12):
        ra = a.step()
    with tempfile.TemporaryDirectory() as d:
        pT, pA = foamcase.write_time_directory(d, "%g" % ra.time, case, a.field(abi.FIELD_T), a.field(abi.FIELD_ALPHA_POWDER), binary=True)
        T0, _ = foamcase.read_field(pT, a.n)
        P0, _ = foamcase.read_field(pA, a.n)
    assert np.array_equal(T0, a.field(abi.FIELD_T))
    b = ol.OracleCase(case, lib=oracle)
    b.restart(T0, P0)
    alpha = b.field(abi.FIELD_ALPHA_SOLID)
    assert alpha.min() >= 0.0 and alpha.max() <= 1.0
    table = np.clip(np.interp(T0, case["material"]["thermoT"], case["material"]["thermoAlpha"]), 0.0, 1.0)
    assert np.allclose(alpha, table, rtol=0, atol=1e-14)
    assert np.array_equal(b.field(abi.FIELD_T_OLD), T0)


@pytest.mark.gpu
def test_gpu_restart_from_a_written_time_directory(ctx, oracle, tmp_path):
    """GPU side: a run is stopped in the mushy-zone phase, its time directory written in the tutorials' binary writeFormat, read back,
    and a fresh resident case restarted from it (b200_case_set_field + b200_case_reset_alpha_solid); the oracle does the same.  The
    rebuilt alpha.solid is bit-equal (pure table arithmetic), and six further steps stay within the north-star bound."""
    import oracle_lib as ol
    from additivefoam_b200 import lib
    case = cases.amb2018_02_b(n=(60, 12, 8), explicitSolve=0, maxDi=10.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC, writeInterval=0.0)
    g = lib.Case(ctx, case)
    for _ in range(30):
        rep = g.step()
    T = g.field(abi.FIELD_T)
    assert T.max() > 1410.0
    pT, pA = foamcase.write_time_directory(str(tmp_path), "%g" % rep.time, case, T, g.field(abi.FIELD_ALPHA_POWDER), binary=True)
    g.close()
    T0, _ = foamcase.read_field(pT, T.size)
    P0, _ = foamcase.read_field(pA, T.size)
    assert np.array_equal(T0, T)
    g2, o2 = lib.Case(ctx, case), ol.OracleCase(case, lib=oracle)
    g2.restart(T0, P0); o2.restart(T0, P0)
    assert np.array_equal(g2.field(abi.FIELD_ALPHA_SOLID), o2.field(abi.FIELD_ALPHA_SOLID))
    assert np.array_equal(g2.field(abi.FIELD_KAPPA), o2.field(abi.FIELD_KAPPA))
    for s in range(6):
        rg, ro = g2.step(), o2.step()
        assert rg.nThermoCorr == ro.nThermoCorr and abs(rg.lastSolveIterations - ro.lastSolveIterations) <= 1
        eT = float(np.max(np.abs(g2.field(abi.FIELD_T) - o2.field(abi.FIELD_T))) / np.max(np.abs(o2.field(abi.FIELD_T))))
        eA = float(np.max(np.abs(g2.field(abi.FIELD_ALPHA_SOLID) - o2.field(abi.FIELD_ALPHA_SOLID))))
        assert eT < 1e-9 and eA < 1e-9, (s, eT, eA)
    g2.close(); o2.close()


def test_crank_nicolson_ddt_scheme_and_its_coefficient_are_read(tmp_path):
    """ddtSchemes default CrankNicolson psi (thermoScheme.H:51-65): the scheme and its off-centring coefficient; none given means 1
    ([UPSTREAM] CrankNicolsonDdtScheme)."""
    for entry, psi in (("default CrankNicolson 0.9;", 0.9), ("default CrankNicolson;", 1.0), ("default CrankNicolson 0;", 0.0)):
        root = write_case(tmp_path / ("cn%g" % psi))
        path = os.path.join(root, "system", "fvSchemes")
        text = open(path).read()
        open(path, "w").write(text.replace("default Euler;", entry))
        ctl = foamcase.read_case(root)["controls"]
        assert ctl["ddtScheme"] == abi.DDT_CRANK_NICOLSON and ctl["ocCoeff"] == psi


def test_backward_ddt_scheme_is_read(tmp_path):
    """ddtSchemes default backward (the reference accepts it, thermoScheme.H:4-15,43-49) reaches the case description."""
    root = write_case(tmp_path)
    path = os.path.join(root, "system", "fvSchemes")
    text = open(path).read()
    open(path, "w").write(text.replace("default Euler;", "default backward;"))
    assert foamcase.read_case(root)["controls"]["ddtScheme"] == abi.DDT_BACKWARD
    assert foamcase.read_case(write_case(tmp_path / "e"))["controls"]["ddtScheme"] == abi.DDT_EULER
