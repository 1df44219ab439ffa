"""Read an AdditiveFOAM case directory -- the OpenFOAM dictionaries a user of the reference already has -- into the case
description ``lib.Case`` takes (the ``b200_case_desc`` of include/additivefoam_b200.h).

What is read, and where the reference reads it (paths relative to the AdditiveFOAM tree, SOLVER = applications/solvers/additiveFoam):

  system/blockMeshDict            one uniform hex block: vertices, ``blocks (hex (...) (nx ny nz) simpleGrading (1 1 1))``, the patch -> block-side map
  0/T                             internalField, boundaryField types (zeroGradient | fixedValue | mixedTemperature: h, emissivity, Tinf)
                                  -- SOLVER/createFields.H:1-14, derivedFvPatchFields/mixedTemperature
  constant/transportProperties    solid/liquid/powder kappa, Cp (Polynomial<3>), rho, Lf            -- SOLVER/readTransportProperties.H:13-73
  constant/thermoPath             (T alpha.solid) rows                                               -- SOLVER/createFields.H:46-57
  constant/heatSourceDict         sources, per beam: pathName, absorptionModel + Coeffs, heatSourceModel + Coeffs
                                  -- movingHeatSource/heatSourceModels/heatSourceModel/heatSourceModel.C:89-118, movingBeam/movingBeam.C:58-63,91
  constant/<pathName>             the scan path (movingBeam.C:89-148); if it is missing and constant/createScanPathDict exists, the file
                                  applications/utilities/createScanPath would write is generated (b200_create_scan_path)
  system/controlDict              startTime, endTime, deltaT, adjustTimeStep, maxCo, maxDi, maxAlphaCo, maxDeltaT, write interval
  system/fvSolution               solvers."T.*" (solver, preconditioner, tolerance, relTol, minIter, maxIter); PIMPLE: nThermoCorrectors,
                                  thermoTolerance, Tmax, explicitSolve                                 -- SOLVER/solutionControls.H:2-33
  system/setFieldsDict            (optional) the powder layer: a boxToCell region setting alpha.powder 1

Only what the thermal path uses is interpreted; anything the resident stepper cannot represent (several blocks, graded or
non-axis-aligned blocks, other boundary types, non-uniform initial fields) raises ``FoamCaseError`` instead of being approximated.
The dictionary grammar handled is the subset those files use: comments, ``key value...;``, sub-dictionaries, (nested) lists,
``$macro`` substitution, dimension sets, ``uniform``/quoted keys; ``#...`` directive lines are skipped.
"""
import os
import re

from . import abi

GREAT = 1e15          # [UPSTREAM] doubleScalarGreat


class FoamCaseError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ dictionary parser
_TOKEN = re.compile(r'"[^"]*"|[{}()\[\];]|[^\s{}()\[\];"]+')


def _tokens(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    out = []
    for line in text.splitlines():
        line = line.split("//", 1)[0]
        if line.lstrip().startswith("#"):
            continue                                  # #include / #ifeq ... : nothing the thermal path needs
        out += _TOKEN.findall(line)
    return out


def _atom(tok):
    if tok.startswith('"'):
        return tok[1:-1]
    try:
        return int(tok)
    except ValueError:
        pass
    try:
        return float(tok)
    except ValueError:
        return tok


class _Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def next(self):
        tok = self.peek()
        if tok is None:
            raise FoamCaseError("unexpected end of dictionary")
        self.i += 1
        return tok

    def parse_dict(self, closing):
        d = {}
        while True:
            tok = self.peek()
            if tok is None:
                if closing is None:
                    return d
                raise FoamCaseError("missing '}'")
            if tok == closing:
                self.next()
                return d
            key = _atom(self.next())
            if self.peek() == "{":
                self.next()
                d[key] = self.parse_dict("}")
                continue
            vals = []
            while self.peek() != ";":
                vals.append(self.parse_value())
            self.next()
            if isinstance(key, str) and key.startswith("$") and not vals:
                d.setdefault("$include", []).append(key[1:])      # `$p_rgh;` : copy of another sub-dictionary
            else:
                d[key] = vals[0] if len(vals) == 1 else vals

    def parse_value(self):
        tok = self.next()
        if tok == "(":
            return self.parse_list(")")
        if tok == "[":
            return ("dimensions", self.parse_list("]"))
        if tok == "{":
            return self.parse_dict("}")
        if tok in (")", "]", "}"):
            raise FoamCaseError(f"unexpected '{tok}'")
        return _atom(tok)

    def parse_list(self, closing):
        out = []
        while self.peek() != closing:
            if self.peek() is None:
                raise FoamCaseError(f"missing '{closing}'")
            out.append(self.parse_value())
        self.next()
        return out


def _substitute(node, scopes):
    """``$name`` -> the value of `name` in the nearest enclosing dictionary."""
    if isinstance(node, dict):
        for k in list(node):
            node[k] = _substitute(node[k], scopes + [node])
        return node
    if isinstance(node, list):
        return [_substitute(v, scopes) for v in node]
    if isinstance(node, tuple):
        return tuple(_substitute(v, scopes) for v in node)
    if isinstance(node, str) and node.startswith("$"):
        name = node[1:].strip("{}")
        for s in reversed(scopes):
            if name in s:
                return s[name]
        raise FoamCaseError(f"undefined macro {node}")
    return node


def parse_dictionary(text):
    """An OpenFOAM dictionary file as nested dict / list / number / word values (FoamFile header included)."""
    return _substitute(_Parser(_tokens(text)).parse_dict(None), [])


def parse_list_file(text):
    """A file that is one bare list (constant/thermoPath)."""
    p = _Parser(_tokens(text))
    if p.next() != "(":
        raise FoamCaseError("expected a list")
    return p.parse_list(")")


def _read(case_dir, *rel):
    path = os.path.join(case_dir, *rel)
    if not os.path.exists(path):
        raise FoamCaseError(f"{path} is missing")
    with open(path) as f:
        return f.read()


def _scalar(v, what):
    """`rho [1 -3 0 0 0 0 0] 7569.92;` or `h 10.0;` or `Tinf uniform 300;`."""
    if isinstance(v, list):
        v = [x for x in v if not (isinstance(x, tuple) and x and x[0] == "dimensions") and x != "uniform" and x != what]
        if len(v) != 1:
            raise FoamCaseError(f"cannot read a scalar from {what}")
        v = v[0]
    if isinstance(v, bool) or not isinstance(v, (int, float)):
        raise FoamCaseError(f"{what}: expected a number, found {v!r}")
    return float(v)


def _switch(v, what):
    if isinstance(v, str) and v.lower() in ("true", "yes", "on", "y", "t"):
        return 1
    if isinstance(v, str) and v.lower() in ("false", "no", "off", "n", "f", "none"):
        return 0
    raise FoamCaseError(f"{what}: expected a Switch, found {v!r}")


# ------------------------------------------------------------------------------------------------ the files
def _read_mesh(case_dir):
    d = parse_dictionary(_read(case_dir, "system", "blockMeshDict"))
    scale = float(d.get("convertToMeters", d.get("scale", 1.0)))
    verts = [[scale * float(c) for c in v] for v in d["vertices"]]
    blocks = d["blocks"]
    if len(blocks) != 5 or blocks[0] != "hex" or blocks[3] != "simpleGrading":
        raise FoamCaseError("the resident stepper takes exactly one `hex (...) (nx ny nz) simpleGrading (...)` block")
    if any(float(g) != 1.0 for g in blocks[4]):
        raise FoamCaseError("graded blocks are not uniform hex meshes")
    ids, n = [int(v) for v in blocks[1]], [int(v) for v in blocks[2]]
    corner = [verts[i] for i in ids]
    lo, hi = [min(c[q] for c in corner) for q in range(3)], [max(c[q] for c in corner) for q in range(3)]
    # blockMesh vertex order: 0-3 the z-min face counter-clockwise from (xmin, ymin), 4-7 above them
    want = [[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]],
            [lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]]]
    if corner != want:
        raise FoamCaseError("the block is not an axis-aligned box in blockMesh's canonical vertex order")
    patches = []                       # (name, [sides]) in blockMeshDict order: that is OpenFOAM's patch order
    entries = d["boundary"]
    for name, spec in zip(entries[0::2], entries[1::2]):
        sides = []
        for face in spec["faces"]:
            pts = [verts[int(i)] for i in face]
            side = None
            for q, (a, b) in enumerate(((abi.XMIN, abi.XMAX), (abi.YMIN, abi.YMAX), (abi.ZMIN, abi.ZMAX))):
                if all(p[q] == lo[q] for p in pts):
                    side = a
                elif all(p[q] == hi[q] for p in pts):
                    side = b
            if side is None:
                raise FoamCaseError(f"face {face} of patch {name} is not a side of the block")
            sides.append(side)
        patches.append((name, sides))
    if sorted(s for _, ss in patches for s in ss) != sorted([abi.XMIN, abi.XMAX, abi.YMIN, abi.YMAX, abi.ZMIN, abi.ZMAX]):
        raise FoamCaseError("the patches must cover each side of the block exactly once")
    return dict(n=n, min=lo, max=hi), patches


def _read_T(case_dir, mesh_patches, start_dir="0"):
    d = parse_dictionary(_read(case_dir, start_dir, "T"))
    internal = d["internalField"]
    if not (isinstance(internal, list) and internal[0] == "uniform"):
        raise FoamCaseError("only a uniform initial temperature is read; upload other fields with Case.set_field")
    bf, out = d["boundaryField"], []
    for name, sides in mesh_patches:
        spec = bf.get(name)
        if spec is None:                                   # regular-expression keys
            spec = next((v for k, v in bf.items() if isinstance(k, str) and re.fullmatch(k, name)), None)
        if spec is None:
            raise FoamCaseError(f"0/T has no boundaryField entry for patch {name}")
        t = spec["type"]
        if t == "zeroGradient":
            out.append(dict(type=abi.BC_ZERO_GRADIENT, sides=sides, value=0.0, h=0.0, emissivity=0.0, Tinf=0.0))
        elif t == "fixedValue":
            out.append(dict(type=abi.BC_FIXED_VALUE, sides=sides, value=_scalar(spec["value"], "value"), h=0.0, emissivity=0.0, Tinf=0.0))
        elif t == "mixedTemperature":
            out.append(dict(type=abi.BC_MIXED_TEMPERATURE, sides=sides, value=0.0, h=_scalar(spec["h"], "h"),
                            emissivity=_scalar(spec["emissivity"], "emissivity"), Tinf=_scalar(spec["Tinf"], "Tinf")))
        else:
            raise FoamCaseError(f"boundary type {t} of patch {name} is not one the thermal path implements")
    return float(internal[1]), out


def _read_material(case_dir):
    d = parse_dictionary(_read(case_dir, "constant", "transportProperties"))
    poly = lambda phase, key: [float(v) for v in d[phase][key]]       # noqa: E731
    rows = parse_list_file(_read(case_dir, "constant", "thermoPath"))
    flat = [float(v) for v in rows]
    if len(flat) < 4 or len(flat) % 2:
        raise FoamCaseError("thermoPath needs at least two (T alpha.solid) rows")
    return dict(kappa=[poly(p, "kappa") for p in ("solid", "liquid", "powder")], Cp=[poly(p, "Cp") for p in ("solid", "liquid", "powder")],
                rho=_scalar(d["rho"], "rho"), Lf=_scalar(d["Lf"], "Lf"), thermoT=flat[0::2], thermoAlpha=flat[1::2])


_SHAPES = {"superGaussian": (abi.SUPER_GAUSSIAN, ("k",)), "modifiedSuperGaussian": (abi.MODIFIED_SUPER_GAUSSIAN, ("k", "m")),
           "projectedGaussian": (abi.PROJECTED_GAUSSIAN, ("A", "B"))}


def _scan_path(case_dir, name):
    from . import lib
    path = os.path.join(case_dir, "constant", name)
    if os.path.exists(path):
        with open(path) as f:
            return lib.scan_path_parse(f.read()).tolist()
    dict_path = os.path.join(case_dir, "constant", "createScanPathDict")
    m = re.fullmatch(r"(.*)_(\d+)", name)
    if not os.path.exists(dict_path):
        raise FoamCaseError(f"{path} is missing (and there is no constant/createScanPathDict to make it from)")
    # the file createScanPath would write: constant/scanPath_<i>, rotation i*angle (createScanPath.C:74-101)
    g = parse_dictionary(_read(case_dir, "constant", "createScanPathDict"))
    index = int(m.group(2)) if m else 0
    texts = lib.create_scan_path(g["minPoint"], g["maxPoint"], float(g["hatch"]), float(g["angle"]), max(int(g["nRotations"]), index + 1),
                                 float(g["power"]), float(g["speed"]), float(g["dwellTime"]),
                                 _switch(g["biDirection"], "biDirection") if "biDirection" in g else True)
    return lib.scan_path_parse(texts[index]).tolist()


def _read_beams(case_dir):
    d = parse_dictionary(_read(case_dir, "constant", "heatSourceDict"))
    names = d["sources"]
    beams = []
    for name in names if isinstance(names, list) else [names]:
        s = d[name]
        shape = s["heatSourceModel"]
        if shape.startswith("b200"):                       # the GPU plug-in types carry the same coefficients
            shape = shape[4].lower() + shape[5:]
        if shape not in _SHAPES:
            raise FoamCaseError(f"unknown heatSourceModel {s['heatSourceModel']}")
        model, keys = _SHAPES[shape]
        c = s.get(s["heatSourceModel"] + "Coeffs", s)
        b = dict(heatSourceModel=model, dimensions=[float(v) for v in c["dimensions"]],
                 nPoints=[int(v) for v in c["nPoints"]] if "nPoints" in c else [1, 1, 1],       # heatSourceModel.C:112-117
                 transient=_switch(c["transient"], "transient") if "transient" in c else 0,
                 isoValue=float(c.get("isoValue", GREAT)), k=0.0, m=0.0, A=0.0, B=0.0,
                 absorptionModel=abi.ABSORPTION_CONSTANT, eta=0.0, kellyGeometry=abi.KELLY_CONE, eta0=0.0, etaMin=0.0,
                 deltaT=float(s["deltaT"]) if "deltaT" in s else 0.0,                              # movingBeam.C:62 (0 = GREAT)
                 hitPathIntervals=_switch(s["hitPathIntervals"], "hitPathIntervals") if "hitPathIntervals" in s else 1,
                 path=_scan_path(case_dir, s["pathName"]))
        for key in keys:
            b[key] = float(c[key])
        absorption = s["absorptionModel"]
        ac = s.get(absorption + "Coeffs", s)
        if absorption == "constant":
            b["eta"] = float(ac["eta"])
        elif absorption == "Kelly":
            b["absorptionModel"] = abi.ABSORPTION_KELLY
            geometry = ac["geometry"]
            if geometry not in ("cone", "cylinder"):
                raise FoamCaseError(f"Kelly geometry {geometry}")
            b["kellyGeometry"] = abi.KELLY_CONE if geometry == "cone" else abi.KELLY_CYLINDER
            b["eta0"], b["etaMin"] = float(ac["eta0"]), float(ac["etaMin"])
        else:
            raise FoamCaseError(f"unknown absorptionModel {absorption}")
        beams.append(b)
    return beams


_SOLVERS = {"PCG": abi.SOLVER_PCG, "bPCG": abi.SOLVER_PCG, "PBiCGStab": abi.SOLVER_PBICGSTAB, "bPBiCGStab": abi.SOLVER_PBICGSTAB}
_PRECONDS = {"DIC": abi.PRECOND_DIC, "DILU": abi.PRECOND_DILU, "diagonal": abi.PRECOND_DIAGONAL, "none": abi.PRECOND_NONE}


def _read_controls(case_dir):
    c = parse_dictionary(_read(case_dir, "system", "controlDict"))
    f = parse_dictionary(_read(case_dir, "system", "fvSolution"))
    solvers = f["solvers"]
    ts = solvers.get("T")
    if ts is None:
        ts = next((v for k, v in solvers.items() if isinstance(k, str) and isinstance(v, dict) and re.fullmatch(k, "T")), None)
    if ts is None:
        raise FoamCaseError('fvSolution has no solver entry matching "T"')
    if ts["solver"] not in _SOLVERS:
        raise FoamCaseError(f"solver {ts['solver']} is not one the library implements (PCG, PBiCGStab)")
    precond = ts.get("preconditioner", "none")
    if precond not in _PRECONDS:
        raise FoamCaseError(f"preconditioner {precond}")
    pimple = f.get("PIMPLE", {})
    # pU (SOLVER/pU/*, additiveFoam.C:78-88) is not built: a case that asks for the momentum/pressure loop would silently run as
    # pure conduction, so it is refused (tutorials/verificationAndValidation/marangoni has nOuterCorrectors 1)
    if int(pimple.get("nOuterCorrectors", 0)) > 0:
        raise FoamCaseError("PIMPLE nOuterCorrectors > 0 asks for the pU loop and fvm::div(phi,T) with phi != 0, which the library "
                            "does not implement (thermal path only)")
    ddt_scheme, oc_coeff = _check_schemes(case_dir)
    adjustable = c.get("writeControl") == "adjustableRunTime"
    return dict(startTime=float(c.get("startTime", 0.0)), endTime=float(c["endTime"]), deltaT=float(c["deltaT"]),
                adjustTimeStep=_switch(c["adjustTimeStep"], "adjustTimeStep") if "adjustTimeStep" in c else 0,
                maxCo=float(c.get("maxCo", 1.0)), maxDi=float(c.get("maxDi", 10.0)), maxAlphaCo=float(c.get("maxAlphaCo", 1.0)),
                maxDeltaT=float(c.get("maxDeltaT", GREAT)), writeInterval=float(c["writeInterval"]) if adjustable else 0.0,
                nThermoCorrectors=int(pimple.get("nThermoCorrectors", 20)),                          # solutionControls.H:2-6
                thermoTolerance=float(pimple.get("thermoTolerance", 1e-6)), Tmax=float(pimple.get("Tmax", 0.0)),
                explicitSolve=_switch(pimple["explicitSolve"], "explicitSolve") if "explicitSolve" in pimple else 0,
                solver=_SOLVERS[ts["solver"]], preconditioner=_PRECONDS[precond], tolerance=float(ts.get("tolerance", 1e-6)),
                relTol=float(ts.get("relTol", 0.0)), minIter=int(ts.get("minIter", 0)), maxIter=int(ts.get("maxIter", 1000)),
                ddtScheme=ddt_scheme, ocCoeff=oc_coeff)


def _scheme(d, section, *keys):
    """The entry of fvSchemes `section` for the first of `keys` present, else its `default` (a list of words or one word)."""
    sec = d.get(section, {})
    for k in keys + ("default",):
        if k in sec:
            v = sec[k]
            return " ".join(str(w) for w in v) if isinstance(v, list) else str(v)
    return None


def _check_schemes(case_dir):
    """The kernels hard-code the schemes every shipped case selects (tutorials/*/system/fvSchemes:18-47): Euler, backward or
    CrankNicolson ddt (thermoScheme.H:2-15,41-65), Gauss harmonic for laplacian(kappa,T), linear interpolation (setDeltaT.H:21-27).
    Anything else is refused instead of being approximated.  Returns (ddtScheme, ocCoeff)."""
    path = os.path.join(case_dir, "system", "fvSchemes")
    if not os.path.exists(path):
        raise FoamCaseError("system/fvSchemes is missing")
    d = parse_dictionary(_read(case_dir, "system", "fvSchemes"))
    # thermoScheme.H:2 looks the scheme up by the field's name: mesh.schemes().ddt("T") -> the entry `T` if there is one, else `default`
    ddt = _scheme(d, "ddtSchemes", "T")
    words = (ddt or "").split()
    if not words or words[0] not in ("Euler", "backward", "CrankNicolson"):
        raise FoamCaseError(f"ddtSchemes {ddt!r}: thermoScheme.H:4-15 accepts Euler, CrankNicolson or backward")
    oc_coeff = 1.0
    if words[0] == "CrankNicolson":
        # thermoScheme.H:53-64 builds the scheme a second time from mesh.schemes().ddt("CrankNicolson"), i.e. from `default`: the
        # coefficient of the Sp terms is the default entry's, the one of fvm::ddt(T) the entry of T -- only one and the same is accepted
        if _scheme(d, "ddtSchemes") != ddt:
            raise FoamCaseError("ddtSchemes: with CrankNicolson for T the `default` entry must be the same scheme and coefficient "
                                "(thermoScheme.H:53-64 reads the coefficient from the default entry)")
        if len(words) > 2:
            raise FoamCaseError(f"ddtSchemes {ddt!r}: a constant off-centring coefficient is implemented, not a function of time")
        oc_coeff = float(words[1]) if len(words) == 2 else 1.0       # [UPSTREAM] CrankNicolsonDdtScheme: psi = 1 when none is given
        if not 0.0 <= oc_coeff <= 1.0:
            raise FoamCaseError(f"ddtSchemes {ddt!r}: the off-centring coefficient must lie in [0, 1]")
    ddt = words[0]
    # `laplacian(kappa,T)` is one word to OpenFOAM but a word plus a list to this module's grammar: look for it in the text itself
    text = re.sub(r"//[^\n]*|/\*.*?\*/", "", _read(case_dir, "system", "fvSchemes"), flags=re.S)
    m = re.search(r"laplacian\(\s*kappa\s*,\s*T\s*\)\s+([^;]+);", text)
    lap = m.group(1).strip() if m else _scheme(d, "laplacianSchemes")
    if lap is None or lap.split()[:2] != ["Gauss", "harmonic"]:
        raise FoamCaseError(f"laplacianSchemes laplacian(kappa,T) {lap!r}: the assembly kernels implement Gauss harmonic")
    interp = _scheme(d, "interpolationSchemes")
    if interp != "linear":
        raise FoamCaseError(f"interpolationSchemes default {interp!r}: the diffusion-number kernel implements linear interpolation")
    return {"Euler": abi.DDT_EULER, "backward": abi.DDT_BACKWARD, "CrankNicolson": abi.DDT_CRANK_NICOLSON}[ddt], oc_coeff


def _read_powder(case_dir, mesh):
    path = os.path.join(case_dir, "system", "setFieldsDict")
    if not os.path.exists(path):
        return 1e300
    d = parse_dictionary(_read(case_dir, "system", "setFieldsDict"))
    z_min = 1e300
    regions = d.get("regions", [])
    for kind, spec in zip(regions[0::2], regions[1::2]):
        values = spec.get("fieldValues", [])
        hits = [values[i + 2] for i in range(0, len(values) - 2, 3) if values[i] == "volScalarFieldValue" and values[i + 1] == "alpha.powder"]
        if not hits or float(hits[0]) == 0.0:
            continue
        if kind != "boxToCell" or float(hits[0]) != 1.0:
            raise FoamCaseError("the powder layer must be a boxToCell region with alpha.powder 1")
        lo, hi = spec["box"]
        covers = all(float(lo[q]) <= mesh["min"][q] and float(hi[q]) >= mesh["max"][q] for q in (0, 1)) and float(hi[2]) >= mesh["max"][2]
        if not covers:
            raise FoamCaseError("the powder box must span the block in x and y and reach its top (alpha.powder = 1 above a plane)")
        z_min = float(lo[2])
    return z_min


def read_case(case_dir, start_dir="0"):
    """The case description of an AdditiveFOAM case directory (see the module docstring)."""
    mesh, mesh_patches = _read_mesh(case_dir)
    t_initial, patches = _read_T(case_dir, mesh_patches, start_dir)
    return dict(mesh=mesh, patches=patches, material=_read_material(case_dir), beams=_read_beams(case_dir),
                controls=_read_controls(case_dir), Tinitial=t_initial, powderZMin=_read_powder(case_dir, mesh))


# ------------------------------------------------------------------------------------------------ field files (restart I/O)
# What additiveFoam writes at a write time and reads back on restart: T and alpha.powder (SOLVER/createFields.H:2-29, AUTO_WRITE;
# alpha.solid is NO_READ / NO_WRITE and is rebuilt from T, :31-44, :74-78), as [UPSTREAM] volScalarField files in the writeFormat of
# system/controlDict (`ascii` or `binary`; the tutorials ship `writeFormat binary`).  The internal field is what the resident stepper
# exchanges; patch `value` entries are written from the caller's patch values where given and as the neighbouring cell otherwise.
_HEADER = """/*--------------------------------*- C++ -*----------------------------------*\\
  =========                 |
  \\\\      /  F ield         | OpenFOAM: The Open Source CFD Toolbox
   \\\\    /   O peration     | Website:  https://openfoam.org
    \\\\  /    A nd           | Version:  10
     \\\\/     M anipulation  |
\\*---------------------------------------------------------------------------*/
FoamFile
{
    format      %s;
    class       volScalarField;
    location    "%s";
    object      %s;
}
// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //

"""


def _list_bytes(values, binary, precision):
    import numpy as np
    v = np.ascontiguousarray(values, dtype=np.float64)
    if binary:
        return b"nonuniform List<scalar> \n%d\n(" % v.size + v.astype("<f8").tobytes() + b")"
    fmt = "%%.%dg" % precision
    return ("nonuniform List<scalar> \n%d\n(\n" % v.size + "\n".join(fmt % x for x in v) + "\n)\n").encode()


def write_field(case_dir, time_name, name, internal, dimensions, patches, binary=True, precision=8):
    """Write ``<case>/<time>/<name>`` as OpenFOAM does.  ``patches``: ordered {patch name: dict(type=..., entries...)}; an entry whose
    value is an array is written as a (nonuniform) List<scalar>, a float as ``uniform``; everything else verbatim."""
    import numpy as np
    os.makedirs(os.path.join(case_dir, time_name), exist_ok=True)
    out = [(_HEADER % ("binary" if binary else "ascii", time_name, name)).encode(),
           ("dimensions      [%s];\n\n" % " ".join(str(d) for d in dimensions)).encode(), b"internalField   "]
    internal = np.asarray(internal, dtype=np.float64)
    if internal.size and np.all(internal == internal.flat[0]):
        out.append(("uniform %.*g" % (precision if not binary else 17, float(internal.flat[0]))).encode())
    else:
        out.append(_list_bytes(internal, binary, precision))
    out.append(b";\n\nboundaryField\n{\n")
    for pname, entries in patches.items():
        out.append(("    %s\n    {\n" % pname).encode())
        for k, v in entries.items():
            if isinstance(v, (list, tuple)) or hasattr(v, "dtype"):
                out.append(("        %-15s " % k).encode() + _list_bytes(v, binary, precision) + b";\n")
            elif isinstance(v, float):
                out.append(("        %-15s uniform %.*g;\n" % (k, 17 if binary else precision, v)).encode())
            else:
                out.append(("        %-15s %s;\n" % (k, v)).encode())
        out.append(b"    }\n")
    out.append(b"}\n\n\n// ************************************************************************* //\n")
    path = os.path.join(case_dir, time_name, name)
    with open(path, "wb") as f:
        f.write(b"".join(out))
    return path


def _read_scalar_entry(data, pos, binary, n_expected=None):
    """Parse ``uniform x`` or ``nonuniform List<scalar> N ( ... )`` starting at byte offset ``pos`` (after the keyword); returns (value or array, end)."""
    import numpy as np
    m = re.compile(rb"\s*(uniform|nonuniform)\s*").match(data, pos)
    if not m:
        raise FoamCaseError("field entry is neither uniform nor nonuniform")
    if m.group(1) == b"uniform":
        e = data.index(b";", m.end())
        return float(data[m.end():e].decode()), e + 1
    m2 = re.compile(rb"List<scalar>\s*(\d+)\s*\(").match(data, m.end())
    if not m2:
        raise FoamCaseError("only List<scalar> fields are read")
    n = int(m2.group(1))
    if n_expected is not None and n != n_expected:
        raise FoamCaseError("field has %d values, the mesh %d cells" % (n, n_expected))
    if binary:
        raw = data[m2.end():m2.end() + 8 * n]
        if len(raw) != 8 * n or data[m2.end() + 8 * n:m2.end() + 8 * n + 1] != b")":
            raise FoamCaseError("truncated binary List<scalar>")
        return np.frombuffer(raw, dtype="<f8").copy(), data.index(b";", m2.end() + 8 * n) + 1
    e = data.index(b")", m2.end())
    vals = np.array(data[m2.end():e].split(), dtype=np.float64)
    if vals.size != n:
        raise FoamCaseError("List<scalar> announces %d values and holds %d" % (n, vals.size))
    return vals, data.index(b";", e) + 1


def read_field(path, n_cells=None):
    """The internal field of a volScalarField file (ascii or binary, uniform or nonuniform) -> numpy array (``n_cells`` values when the
    field is uniform and ``n_cells`` is given, else the scalar), plus {patch: {key: value}} with ``value`` entries parsed the same way."""
    import numpy as np
    data = open(path, "rb").read()
    head = re.search(rb"FoamFile\s*\{(.*?)\}", data, re.S)
    if not head:
        raise FoamCaseError("%s: no FoamFile header" % path)
    binary = re.search(rb"format\s+binary\s*;", head.group(1)) is not None
    if not re.search(rb"class\s+volScalarField\s*;", head.group(1)):
        raise FoamCaseError("%s: not a volScalarField" % path)
    m = re.compile(rb"internalField").search(data, head.end())
    if not m:
        raise FoamCaseError("%s: no internalField" % path)
    internal, pos = _read_scalar_entry(data, m.end(), binary, n_cells if n_cells else None)
    if np.isscalar(internal) and n_cells:
        internal = np.full(n_cells, internal)
    patches = {}
    b = re.compile(rb"boundaryField\s*\{").search(data, pos)
    pos = b.end() if b else len(data)
    pat = re.compile(rb"\s*([A-Za-z_][\w.]*)\s*\{")
    while b:
        pm = pat.match(data, pos)
        if not pm:
            break
        pos = pm.end()
        entries = {}
        while True:
            km = re.compile(rb"\s*(\}|[A-Za-z_]\w*)").match(data, pos)
            if not km:
                raise FoamCaseError("%s: malformed patch %s" % (path, pm.group(1).decode()))
            pos = km.end()
            if km.group(1) == b"}":
                break
            key = km.group(1).decode()
            if re.compile(rb"\s*(uniform|nonuniform)\b").match(data, pos):
                entries[key], pos = _read_scalar_entry(data, pos, binary)
            else:
                e = data.index(b";", pos)
                entries[key] = data[pos:e].decode().strip()
                pos = e + 1
        patches[pm.group(1).decode()] = entries
    return internal, patches


def write_time_directory(case_dir, time_name, case, T, alpha_powder, binary=True, precision=8):
    """What `runTime.write()` leaves for a restart: ``<time>/T`` with the boundary types of the case, and ``<time>/alpha.powder``."""
    names = case.get("patchNames") or ["patch%d" % i for i in range(len(case["patches"]))]
    tp, ap = {}, {}
    for name, p in zip(names, case["patches"]):
        if p["type"] == abi.BC_FIXED_VALUE:
            tp[name] = dict(type="fixedValue", value=float(p["value"]))
        elif p["type"] == abi.BC_MIXED_TEMPERATURE:
            tp[name] = dict(type="mixedTemperature", h=repr(float(p["h"])), emissivity=repr(float(p["emissivity"])), Tinf=float(p["Tinf"]))
        else:
            tp[name] = dict(type="zeroGradient")
        ap[name] = dict(type="zeroGradient")
    return (write_field(case_dir, time_name, "T", T, [0, 0, 0, 1, 0, 0, 0], tp, binary, precision),
            write_field(case_dir, time_name, "alpha.powder", alpha_powder, [0, 0, 0, 0, 0, 0, 0], ap, binary, precision))
