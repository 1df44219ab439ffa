"""Case descriptions: the dictionaries of an AdditiveFOAM case restated as a ``b200_case_desc``.

Values come from the reference tutorials (paths relative to the AdditiveFOAM tree):
  material   tutorials/AMB2018-02-B/constant/transportProperties:18-41, constant/thermoPath:2-3
  mesh       tutorials/AMB2018-02-B/system/blockMeshDict:16-74
  T BCs      tutorials/AMB2018-02-B/0/T:19-41
  controls   tutorials/AMB2018-02-B/system/controlDict:22-54, system/fvSolution:34-58
  beam       tutorials/AMB2018-02-B/constant/heatSourceDict:18-46, constant/scanPath:1-3
The synthetic configurations are the ones SURVEY.md section 8(d) defines for BASELINE.json.
"""
import ctypes as C
import copy

from . import abi

IN625 = dict(
    kappa=[[8.275, 0.01472, 0.0], [4.889, 0.014743, 0.0], [-0.07707, 0.00075, 0.0]],
    Cp=[[579.28, 0.0, 0.0], [750.65, 0.0, 0.0], [747.568, 0.0, 0.0]],
    rho=7569.92, Lf=2.1754e5,
    thermoT=[1410.0, 1620.0], thermoAlpha=[1.0, 0.0],
)

# tutorials/multiLayerPBF/constant/transportProperties:30-33 (powder block differs)
IN625_PBF = copy.deepcopy(IN625)


def _patch(ptype, sides, value=0.0, h=0.0, emissivity=0.0, Tinf=0.0):
    return dict(type=ptype, sides=list(sides), value=value, h=h, emissivity=emissivity, Tinf=Tinf)


def default_patches():
    """bottom zeroGradient, top mixedTemperature(h=10, eps=0.4, Tinf=300), sides zeroGradient."""
    return [
        _patch(abi.BC_ZERO_GRADIENT, [abi.ZMIN]),
        _patch(abi.BC_MIXED_TEMPERATURE, [abi.ZMAX], h=10.0, emissivity=0.4, Tinf=300.0),
        _patch(abi.BC_ZERO_GRADIENT, [abi.XMIN, abi.XMAX, abi.YMIN, abi.YMAX]),
    ]


def super_gaussian_beam(path, dims=(85e-6, 85e-6, 30e-6), nPoints=(10, 10, 10), k=2.0, eta=0.33):
    return dict(heatSourceModel=abi.SUPER_GAUSSIAN, dimensions=list(dims), nPoints=list(nPoints), transient=0,
                isoValue=0.0, k=k, m=0.0, A=0.0, B=0.0, absorptionModel=abi.ABSORPTION_CONSTANT, eta=eta,
                kellyGeometry=abi.KELLY_CONE, eta0=0.0, etaMin=0.0, deltaT=0.0, hitPathIntervals=1, path=path)


def modified_super_gaussian_beam(path, dims=(40e-6, 40e-6, 30e-6), nPoints=(10, 10, 10), k=7.95, m=2.72,
                                 transient=1, isoValue=1620.0, eta0=0.28, etaMin=0.35):
    """tutorials/multiBeam/constant/heatSourceDict:50-74 (beam2)."""
    return dict(heatSourceModel=abi.MODIFIED_SUPER_GAUSSIAN, dimensions=list(dims), nPoints=list(nPoints),
                transient=transient, isoValue=isoValue, k=k, m=m, A=0.0, B=0.0,
                absorptionModel=abi.ABSORPTION_KELLY, eta=0.0, kellyGeometry=abi.KELLY_CONE, eta0=eta0,
                etaMin=etaMin, deltaT=0.0, hitPathIntervals=1, path=path)


def projected_gaussian_beam(path, dims=(50e-6, 50e-6, 30e-6), nPoints=(10, 10, 10), A=2.0, B=1.0, transient=1,
                            isoValue=1620.0, eta=0.35):
    return dict(heatSourceModel=abi.PROJECTED_GAUSSIAN, dimensions=list(dims), nPoints=list(nPoints),
                transient=transient, isoValue=isoValue, k=0.0, m=0.0, A=A, B=B,
                absorptionModel=abi.ABSORPTION_CONSTANT, eta=eta, kellyGeometry=abi.KELLY_CONE, eta0=0.0,
                etaMin=0.0, deltaT=0.0, hitPathIntervals=1, path=path)


def default_controls(**kw):
    c = dict(startTime=0.0, endTime=0.004, deltaT=1e-7, adjustTimeStep=1, maxCo=0.5, maxDi=1.0, maxAlphaCo=1.0,
             maxDeltaT=1e15, writeInterval=0.001, nThermoCorrectors=20, thermoTolerance=1e-8, Tmax=0.0,
             explicitSolve=1, solver=abi.SOLVER_PBICGSTAB, preconditioner=abi.PRECOND_DILU, tolerance=1e-15,
             relTol=0.0, minIter=1, maxIter=20, ddtScheme=abi.DDT_EULER, ocCoeff=1.0)
    c.update(kw)
    return c


def amb2018_02_b(n=(150, 25, 15), **controls):
    """BASELINE configs[0]: tutorials/AMB2018-02-B as shipped (explicitSolve true, maxDi 1, PBiCGStab/DILU)."""
    return dict(
        mesh=dict(n=list(n), min=[-0.0005, -0.00025, -0.0003], max=[0.0025, 0.00025, 0.0]),
        patches=default_patches(), material=copy.deepcopy(IN625),
        beams=[super_gaussian_beam([[1, 0.0, 0.0, 0.0, 0.0, 0.0], [0, 0.002, 0.0, 0.0, 179.2, 0.8]])],
        controls=default_controls(**controls), Tinitial=300.0, powderZMin=1e300)


def synthetic_track(n, cell=20e-6, track_start=0.5e-3, track_margin=0.5e-3, power=179.2, speed=0.8, **controls):
    """BASELINE configs[2] family (SURVEY.md 8d cfg 3): uniform-hex block of 20 um cubes, one superGaussian
    track along +x on the top surface at y = 0, conduction + solidification, implicit PCG/DIC by default."""
    nx, ny, nz = n
    lx, ly, lz = nx * cell, ny * cell, nz * cell
    ctl = dict(explicitSolve=0, maxDi=10.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC, endTime=1.0,
               writeInterval=0.0)
    ctl.update(controls)
    x0 = -track_start
    path = [[1, 0.0, 0.0, 0.0, 0.0, 0.0], [0, lx - track_start - track_margin, 0.0, 0.0, power, speed]]
    return dict(
        mesh=dict(n=[nx, ny, nz], min=[x0, -0.5 * ly, -lz], max=[x0 + lx, 0.5 * ly, 0.0]),
        patches=default_patches(), material=copy.deepcopy(IN625), beams=[super_gaussian_beam(path)],
        controls=default_controls(**ctl), Tinitial=300.0, powderZMin=1e300)


def rosenthal(n=(500, 200, 100), cell=20e-6, kappa=25.0, Cp=600.0, power=179.2, speed=0.8, eta=0.33, **controls):
    """BASELINE configs[1] (SURVEY.md 8d cfg 2, synthetic -- not shipped by the reference, finding F7):
    conduction only, constant kappa/Cp, Lf = 0, adiabatic top, one Gaussian track."""
    case = synthetic_track(n, cell=cell, power=power, speed=speed, **controls)
    mat = case["material"]
    mat["kappa"] = [[kappa, 0, 0]] * 3
    mat["Cp"] = [[Cp, 0, 0]] * 3
    mat["Lf"] = 0.0
    case["patches"] = [_patch(abi.BC_ZERO_GRADIENT, [abi.ZMIN]), _patch(abi.BC_ZERO_GRADIENT, [abi.ZMAX]),
                       _patch(abi.BC_ZERO_GRADIENT, [abi.XMIN, abi.XMAX, abi.YMIN, abi.YMAX])]
    case["beams"][0]["eta"] = eta
    return case


def multi_beam(n=(1000, 250, 200), cell=20e-6, **controls):
    """BASELINE configs[3] (SURVEY.md 8d cfg 4): beam1/beam2 of tutorials/multiBeam/constant/heatSourceDict:26-74
    duplicated on four parallel tracks y = +-50 um, +-150 um, alternating directions."""
    case = synthetic_track(n, cell=cell, **controls)
    nx = n[0]
    x1 = nx * cell - 1.0e-3
    beams = []
    for q, y in enumerate([50e-6, -50e-6, 150e-6, -150e-6]):
        fwd = [[1, 0.0, y, 0.0, 0.0, 0.0], [0, x1, y, 0.0, 179.2, 0.8]]
        bwd = [[1, x1, y, 0.0, 0.0, 0.0], [0, 0.0, y, 0.0, 179.2, 0.8]]
        path = fwd if q % 2 == 0 else bwd
        beams.append(super_gaussian_beam(path) if q % 2 == 0 else modified_super_gaussian_beam(path))
    case["beams"] = beams
    return case


def multi_layer_pbf(n=(1000, 500, 400), cell=20e-6, hatch=100e-6, nHatch=4, **controls):
    """BASELINE configs[4] (SURVEY.md 8d cfg 5): one powder layer (alpha.powder = 1 for z > -40 um,
    tutorials/multiLayerPBF/system/setFieldsDict:18-29), fixedValue 300 bottom/sides (0/T:23-40), transient
    modifiedSuperGaussian (constant/heatSourceDict:26-47), bidirectional raster (createScanPathDict:18-34)."""
    ctl = dict(maxDi=100.0, Tmax=3300.0)
    ctl.update(controls)
    case = synthetic_track(n, cell=cell, **ctl)
    nx = n[0]
    x1 = nx * cell - 1.0e-3
    path = []
    for hline in range(nHatch):
        y = (hline - 0.5 * (nHatch - 1)) * hatch
        a, b = (0.0, x1) if hline % 2 == 0 else (x1, 0.0)
        path.append([1, a, y, 0.0, 0.0, 0.5e-3 if hline else 0.0])
        path.append([0, b, y, 0.0, 195.0, 0.8])
    case["beams"] = [modified_super_gaussian_beam(path)]
    case["patches"] = [_patch(abi.BC_FIXED_VALUE, [abi.ZMIN], value=300.0),
                       _patch(abi.BC_MIXED_TEMPERATURE, [abi.ZMAX], h=10.0, emissivity=0.4, Tinf=300.0),
                       _patch(abi.BC_FIXED_VALUE, [abi.XMIN, abi.XMAX, abi.YMIN, abi.YMAX], value=300.0)]
    case["material"]["kappa"][2] = [0.21767, 0.00102, 0.0]   # tutorials/multiLayerPBF/constant/transportProperties:30-33
    case["material"]["Cp"][2] = [317.53, 0.0, 0.0]
    case["powderZMin"] = -40e-6
    return case


class BuiltCase:
    """Owns the ctypes buffers a b200_case_desc points into."""

    def __init__(self, case):
        self.case = case
        self._keep = []
        d = abi.CaseDesc()
        m = case["mesh"]
        d.mesh.nx, d.mesh.ny, d.mesh.nz = m["n"]
        for q in range(3):
            d.mesh.min[q] = m["min"][q]
            d.mesh.max[q] = m["max"][q]
        patches = (abi.Patch * len(case["patches"]))()
        for i, p in enumerate(case["patches"]):
            patches[i].type = p["type"]
            patches[i].nSides = len(p["sides"])
            for q, s in enumerate(p["sides"]):
                patches[i].sides[q] = s
            patches[i].value, patches[i].h = p["value"], p["h"]
            patches[i].emissivity, patches[i].Tinf = p["emissivity"], p["Tinf"]
        d.nPatches, d.patches = len(case["patches"]), patches
        self._keep.append(patches)
        mat = case["material"]
        for ph in range(3):
            for q in range(3):
                d.material.kappa[ph][q] = mat["kappa"][ph][q]
                d.material.Cp[ph][q] = mat["Cp"][ph][q]
        d.material.rho, d.material.Lf = mat["rho"], mat["Lf"]
        nT = len(mat["thermoT"])
        tx = (C.c_double * nT)(*mat["thermoT"])
        ty = (C.c_double * nT)(*mat["thermoAlpha"])
        d.material.nThermo, d.material.thermoT, d.material.thermoAlpha = nT, tx, ty
        self._keep += [tx, ty]
        beams = (abi.Beam * max(1, len(case["beams"])))()
        for i, b in enumerate(case["beams"]):
            fill_beam(beams[i], b, self._keep)
        d.nBeams, d.beams = len(case["beams"]), beams
        self._keep.append(beams)
        for k, v in case["controls"].items():
            setattr(d.controls, k, v)
        d.Tinitial, d.powderZMin = case["Tinitial"], case["powderZMin"]
        self.desc = d

    @property
    def n_cells(self):
        n = self.case["mesh"]["n"]
        return n[0] * n[1] * n[2]


def fill_beam(dst, b, keep):
    for k in ("heatSourceModel", "transient", "isoValue", "k", "m", "A", "B", "absorptionModel", "eta",
              "kellyGeometry", "eta0", "etaMin", "deltaT", "hitPathIntervals"):
        setattr(dst, k, b[k])
    for q in range(3):
        dst.dimensions[q] = b["dimensions"][q]
        dst.nPoints[q] = b["nPoints"][q]
    flat = [float(v) for seg in b["path"] for v in seg]
    buf = (C.c_double * max(1, len(flat)))(*flat)
    keep.append(buf)
    dst.nSegments, dst.path = len(b["path"]), buf
    return dst


def make_beam(b):
    keep = []
    dst = abi.Beam()
    fill_beam(dst, b, keep)
    dst._keep = keep
    return dst
