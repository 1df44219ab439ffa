"""ctypes mirror of ``include/additivefoam_b200.h`` (the C ABI / drop-in boundary).

Only structure layouts and enum values live here; no computation.  The same structures
describe a case to the CUDA library (``additivefoam_b200/csrc`` -> ``libadditivefoam_b200.so``)
and, in tests only, to the CPU oracle.
"""
import ctypes as C

# enums (keep in sync with the header)
SOLVER_PCG, SOLVER_PBICGSTAB, SOLVER_DIAGONAL = 0, 1, 2
PRECOND_NONE, PRECOND_DIAGONAL, PRECOND_DIC, PRECOND_DILU = 0, 1, 2, 3
BC_ZERO_GRADIENT, BC_FIXED_VALUE, BC_MIXED_TEMPERATURE = 0, 1, 2
XMIN, XMAX, YMIN, YMAX, ZMIN, ZMAX = range(6)
SUPER_GAUSSIAN, MODIFIED_SUPER_GAUSSIAN, PROJECTED_GAUSSIAN = 0, 1, 2
ABSORPTION_CONSTANT, ABSORPTION_KELLY = 0, 1
KELLY_CONE, KELLY_CYLINDER = 0, 1
FIELD_T, FIELD_ALPHA_SOLID, FIELD_ALPHA_POWDER, FIELD_KAPPA, FIELD_CP, FIELD_QDOT, FIELD_DFDT, FIELD_T0, FIELD_T_OLD, FIELD_ALPHA_SOLID_OLD, \
    FIELD_T_OLD_OLD, FIELD_ALPHA_SOLID_OLD_OLD, FIELD_DDT0_T, FIELD_DDT0_ALPHA_SOLID = range(14)
DDT_EULER, DDT_BACKWARD, DDT_CRANK_NICOLSON = 0, 1, 2

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class BlockMesh(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("min", C.c_double * 3), ("max", C.c_double * 3)]


class Patch(C.Structure):
    _fields_ = [("type", C.c_int32), ("nSides", C.c_int32), ("sides", C.c_int32 * 6),
                ("value", C.c_double), ("h", C.c_double), ("emissivity", C.c_double), ("Tinf", C.c_double)]


class Material(C.Structure):
    _fields_ = [("kappa", (C.c_double * 3) * 3), ("Cp", (C.c_double * 3) * 3),
                ("rho", C.c_double), ("Lf", C.c_double),
                ("nThermo", C.c_int32), ("thermoT", c_double_p), ("thermoAlpha", c_double_p)]


class Beam(C.Structure):
    _fields_ = [("heatSourceModel", C.c_int32), ("dimensions", C.c_double * 3), ("nPoints", C.c_int32 * 3),
                ("transient", C.c_int32), ("isoValue", C.c_double),
                ("k", C.c_double), ("m", C.c_double), ("A", C.c_double), ("B", C.c_double),
                ("absorptionModel", C.c_int32), ("eta", C.c_double),
                ("kellyGeometry", C.c_int32), ("eta0", C.c_double), ("etaMin", C.c_double),
                ("deltaT", C.c_double), ("hitPathIntervals", C.c_int32),
                ("nSegments", C.c_int32), ("path", c_double_p)]


class Controls(C.Structure):
    _fields_ = [("startTime", C.c_double), ("endTime", C.c_double), ("deltaT", C.c_double),
                ("adjustTimeStep", C.c_int32),
                ("maxCo", C.c_double), ("maxDi", C.c_double), ("maxAlphaCo", C.c_double), ("maxDeltaT", C.c_double),
                ("writeInterval", C.c_double),
                ("nThermoCorrectors", C.c_int32), ("thermoTolerance", C.c_double), ("Tmax", C.c_double),
                ("explicitSolve", C.c_int32),
                ("solver", C.c_int32), ("preconditioner", C.c_int32),
                ("tolerance", C.c_double), ("relTol", C.c_double), ("minIter", C.c_int32), ("maxIter", C.c_int32),
                ("ddtScheme", C.c_int32), ("ocCoeff", C.c_double)]


class CaseDesc(C.Structure):
    _fields_ = [("mesh", BlockMesh), ("nPatches", C.c_int32), ("patches", C.POINTER(Patch)),
                ("material", Material), ("nBeams", C.c_int32), ("beams", C.POINTER(Beam)),
                ("controls", Controls), ("Tinitial", C.c_double), ("powderZMin", C.c_double)]


class StepReport(C.Structure):
    _fields_ = [("time", C.c_double), ("deltaT", C.c_double), ("alphaCoNum", C.c_double), ("DiNum", C.c_double),
                ("totalPower", C.c_double), ("explicitStep", C.c_int32), ("nThermoCorr", C.c_int32),
                ("thermoResidual", C.c_double), ("nSolveIterations", C.c_int32), ("lastSolveIterations", C.c_int32),
                ("lastInitialResidual", C.c_double), ("lastFinalResidual", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SolverPerf(C.Structure):
    _fields_ = [("initialResidual", C.c_double), ("finalResidual", C.c_double),
                ("nIterations", C.c_int32), ("converged", C.c_int32), ("singular", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}
