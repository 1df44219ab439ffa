"""Regenerates tests/golden/*.npz and known_answers.json.   python tests/golden/make_golden.py

What the vectors are -- and are not.  AdditiveFOAM ships no tests or golden vectors, and it cannot be built or imported in
this image (it is an OpenFOAM-10 application), so these are NOT outputs of the reference binary.
They are
  * known_answers.json: closed-form values derived by hand from the reference's formulas (each with the
    reference file:line of the formula), written out here with plain Python `math` -- independent of oracle/ and of the CUDA code;
  * *.npz: outputs of the CPU oracle (oracle/, the restatement of the reference algorithm) on small seeded / tutorial-derived
    inputs, frozen so that (a) a change of the oracle is noticed (tests/test_golden.py, CPU) and (b) the CUDA path is compared
    with vectors that exist independently of the oracle build on the GPU box (tests/test_golden.py, -m gpu).
The inputs are rebuilt by tests/test_golden.py from the same functions below, so only outputs are stored.
"""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


# ------------------------------------------------------------------------------------------------ inputs (shared with the tests)
def ldu_problem(which):
    import helpers
    if which == "block_sym":
        dims = (23, 11, 7)
        lo, up = helpers.block_addressing_np(dims)
        asym = False
    elif which == "block_asym":
        dims = (17, 9, 5)
        lo, up = helpers.block_addressing_np(dims)
        asym = True
    else:
        dims = None
        lo, up = helpers.random_ldu_addressing(3000, 7000, 21)
        asym = which == "general_asym"
    n = dims[0] * dims[1] * dims[2] if dims else 3000
    diag, upper, lower, x, b = helpers.spd_coefficients(n, lo, up, seed=1234, asym=asym)
    return dims, n, lo, up, diag, upper, lower, x, b


LDU_PROBLEMS = ["block_sym", "block_asym", "general_sym", "general_asym"]


def case_inputs(which):
    from additivefoam_b200 import abi, cases
    if which == "amb2018_explicit":            # configs[0] as shipped, first 60 steps
        return cases.amb2018_02_b(), 60
    if which == "amb2018_implicit_pcg_dic":    # configs[0] geometry coarsened, explicitSolve false, bPCG / DIC
        return cases.amb2018_02_b(n=(60, 12, 8), explicitSolve=0, maxDi=10.0, solver=abi.SOLVER_PCG,
                                  preconditioner=abi.PRECOND_DIC), 40
    if which == "amb2018_implicit_bicg_dilu":  # as-shipped solver choice (fvSolution:34-42)
        return cases.amb2018_02_b(n=(60, 12, 8), explicitSolve=0, maxDi=10.0), 40
    if which == "multibeam":
        c = cases.multi_beam(n=(80, 30, 10), endTime=0.002)
        c["controls"]["deltaT"] = 1e-6
        return c, 25
    if which == "multilayer":
        c = cases.multi_layer_pbf(n=(60, 24, 12), nHatch=2, endTime=0.002)
        c["controls"]["deltaT"] = 1e-6
        return c, 30
    raise KeyError(which)


CASES = ["amb2018_explicit", "amb2018_implicit_pcg_dic", "amb2018_implicit_bicg_dilu", "multibeam", "multilayer"]
REPORT_FIELDS = ["time", "deltaT", "alphaCoNum", "DiNum", "totalPower", "explicitStep", "nThermoCorr", "thermoResidual",
                 "nSolveIterations", "lastSolveIterations", "lastInitialResidual", "lastFinalResidual"]


def beam_inputs():
    """(name, beam dict, position, dims) for a qDot evaluation on a 40 x 30 x 12 block of 20 um cells."""
    from additivefoam_b200 import cases
    path = [[1, 0, 0, 0, 100.0, 0.0], [0, 1e-3, 0, 0, 100.0, 0.8]]
    out = [("superGaussian", cases.super_gaussian_beam(path), (4.1e-4, 3.07e-4, 2.4e-4), (85e-6, 85e-6, 30e-6)),
           ("modifiedSuperGaussian", cases.modified_super_gaussian_beam(path), (3.9e-4, 2.9e-4, 2.4e-4), (40e-6, 40e-6, 50e-6)),
           ("projectedGaussian", cases.projected_gaussian_beam(path), (4.0e-4, 3.0e-4, 2.4e-4), (50e-6, 50e-6, 80e-6))]
    return out


BEAM_MESH = dict(n=(40, 30, 12), min=(0.0, 0.0, 0.0), max=(8e-4, 6e-4, 2.4e-4))
BEAM_POWER = 100.0


# ------------------------------------------------------------------------------------------------ hand-derived known answers
def known_answers():
    g = math.gamma
    ka = {}
    # superGaussian::V0 (superGaussian.C:75-82): s = dims / 2^(1/k); V0 = (2/3) sx sy sz pi Gamma(1 + 3/k)
    k = 2.0
    s = [d / 2 ** (1 / k) for d in (85e-6, 85e-6, 30e-6)]
    ka["superGaussian_V0_k2_85_85_30um"] = (2.0 / 3.0) * s[0] * s[1] * s[2] * math.pi * g(1 + 3 / k)
    # modifiedSuperGaussian::V0 (modifiedSuperGaussian.C:89-100)
    k, m = 7.95, 2.72
    a = 2 ** (1 / k)
    sx, sy, sz = 40e-6 / a, 40e-6 / a, 30e-6
    ka["modifiedSuperGaussian_V0_k7.95_m2.72_40_40_30um"] = sx * sy * sz * math.pi * g(1 + 2 / k) * g(1 + 1 / m) * g(1 + 2 / m) / g(1 + 3 / m)
    # absorbed power of configs[0] (heatSourceDict:34 eta 0.33, scanPath:3 179.2 W) and peak qDot
    ka["amb2018_absorbed_power_W"] = 0.33 * 179.2
    ka["amb2018_peak_qdot_W_m3"] = 0.33 * 179.2 / ka["superGaussian_V0_k2_85_85_30um"]
    # Kelly cone (KellyAbsorption.C:73-98), eta0 = 0.28, etaMin = 0.35
    ka["kelly_cone_eta_aspect_0.75"] = 0.35
    ka["kelly_cone_eta_aspect_2"] = 0.6453157893721592
    # block counts (blockMeshDict:38): cells / internal faces / boundary faces of 150 x 25 x 15
    nx, ny, nz = 150, 25, 15
    ka["amb2018_cells"] = nx * ny * nz
    ka["amb2018_internal_faces"] = (nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1)
    ka["amb2018_boundary_faces"] = 2 * (nx * ny + ny * nz + nx * nz)
    ka["Tliq"], ka["Tsol"] = 1620.0, 1410.0         # thermoPath:2-3 through createFields.H:59-71
    ka["mushy_dFdT"] = -1.0 / 210.0                  # thermoSource.H:47 on the two-row table
    return ka


# ------------------------------------------------------------------------------------------------ oracle outputs
def oracle_ldu(L, which):
    import ctypes as C
    import oracle_lib as ol
    from additivefoam_b200 import abi
    P = ol.P
    dims, n, lo, up, diag, upper, lower, x, b = ldu_problem(which)
    out = {}
    y = np.empty(n)
    L.afo_ldu_amul(n, len(lo), P(lo), P(up), P(diag), P(upper), P(lower), P(x), P(y))
    out["amul"] = y
    for name, kind in (("dic", abi.PRECOND_DIC), ("dilu", abi.PRECOND_DILU), ("diagonal", abi.PRECOND_DIAGONAL)):
        if name == "dic" and lower is not None:
            continue
        w = np.empty(n)
        L.afo_ldu_precondition(n, len(lo), P(lo), P(up), P(diag), P(upper), P(lower), kind, P(b), P(w), None)
        out["precond_" + name] = w
    solver, precond = (abi.SOLVER_PCG, abi.PRECOND_DIC) if lower is None else (abi.SOLVER_PBICGSTAB, abi.PRECOND_DILU)
    psi = x.copy()
    perf = abi.SolverPerf()
    log = np.empty(4096)
    nlog = C.c_int32()
    L.afo_ldu_solve(n, len(lo), P(lo), P(up), P(diag), P(upper), P(lower), P(b), P(psi), solver, precond, 1e-10, 0.0, 1, 60,
                    C.byref(perf), P(log), 4096, C.byref(nlog))
    out["solve_psi"] = psi
    out["solve_perf"] = np.array([perf.initialResidual, perf.finalResidual, perf.nIterations, perf.converged, perf.singular], dtype=np.float64)
    return out


def oracle_case(L, which):
    import oracle_lib as ol
    from additivefoam_b200 import abi
    case, steps = case_inputs(which)
    o = ol.OracleCase(case, lib=L)
    reps = np.zeros((steps, len(REPORT_FIELDS)))
    for s in range(steps):
        r = o.step()
        reps[s] = [float(getattr(r, f)) for f in REPORT_FIELDS]
    out = {"reports": reps, "T": o.field(abi.FIELD_T), "alpha_solid": o.field(abi.FIELD_ALPHA_SOLID)}
    o.close()
    return out


def oracle_beams(L):
    import ctypes as C
    import oracle_lib as ol
    from additivefoam_b200 import abi, cases
    P = ol.P
    out = {}
    bm = abi.BlockMesh()
    bm.nx, bm.ny, bm.nz = BEAM_MESH["n"]
    for q in range(3):
        bm.min[q], bm.max[q] = BEAM_MESH["min"][q], BEAM_MESH["max"][q]
    n = bm.nx * bm.ny * bm.nz
    for name, b, pos, dims in beam_inputs():
        beam = cases.make_beam(b)
        p, d = np.array(pos), np.array(dims)
        q = np.zeros(n)
        sumw, vol, absorbed = np.zeros(1), np.zeros(1), np.zeros(1)
        L.afo_beam_qdot(C.byref(bm), C.byref(beam), P(p), P(d), BEAM_POWER, P(q), P(sumw), P(vol), P(absorbed))
        out[name + "_qdot"] = q
        out[name + "_scalars"] = np.array([sumw[0], vol[0], absorbed[0]])     # sum of weights, volume used, absorbed power
    return out


def main():
    import oracle_lib as ol
    L = ol.load()
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(known_answers(), f, indent=1, sort_keys=True)
    for w in LDU_PROBLEMS:
        np.savez_compressed(os.path.join(HERE, f"ldu_{w}.npz"), **oracle_ldu(L, w))
    for w in CASES:
        np.savez_compressed(os.path.join(HERE, f"case_{w}.npz"), **oracle_case(L, w))
    np.savez_compressed(os.path.join(HERE, "beams.npz"), **oracle_beams(L))
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
