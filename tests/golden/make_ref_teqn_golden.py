"""Regenerates tests/golden/ref_teqn.json  --  python tests/golden/make_ref_teqn_golden.py   (needs /root/reference)

These ARE outputs of the reference's code: oracle/_ref/refTEqn is the reference's own temperature equation -- thermo/TEqn.H, which
#includes thermo/thermoScheme.H and thermo/thermoSource.H -- included verbatim where it lies against the fvScalarMatrix stand-in
of oracle/ref_stubs/teqn (oracle/Makefile).  The fragment is a function of the state after runTime++ (T = T.oldTime(), alpha.solid
and its old time, kappa, Cp, this step's qDot, the patch values of T, deltaT) and returns T, alpha.solid, the patch values of T and
the Krylov iterations of every corrector pass.  The states come from running the CPU oracle on small cases; tests/test_ref_teqn.py
re-runs the same cases, so only digests of the fragment's outputs are stored: per step the iteration list, and for T, alpha.solid
and every patch's face values a few sampled entries as hex floats plus the sum of all entries (exact in hex)."""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "refTEqn")
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def configs():
    """name -> (case, steps).  writeInterval 0 and hitPathIntervals off: nothing but the fragments decides the time step."""
    from additivefoam_b200 import abi, cases
    out = {}
    c = cases.amb2018_02_b(n=(40, 10, 6), writeInterval=0.0)                                   # as shipped: explicitSolve true (fvc::laplacian, diagonal solve)
    out["amb_explicit"] = (c, 30)
    c = cases.amb2018_02_b(n=(40, 10, 6), writeInterval=0.0, explicitSolve=0, maxDi=10.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC)
    out["amb_implicit_pcg"] = (c, 25)
    c = cases.amb2018_02_b(n=(30, 8, 6), writeInterval=0.0, explicitSolve=0, maxDi=10.0)       # as-shipped solver: PBiCGStab / DILU
    out["amb_implicit_bicg"] = (c, 20)
    c = cases.synthetic_track((36, 12, 8), writeInterval=0.0, deltaT=2e-6, Tmax=2500.0, nThermoCorrectors=6, thermoTolerance=1e-6,
                              explicitSolve=0, maxDi=20.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC)   # Tmax penalty active
    out["track_tmax"] = (c, 40)
    c = cases.multi_layer_pbf(n=(32, 16, 10), nHatch=2, writeInterval=0.0, deltaT=1e-6)        # fixedValue walls, powder layer
    out["pbf"] = (c, 30)
    # ddtSchemes default backward (thermoScheme.H:41-49): backward fvm::ddt(T) and fvc::ddt(alpha1), rDeltaT of the Sp terms scaled by coefft
    c = cases.amb2018_02_b(n=(40, 10, 6), writeInterval=0.0, explicitSolve=0, maxDi=10.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC,
                           ddtScheme=abi.DDT_BACKWARD)
    out["amb_backward_pcg"] = (c, 25)
    c = cases.synthetic_track((36, 12, 8), writeInterval=0.0, deltaT=2e-6, Tmax=2500.0, nThermoCorrectors=6, thermoTolerance=1e-6,
                              explicitSolve=0, maxDi=20.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC, ddtScheme=abi.DDT_BACKWARD)
    out["track_backward_tmax"] = (c, 40)
    c = cases.multi_layer_pbf(n=(32, 16, 10), nHatch=2, writeInterval=0.0, deltaT=1e-6, ddtScheme=abi.DDT_BACKWARD)
    out["pbf_backward"] = (c, 30)
    # ddtSchemes default CrankNicolson psi (thermoScheme.H:51-65): the scheme's ddt0 fields, coefft = 1 + ocCoeff on the Sp terms
    c = cases.amb2018_02_b(n=(40, 10, 6), writeInterval=0.0, explicitSolve=0, maxDi=10.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC,
                           ddtScheme=abi.DDT_CRANK_NICOLSON, ocCoeff=0.9)
    out["amb_cn09_pcg"] = (c, 25)
    c = cases.synthetic_track((36, 12, 8), writeInterval=0.0, deltaT=2e-6, Tmax=2500.0, nThermoCorrectors=6, thermoTolerance=1e-6,
                              explicitSolve=0, maxDi=20.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC, ddtScheme=abi.DDT_CRANK_NICOLSON,
                              ocCoeff=1.0)
    out["track_cn1_tmax"] = (c, 40)
    c = cases.multi_layer_pbf(n=(32, 16, 10), nHatch=2, writeInterval=0.0, deltaT=1e-6, ddtScheme=abi.DDT_CRANK_NICOLSON, ocCoeff=0.5)
    out["pbf_cn05"] = (c, 30)
    c = cases.amb2018_02_b(n=(40, 10, 6), writeInterval=0.0, ddtScheme=abi.DDT_CRANK_NICOLSON, ocCoeff=0.9)   # as shipped: explicit steps are Euler, ddt0(T) lives on in setDeltaT.H
    out["amb_cn09_as_shipped_explicit"] = (c, 30)
    for case, _ in out.values():
        for b in case["beams"]:
            b["hitPathIntervals"] = 0
    return out


def request(case, ddt, explicit, dt, dt0, Tliq, Tsol, nThermoCorr, T, Told_levels, alpha1, alpha_old_levels, kappa, Cp, qDot, patchTb, cn=None):
    m, ct, mat = case["mesh"], case["controls"], case["material"]
    n = m["n"]
    d = [(m["max"][q] - m["min"][q]) / n[q] for q in range(3)]
    Tmax = ct["Tmax"] if ct["Tmax"] > 0 else 1e300
    lines = ["teqn %d %d %d %r %r %r  %s %d  %r %r  %r %r %r %r %r %r %r" % (*n, *d, ddt, int(explicit), dt, dt0, mat["rho"], mat["Lf"], Tliq, Tsol,
                                                                           ct["thermoTolerance"], float(nThermoCorr), Tmax),
             "%d %d %r %r %d %d" % (ct["solver"], ct["preconditioner"], ct["tolerance"], ct["relTol"], ct["minIter"], ct["maxIter"]),
             "%d %s" % (len(mat["thermoT"]), " ".join("%r %r" % (x, y) for x, y in zip(mat["thermoT"], mat["thermoAlpha"]))),
             "%d" % len(case["patches"])]
    for p in case["patches"]:
        lines.append("%d %d %s %r %r %r %r" % (p["type"], len(p["sides"]), " ".join(str(s) for s in p["sides"]), p["value"], p["h"], p["emissivity"], p["Tinf"]))
    lines.append("%d %d" % (len(Told_levels), len(alpha_old_levels)))
    ddt0 = []
    if cn is None:
        lines.append("1.0 0  0 0 0  0 0 0")
    else:
        st = cn["state"]
        lines.append("%r %d  %d %d %d  %d %d %d" % (ct["ocCoeff"], st["timeIndex"], st["existsT"], st["startT"], st["tiT"], st["existsA"], st["startA"], st["tiA"]))
        ddt0 = ([cn["ddt0T"]] if st["existsT"] else []) + ([cn["ddt0A"]] if st["existsA"] else [])
    for f in [T, *Told_levels, alpha1, *alpha_old_levels, kappa, Cp, qDot, *ddt0, *patchTb]:
        lines.append(" ".join(repr(float(v)) for v in f))
    return "\n".join(lines) + "\n"


def run_ref(req):
    out = subprocess.run([REF], input=req, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return out.stdout.split("\n")


def parse(lines, nPatches):
    H = float.fromhex
    its = [int(v) for v in lines[0].split()][1:]
    T = np.array([H(v) for v in lines[1].split()])
    A = np.array([H(v) for v in lines[2].split()])
    Tb = [np.array([H(v) for v in lines[3 + 2 * p].split()]) for p in range(nPatches)]
    vf = [np.array([H(v) for v in lines[4 + 2 * p].split()]) for p in range(nPatches)]
    return its, T, A, Tb, vf


def digest(a):
    """A few entries and the (sequential, hence reproducible) sum, as hex floats."""
    a = np.asarray(a, dtype=np.float64)
    idx = sorted({0, len(a) // 3, len(a) // 2, (2 * len(a)) // 3, len(a) - 1, int(np.argmax(a))})
    s = 0.0
    for v in a:
        s += float(v)
    return {"n": len(a), "idx": idx, "val": [float(a[i]).hex() for i in idx], "sum": float(s).hex()}


def step_inputs(o, case, abi):
    """State the fragment sees in the step that `o.step()` is about to take: call BEFORE the step for the start-of-step fields.
    After runTime++ the first old-time level of T and alpha.solid is the start-of-step field and the second one (backward scheme) the
    first level of before; alpha.solid has no levels at all until its first fvc::ddt (the fragment then creates them, and the scheme
    uses deltaT0 = great for that one call); deltaT0 is the deltaT saved at the previous runTime++."""
    dt_save, TnOld, a1nOld = o.time_scheme_state()
    backward = case["controls"]["ddtScheme"] == abi.DDT_BACKWARD
    T, A = o.field(abi.FIELD_T), o.field(abi.FIELD_ALPHA_SOLID)
    T_levels, A_levels = [T], [A]
    if backward:
        T_levels = [T, o.field(abi.FIELD_T_OLD)]
        assert a1nOld in (0, 2)
        A_levels = [] if a1nOld == 0 else [A, o.field(abi.FIELD_ALPHA_SOLID_OLD)]
    cn = case["controls"]["ddtScheme"] == abi.DDT_CRANK_NICOLSON
    return dict(T=T, alpha1=A, T_levels=T_levels, A_levels=A_levels, deltaT0=dt_save,
                ddt="backward" if backward else "CrankNicolson" if cn else "Euler",
                Told=o.field(abi.FIELD_T_OLD), Aold=o.field(abi.FIELD_ALPHA_SOLID_OLD), patches=[tb.copy() for tb, _ in o.patch_state()])


def cn_inputs(o, pre):
    """CrankNicolson, AFTER the step: the scheme's state as that step's TEqn.H found it (setDeltaT.H has already used ddt0(T) at the
    previous time index; runTime++ has shifted the old-time levels that existed) and the old-time levels of T and alpha.solid it saw."""
    st, d0T, d0A = o.cn_snapshot()
    pre["T_levels"] = [pre["T"]] + ([pre["Told"]] if st["TnOld"] >= 2 else [])
    pre["A_levels"] = [] if st["a1nOld"] == 0 else [pre["alpha1"]] + ([pre["Aold"]] if st["a1nOld"] >= 2 else [])
    return dict(state=st, ddt0T=d0T, ddt0A=d0A)


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    assert os.path.exists(REF), "oracle/_ref/refTEqn was not built: /root/reference is needed"
    import oracle_lib as ol
    from additivefoam_b200 import abi
    out = {}
    for name, (case, steps) in configs().items():
        o = ol.OracleCase(case)
        rows = []
        worst = 0.0
        for s in range(steps):
            pre = step_inputs(o, case, abi)
            rep = o.step()
            kappa, Cp, qDot = (o.field(f) for f in (abi.FIELD_KAPPA, abi.FIELD_CP, abi.FIELD_QDOT))
            # after runTime++ the old-time level of T and alpha.solid is the start-of-step field
            cn = cn_inputs(o, pre) if pre["ddt"] == "CrankNicolson" else None
            req = request(case, pre["ddt"], rep.explicitStep, rep.deltaT, pre["deltaT0"], o.L.afo_case_Tliq(o.h), o.L.afo_case_Tsol(o.h),
                          case["controls"]["nThermoCorrectors"] if (rep.totalPower > 1e-15 or pre["alpha1"].min() < 1 - 1e-15) else 1,
                          pre["T"], pre["T_levels"], pre["alpha1"], pre["A_levels"], kappa, Cp, qDot, pre["patches"], cn)
            its, T, A, Tb, vf = parse(run_ref(req), len(case["patches"]))
            To, Ao = o.field(abi.FIELD_T), o.field(abi.FIELD_ALPHA_SOLID)
            worst = max(worst, float(np.max(np.abs(T - To)) / np.max(np.abs(To))), float(np.max(np.abs(A - Ao))))
            for (tbo, vfo), tbr, vfr in zip(o.patch_state(), Tb, vf):
                worst = max(worst, float(np.max(np.abs(tbo - tbr)) / np.max(np.abs(tbr))), float(np.max(np.abs(vfo - vfr))))
            rows.append({"iters": its, "T": digest(T), "alpha1": digest(A), "Tb": [digest(t) for t in Tb], "valueFraction": [digest(v) for v in vf]})
        out[name] = rows
        print(name, "steps", steps, "correctors of the last step", len(rows[-1]["iters"]), "oracle vs reference text (T, alpha.solid, patch values, value fractions), worst:", worst)
    with open(os.path.join(HERE, "ref_teqn.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
