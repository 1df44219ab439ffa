"""Regenerates tests/golden/ref_ctl.json  --  python tests/golden/make_ref_ctl_golden.py   (needs /root/reference)

These ARE outputs of the reference's code: oracle/_ref/refCtl is the reference's own setDeltaT.H and solutionControls.H, the two
fragments of the time loop that set the time step and the per-step solution controls, included verbatim where they lie against
stand-in types (oracle/ref_stubs/ctl, oracle/Makefile).  The fragments are functions of the state at the start of a time step
(T, T.oldTime(), kappa, Cp, alpha.solid, deltaT) and of this step's qDot; the states come from running the CPU oracle on small
cases (tests/test_ref_ctl.py re-runs the same cases, so only the fragments' outputs are stored, as hex floats): thermo number,
diffusion number, the new deltaT, nThermoCorr, thermoTol, Tmax, absorbed power, beamOn, fluidInDomain, explicitSolve."""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "refCtl")
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def configs():
    """name -> (case, steps).  hitPathIntervals off: sources.adjustDeltaT leaves the step alone (the beams' own cut is pinned by refBeam /
    refMHS); writeInterval 0: no [UPSTREAM] write-time snapping after the fragment."""
    from additivefoam_b200 import cases
    out = {}
    c = cases.amb2018_02_b(n=(60, 10, 6), writeInterval=0.0)                       # as shipped: explicitSolve true, maxDi 1
    out["amb_explicit"] = (c, 60)
    c = cases.amb2018_02_b(n=(60, 10, 6), writeInterval=0.0, maxDi=50.0, maxAlphaCo=1e3, deltaT=3e-5)   # explicit requested, diffusion number grows past 1: switches to implicit
    out["amb_switches"] = (c, 30)
    c = cases.synthetic_track((40, 12, 8), writeInterval=0.0, deltaT=1e-6, maxAlphaCo=0.2, Tmax=3000.0, nThermoCorrectors=7, thermoTolerance=1e-5)
    out["track_implicit"] = (c, 40)
    c = cases.amb2018_02_b(n=(30, 10, 6), writeInterval=0.0, endTime=1.0)          # beam off from the start: nothing to iterate on (nThermoCorr = 1)
    c["beams"][0]["path"] = [[1, 0.0, 0.0, 0.0, 0.0, 0.0], [0, 0.002, 0.0, 0.0, 0.0, 0.8]]
    out["beam_off"] = (c, 6)
    c = cases.multi_layer_pbf(n=(40, 16, 10), nHatch=2, writeInterval=0.0, deltaT=1e-6)
    out["pbf"] = (c, 40)
    for case, _ in out.values():
        for b in case["beams"]:
            b["hitPathIntervals"] = 0
    return out


def request(case, Tliq, Tsol, dt, T, Told, kappa, Cp, alpha1, qDot):
    m, ct, mat = case["mesh"], case["controls"], case["material"]
    cd = {"maxAlphaCo": ct["maxAlphaCo"], "maxDi": ct["maxDi"]}
    pd = {"nThermoCorrectors": ct["nThermoCorrectors"], "thermoTolerance": ct["thermoTolerance"], "explicitSolve": int(ct["explicitSolve"])}
    if ct["Tmax"] > 0:
        pd["Tmax"] = ct["Tmax"]
    head = "ctl %d %d %d %r %r %r %r %r %r  %r %r %r %r  %d %r %r" % (*m["n"], *m["min"], *m["max"], mat["rho"], Tliq, Tsol, dt,
                                                                       int(ct["adjustTimeStep"]), ct["maxCo"], ct["maxDeltaT"])
    kv = lambda d: "%d %s" % (len(d), " ".join("%s %r" % (k, v) for k, v in d.items()))
    return "\n".join([head, kv(cd), kv(pd)] + [" ".join(repr(float(v)) for v in f) for f in (T, Told, kappa, Cp, alpha1, qDot)]) + "\n"


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    assert os.path.exists(REF), "oracle/_ref/refCtl was not built: /root/reference is needed"
    import oracle_lib as ol
    from additivefoam_b200 import abi
    out = {}
    for name, (case, steps) in configs().items():
        o = ol.OracleCase(case)
        dt = case["controls"]["deltaT"]
        rows = []
        for s in range(steps):
            Told = o.pre_step_state()
            T, kappa, Cp, alpha1 = (o.field(f) for f in (abi.FIELD_T, abi.FIELD_KAPPA, abi.FIELD_CP, abi.FIELD_ALPHA_SOLID))
            rep = o.step()
            qDot = o.field(abi.FIELD_QDOT)
            res = subprocess.run([REF], input=request(case, o.L.afo_case_Tliq(o.h), o.L.afo_case_Tsol(o.h), dt, T, Told, kappa, Cp, alpha1, qDot), capture_output=True, text=True, check=True).stdout.split()
            rows.append(res)
            dt = rep.deltaT
        out[name] = rows
        print(name, [float.fromhex(rows[-1][i]) for i in range(7)], rows[-1][7:], sorted({r[9] for r in rows}), sorted({float.fromhex(r[3]) for r in rows}))
    with open(os.path.join(HERE, "ref_ctl.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
