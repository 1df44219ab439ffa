"""Diagnostic: which ingredient of a case makes the GPU and the oracle part ways, and at which step.
Usage: python tools/diag_case_parity.py   (needs a GPU)"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
from additivefoam_b200 import abi, cases, lib  # noqa: E402

ctx = lib.Context(0)
O = ol.load()


def run(name, case, steps=12):
    g, o = lib.Case(ctx, case), ol.OracleCase(case, lib=O)
    first = None
    for s in range(steps):
        rg, ro = g.step(), o.step()
        Tg, To = g.field(abi.FIELD_T), o.field(abi.FIELD_T)
        e = float(np.max(np.abs(Tg - To)) / np.max(np.abs(To)))
        if first is None and e > 1e-12:
            first = (s, e, int(np.argmax(np.abs(Tg - To))), rg.explicitStep, ro.explicitStep, rg.nThermoCorr, ro.nThermoCorr,
                     rg.lastSolveIterations, ro.lastSolveIterations, rg.deltaT, ro.deltaT)
    print(f"{name:40s} final {e:.3e}  first>1e-12: {first}")
    g.close(); o.close()


base = cases.amb2018_02_b(n=(20, 8, 6), explicitSolve=0, maxDi=8.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC)
base["controls"]["deltaT"] = 5e-7
run("baseline", base)
c = copy.deepcopy(base); c["patches"][0] = cases._patch(abi.BC_FIXED_VALUE, [abi.ZMIN], value=300.0); run("fixedValue base = Tinitial", c)
c = copy.deepcopy(base); c["patches"][0] = cases._patch(abi.BC_FIXED_VALUE, [abi.ZMIN], value=290.0); run("fixedValue base 290 != Tinitial 300", c)
c = copy.deepcopy(c); c["controls"]["explicitSolve"] = 1; c["controls"]["maxDi"] = 1.0; run("... explicit", c)
c = copy.deepcopy(base); c["Tinitial"] = 320.0; run("Tinitial 320 (mixed Tinf 300)", c)
c = copy.deepcopy(base); c["patches"][2] = cases._patch(abi.BC_FIXED_VALUE, [abi.XMIN, abi.XMAX, abi.YMIN, abi.YMAX], value=290.0); run("fixedValue sides 290", c)
c = copy.deepcopy(base); c["material"]["thermoT"] = [1400.0, 1500.0, 1600.0]; c["material"]["thermoAlpha"] = [1.0, 0.6, 0.0]; run("3-row thermoPath", c)
c = copy.deepcopy(base); c["powderZMin"] = -5e-5; run("powder layer", c)
