"""Regenerates tests/golden/ref_bc.json  --  python tests/golden/make_ref_bc_golden.py   (needs /root/reference)

These ARE outputs of the reference's code: oracle/_ref/refBC is the reference's own derivedFvPatchFields/mixedTemperature/
mixedTemperatureFvPatchScalarField.C with its real header, compiled where it lies against stand-in headers (oracle/ref_stubs/bc,
oracle/Makefile).  Per patch (h, emissivity, Tinf): faces with patch temperature Tp, conductivity kappa and deltaCoeffs ->
refValue, refGradient, valueFraction after updateCoeffs(), as hex floats."""
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "refBC")

# (h, emissivity, Tinf): the shipped top patch (0/T: h 10, emissivity 0.4, Tinf 300), radiation only, convection only, neither
# (hEff = 0 -> the max(hEff, 1e-15) floor), a strong film, a hot environment
PATCHES = [(10.0, 0.4, 300.0), (0.0, 1.0, 300.0), (250.0, 0.0, 293.15), (0.0, 0.0, 300.0), (1e5, 0.9, 500.0), (25.0, 0.35, 1200.0)]


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    assert os.path.exists(REF), "oracle/_ref/refBC was not built: /root/reference is needed"
    rng = np.random.default_rng(7)
    out = []
    for h, eps, Tinf in PATCHES:
        n = 40
        Tp = np.concatenate([[300.0, Tinf, 1410.0, 1620.0, 3300.0], rng.uniform(250.0, 3500.0, n - 5)])
        kappa = np.concatenate([[9.8, 0.21767, 30.0], rng.uniform(0.2, 35.0, n - 3)])                    # solid, powder, liquid IN625 ranges
        delta = np.concatenate([[1e5, 2e5, 8e4], 1.0 / rng.uniform(2e-6, 5e-5, n - 3)])                 # 1/(half cell): 10 um, 5 um, 12.5 um
        req = ["mixed %r %r %r %d" % (h, eps, Tinf, n)] + ["%r %r %r" % (float(a), float(b), float(c)) for a, b, c in zip(Tp, kappa, delta)]
        rows = [ln.split() for ln in subprocess.run([REF], input="\n".join(req) + "\n", capture_output=True, text=True, check=True).stdout.split("\n") if ln]
        assert len(rows) == n
        out.append(dict(h=h, emissivity=eps, Tinf=Tinf, Tp=[float(v).hex() for v in Tp], kappa=[float(v).hex() for v in kappa],
                        deltaCoeffs=[float(v).hex() for v in delta], refValue=[r[0] for r in rows], refGradient=[r[1] for r in rows],
                        valueFraction=[r[2] for r in rows]))
        print(h, eps, Tinf, [float.fromhex(r[2]) for r in rows[:3]])
    with open(os.path.join(HERE, "ref_bc.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
