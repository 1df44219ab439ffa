"""Regenerates tests/golden/ref_models.json  --  python tests/golden/make_ref_models_golden.py   (needs /root/reference)

These ARE outputs of the reference's code: oracle/_ref/refModels is the reference's own superGaussian.C,
modifiedSuperGaussian.C, projectedGaussian.C (weight, V0) and interpolateXY.C, compiled where they lie against stand-in
headers for the few OpenFOAM types they use (oracle/ref_stubs/models, oracle/Makefile).  Values are stored as hex floats."""
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "refModels")


def shape_requests():
    rng = np.random.default_rng(2024)
    out = []
    specs = [("super", dict(k=2.0, m=0, A=0, B=0), (85e-6, 85e-6, 30e-6)), ("super", dict(k=4.7, m=0, A=0, B=0), (60e-6, 45e-6, 38e-6)),
             ("modified", dict(k=7.95, m=2.72, A=0, B=0), (40e-6, 40e-6, 30e-6)), ("modified", dict(k=3.1, m=1.6, A=0, B=0), (55e-6, 35e-6, 42e-6)),
             ("projected", dict(k=0, m=0, A=2.0, B=1.0), (50e-6, 50e-6, 30e-6)), ("projected", dict(k=0, m=0, A=0.7, B=0.3), (45e-6, 70e-6, 25e-6))]
    for kind, c, dims in specs:
        for depth in (dims[2], 1.9 * dims[2], 6.0 * dims[2]):          # transient beams deepen dimensions.z (heatSourceModel.C:207-213)
            pts = np.abs(rng.normal(0.0, 1.0, (40, 3))) * np.array([dims[0], dims[1], depth]) * 0.7
            pts[0] = 0.0                                                 # the beam centre
            pts[1] = [0.3 * dims[0], 0.2 * dims[1], 1.5 * depth]         # below the modifiedSuperGaussian's depth: weight 0
            pts[2] = [0.0, 0.0, depth]                                   # exactly at the depth
            out.append(dict(kind=kind, k=c["k"], m=c["m"], A=c["A"], B=c["B"], dims=list(dims), depth=depth, points=pts.tolist()))
    return out


def table_requests():
    rng = np.random.default_rng(7)
    out = [dict(x=[1410.0, 1620.0], y=[1.0, 0.0]), dict(x=[1620.0, 1500.0, 1410.0], y=[0.0, 0.55, 1.0]),      # shipped table; descending
           dict(x=[1.0, 0.0], y=[1410.0, 1620.0])]                                                              # roles swapped (createFields.H:59-71)
    for _ in range(4):
        n = int(rng.integers(2, 9))
        x = rng.uniform(1300.0, 1700.0, n)
        if rng.integers(0, 2):
            x = np.sort(x)
        out.append(dict(x=x.tolist(), y=rng.uniform(0.0, 1.0, n).tolist()))
    for t in out:
        lo, hi = min(t["x"]), max(t["x"])
        t["queries"] = [lo - 10.0, lo, hi, hi + 10.0] + t["x"] + rng.uniform(lo, hi, 12).tolist()
    return out


MATERIALS = {
    "IN625": dict(kappa=[[8.275, 0.01472, 0.0], [4.889, 0.014743, 0.0], [-0.07707, 0.00075, 0.0]],
                  Cp=[[579.28, 0.0, 0.0], [750.65, 0.0, 0.0], [747.568, 0.0, 0.0]]),
    "curved": dict(kappa=[[9.1, 0.012, 1.5e-6], [5.5, 0.013, -4e-7], [0.3, 0.0009, 2e-7]],
                   Cp=[[560.0, 0.11, 2e-5], [740.0, -0.02, 0.0], [330.0, 0.05, 1e-5]]),
}
TABLES = {"IN625": ([1410.0, 1620.0], [1.0, 0.0]), "curved": ([1650.0, 1580.0, 1490.0, 1400.0], [0.0, 0.35, 0.8, 1.0])}


def liquidus_solidus(name):
    """createFields.H:59-71 with the reference's own interpolateXY: Tliq = interpolateXY(0, y, x), Tsol = interpolateXY(1, y, x)."""
    x, y = TABLES[name]
    ans = ask([f"table {len(x)}"] + [f"{b!r} {a!r}" for a, b in zip(x, y)] + ["2", "0.0", "1.0"])
    return float.fromhex(ans[0]), float.fromhex(ans[1])


def cell_states(name, n=240):
    rng = np.random.default_rng(11 if name == "IN625" else 12)
    tliq, tsol = liquidus_solidus(name)
    T = rng.uniform(250.0, 3800.0, n)
    T[:n // 2] = rng.uniform(tsol - 120.0, tliq + 120.0, n // 2)          # half of them around the mushy zone
    T[:8] = [tsol, tliq, 300.0, 2.0 * tliq, np.nextafter(tsol, 0), np.nextafter(tliq, 1e9), 299.0, 2.0 * tliq + 1.0]
    T[8:8 + len(TABLES[name][0])] = TABLES[name][0]                          # the table nodes themselves
    a1 = rng.uniform(0.0, 1.0, n)
    a1[::5] = 1.0
    a1[1::7] = 0.0
    a1[2::11] = 1e-9            # below thermoTol 1e-8
    a1[3::13] = 1.0 - 1e-9
    a3 = (rng.uniform(0.0, 1.0, n) < 0.4).astype(float)
    return tliq, tsol, T.tolist(), a1.tolist(), a3.tolist()          # python floats: repr() must be a plain number


def ask(lines):
    assert not any("np." in ln for ln in lines), "numpy scalars print as np.float64(...): convert to float first"
    out = subprocess.run([REF], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    return [ln for ln in out if ln]


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    assert os.path.exists(REF), "oracle/_ref/refModels was not built: /root/reference is needed"
    shapes, tables = shape_requests(), table_requests()
    for s in shapes:
        lines = [f"shape {s['kind']} {s['k']!r} {s['m']!r} {s['A']!r} {s['B']!r} {s['dims'][0]!r} {s['dims'][1]!r} {s['dims'][2]!r} {s['depth']!r} {len(s['points'])}"]
        lines += [f"{p[0]!r} {p[1]!r} {p[2]!r}" for p in s["points"]]
        ans = ask(lines)
        assert ans[0].startswith("V0 ") and len(ans) == 1 + len(s["points"])
        s["V0"], s["weights"] = ans[0][3:], ans[1:]
    for t in tables:
        lines = [f"table {len(t['x'])}"] + [f"{a!r} {b!r}" for a, b in zip(t["x"], t["y"])] + [str(len(t["queries"]))] + [repr(q) for q in t["queries"]]
        t["values"] = ask(lines)
        assert len(t["values"]) == len(t["queries"])
    cells = []
    for name in MATERIALS:
        tliq, tsol, T, a1, a3 = cell_states(name)
        m = MATERIALS[name]
        coeffs = " ".join(repr(v) for row in m["kappa"] + m["Cp"] for v in row)
        ans = ask([f"props {len(T)} {tsol!r} {tliq!r}", coeffs] + [f"{t!r} {x!r} {y!r}" for t, x, y in zip(T, a1, a3)])
        rec = dict(material=name, Tliq=tliq.hex(), Tsol=tsol.hex(), T=T, alpha1=a1, alpha3=a3,
                   kappa=[r.split()[0] for r in ans], Cp=[r.split()[1] for r in ans], alpha3_after=[r.split()[2] for r in ans])
        x, y = TABLES[name]
        for tol in (1e-8, 1e-6):
            ans = ask([f"source {len(T)} {tliq!r} {tsol!r} {tol!r} {len(x)}"] + [f"{a!r} {b!r}" for a, b in zip(x, y)] + [f"{t!r} {v!r}" for t, v in zip(T, a1)])
            rec[f"dFdT_{tol:g}"], rec[f"T0_{tol:g}"] = [r.split()[0] for r in ans], [r.split()[1] for r in ans]
        cells.append(rec)
    # absorptionModels/Kelly/KellyAbsorption.C compiled into oracle/_ref/refKelly
    rng = np.random.default_rng(5)
    kelly = []
    for geometry, eta0, eta_min in (("cone", 0.28, 0.35), ("cylinder", 0.28, 0.35), ("cone", 0.41, 0.3), ("cylinder", 0.12, 0.2)):
        ratios = [0.3, 0.75, 1.0, float(np.nextafter(1.0, 2.0)), 1.0000001, 2.0, 7.5, 40.0] + rng.uniform(0.2, 12.0, 24).tolist()
        lines = [f"kelly {geometry} {eta0!r} {eta_min!r} {len(ratios)}"] + [repr(a) for a in ratios]
        assert not any("np." in ln for ln in lines)
        out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "refKelly")], input="\n".join(lines) + "\n", capture_output=True, text=True,
                             check=True).stdout.split()
        assert len(out) == len(ratios)
        kelly.append(dict(geometry=geometry, eta0=eta0, etaMin=eta_min, ratios=ratios, eta=out))
    with open(os.path.join(HERE, "ref_models.json"), "w") as f:
        json.dump(dict(shapes=shapes, tables=tables, cells=cells, kelly=kelly), f, indent=0)
    print("shapes", len(shapes), "tables", len(tables), "cell sets", len(cells), os.path.getsize(os.path.join(HERE, "ref_models.json")), "bytes")


if __name__ == "__main__":
    main()
