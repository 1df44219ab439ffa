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


def ask(lines):
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
    with open(os.path.join(HERE, "ref_models.json"), "w") as f:
        json.dump(dict(shapes=shapes, tables=tables), f, indent=0)
    print("shapes", len(shapes), "tables", len(tables), os.path.getsize(os.path.join(HERE, "ref_models.json")), "bytes")


if __name__ == "__main__":
    main()
