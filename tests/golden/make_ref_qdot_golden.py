"""Regenerates tests/golden/ref_qdot.json  --  python tests/golden/make_ref_qdot_golden.py   (needs /root/reference)

These ARE outputs of the reference's code: oracle/_ref/refQDot is the reference's own heatSourceModel.C (constructor, qDot():
beam bounding box, per-cell sample-point count, the Riemann sum, the 5 % normalisation rule) with its real header and its
three shape sources, compiled where they lie against a stand-in uniform hex block mesh (oracle/ref_stubs/hsm, oracle/Makefile).
qDot is stored sparsely (cell index, hex float)."""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "refQDot")

# kind, k, m, A, B, dimensions, nPoints, eta, power, position, mesh n, mesh min, mesh max
CASES = [
    ("super", 2.0, 0, 0, 0, (85e-6, 85e-6, 30e-6), (10, 10, 10), 0.33, 179.2, (4.1e-4, 3.07e-4, 2.4e-4), (40, 30, 12), (0, 0, 0), (8e-4, 6e-4, 2.4e-4)),
    ("modified", 7.95, 2.72, 0, 0, (40e-6, 40e-6, 50e-6), (10, 10, 10), 0.47, 100.0, (3.9e-4, 2.9e-4, 2.4e-4), (40, 30, 12), (0, 0, 0), (8e-4, 6e-4, 2.4e-4)),
    ("projected", 0, 0, 2.0, 1.0, (50e-6, 50e-6, 80e-6), (10, 10, 10), 0.35, 100.0, (4.0e-4, 3.0e-4, 2.4e-4), (40, 30, 12), (0, 0, 0), (8e-4, 6e-4, 2.4e-4)),
    # non-cubic cells, few sample points, beam centre off the cell faces
    ("super", 3.3, 0, 0, 0, (70e-6, 55e-6, 35e-6), (3, 5, 2), 0.4, 250.0, (2.13e-4, 1.07e-4, 1.9e-4), (25, 14, 9), (0, -1e-4, -1e-5), (6.5e-4, 3.2e-4, 2.0e-4)),
    # beam at the corner of the block: most of the box is outside, the 5 % rule does not apply (deposited power < eta*P)
    ("super", 2.0, 0, 0, 0, (85e-6, 85e-6, 30e-6), (10, 10, 10), 0.33, 179.2, (1e-5, 2e-5, 2.4e-4), (20, 15, 12), (0, 0, 0), (4e-4, 3e-4, 2.4e-4)),
    # beam dimensions far below the cell size: one sample point per direction and cell
    ("projected", 0, 0, 1.0, 0.5, (6e-6, 6e-6, 9e-6), (10, 10, 10), 0.3, 60.0, (2.05e-4, 1.45e-4, 2.4e-4), (20, 15, 12), (0, 0, 0), (4e-4, 3e-4, 2.4e-4)),
    # unpowered
    ("modified", 5.0, 2.0, 0, 0, (40e-6, 40e-6, 30e-6), (4, 4, 4), 0.35, 0.0, (2e-4, 1.5e-4, 2.4e-4), (10, 8, 6), (0, 0, 0), (4e-4, 3e-4, 2.4e-4)),
]


# updateDimensions(): kind, k, m, A, B, static dimensions, isoValue, beam position, mesh n, min, max, hot-spot (centre, radii, peak)
DEPTH_CASES = [
    ("modified", 7.95, 2.72, 0, 0, (40e-6, 40e-6, 30e-6), 1620.0, (2.0e-4, 1.5e-4, 2.4e-4), (20, 15, 12), (0, 0, 0), (4e-4, 3e-4, 2.4e-4),
     ((2.07e-4, 1.46e-4, 2.4e-4), (70e-6, 60e-6, 95e-6), 3100.0)),
    ("projected", 0, 0, 2.0, 1.0, (50e-6, 50e-6, 30e-6), 1500.0, (2.2e-4, 1.4e-4, 2.4e-4), (20, 15, 12), (0, 0, 0), (4e-4, 3e-4, 2.4e-4),
     ((2.1e-4, 1.5e-4, 2.4e-4), (90e-6, 50e-6, 150e-6), 2600.0)),
    # the pool is shallower than the static depth: the depth stays the static one
    ("modified", 5.0, 2.0, 0, 0, (40e-6, 40e-6, 60e-6), 1620.0, (2.0e-4, 1.5e-4, 2.4e-4), (20, 15, 12), (0, 0, 0), (4e-4, 3e-4, 2.4e-4),
     ((2.0e-4, 1.5e-4, 2.4e-4), (60e-6, 60e-6, 25e-6), 2500.0)),
    # the isotherm lies outside the search radius max(dx, dy) around the beam axis
    ("modified", 5.0, 2.0, 0, 0, (30e-6, 30e-6, 20e-6), 1620.0, (0.8e-4, 1.5e-4, 2.4e-4), (20, 15, 12), (0, 0, 0), (4e-4, 3e-4, 2.4e-4),
     ((3.0e-4, 1.5e-4, 2.4e-4), (50e-6, 50e-6, 120e-6), 3000.0)),
    # non-cubic cells
    ("projected", 0, 0, 1.0, 0.5, (45e-6, 60e-6, 25e-6), 1550.0, (3.0e-4, 1.0e-4, 1.9e-4), (25, 14, 9), (0, -1e-4, -1e-5), (6.5e-4, 3.2e-4, 2.0e-4),
     ((3.05e-4, 1.1e-4, 1.9e-4), (80e-6, 70e-6, 110e-6), 2900.0)),
]


def hot_spot(n, lo, hi, spot):
    """T at the cell centres (cell order x fastest): 300 + (peak - 300) exp(-sum(((c - centre)/radius)^2)).  Plain Python floats."""
    import math
    (cx, cy, cz), (rx, ry, rz), peak = spot
    d = [(hi[q] - lo[q]) / n[q] for q in range(3)]
    T = []
    for k in range(n[2]):
        for j in range(n[1]):
            for i in range(n[0]):
                x, y, z = lo[0] + (i + 0.5) * d[0], lo[1] + (j + 0.5) * d[1], lo[2] + (k + 0.5) * d[2]
                T.append(300.0 + (peak - 300.0) * math.exp(-(((x - cx) / rx) ** 2 + ((y - cy) / ry) ** 2 + ((z - cz) / rz) ** 2)))
    return T


def depth_golden():
    out = []
    for kind, k, m, A, B, dims, iso, pos, n, lo, hi, spot in DEPTH_CASES:
        f = lambda v: repr(float(v))      # noqa: E731
        T = hot_spot(n, lo, hi, spot)
        lines = [" ".join(["depth", kind, f(k), f(m), f(A), f(B)] + [f(v) for v in dims] + [f(iso)] + [f(v) for v in pos] + [str(v) for v in n]
                          + [f(v) for v in lo] + [f(v) for v in hi])] + [repr(t) for t in T]
        ans = subprocess.run([REF], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split()
        assert len(ans) == 3
        out.append(dict(kind=kind, k=k, m=m, A=A, B=B, dims=list(dims), isoValue=iso, position=list(pos), n=list(n), min=list(lo), max=list(hi),
                        spot=[list(spot[0]), list(spot[1]), spot[2]], dimensions=ans))
        print("depth", kind, [float.fromhex(v) for v in ans], "static", dims[2])
    return out


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    assert os.path.exists(REF), "oracle/_ref/refQDot was not built: /root/reference is needed"
    out = []
    for kind, k, m, A, B, dims, npts, eta, power, pos, n, lo, hi in CASES:
        f = lambda v: repr(float(v))      # noqa: E731
        line = " ".join(["qdot", kind, f(k), f(m), f(A), f(B)] + [f(v) for v in dims] + [f(dims[2])] + [str(v) for v in npts] + [f(eta), f(power)]
                        + [f(v) for v in pos] + [str(v) for v in n] + [f(v) for v in lo] + [f(v) for v in hi])
        ans = subprocess.run([REF], input=line + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
        ans = [a for a in ans if a]
        assert ans[0].startswith("sum ") and len(ans) == 1 + n[0] * n[1] * n[2]
        nz = [(i, v) for i, v in enumerate(ans[1:]) if float.fromhex(v) != 0.0]
        out.append(dict(kind=kind, k=k, m=m, A=A, B=B, dims=list(dims), nPoints=list(npts), eta=eta, power=power, position=list(pos), n=list(n),
                        min=list(lo), max=list(hi), power_deposited=ans[0][4:], cells=[i for i, _ in nz], qdot=[v for _, v in nz]))
        print(kind, n, "nonzero cells", len(nz), "deposited", float.fromhex(ans[0][4:]), "eta*P", eta * power)
    with open(os.path.join(HERE, "ref_qdot.json"), "w") as fjson:
        json.dump(dict(qdot=out, depth=depth_golden()), fjson, indent=0)
    print(os.path.getsize(os.path.join(HERE, "ref_qdot.json")), "bytes")


if __name__ == "__main__":
    main()
