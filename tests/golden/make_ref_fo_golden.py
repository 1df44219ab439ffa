"""Regenerates tests/golden/ref_fo.json  --  python tests/golden/make_ref_fo_golden.py   (needs /root/reference)

These ARE outputs of the reference's code: oracle/_ref/refFO is the reference's own functionObjects/meltPoolDimensions/
meltPoolDimensions.C with its real header, compiled where it lies against stand-in headers (oracle/ref_stubs/fo,
oracle/ref_stubs/hsm: one uniform hex block with six physical patches; oracle/Makefile).  Per configuration: a hot-spot
temperature field (tests/golden/make_ref_qdot_golden.hot_spot, recomputed by the tests from its parameters), the patch kinds,
the isotherms and the scan-path angle -> the (length, width, depth) the function object logs, as hex floats."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "refFO")
sys.path.insert(0, HERE)
from make_ref_qdot_golden import hot_spot  # noqa: E402

ZG = ["z"] * 6
BLOCK = dict(n=[40, 20, 10], min=[-2e-4, -2e-4, -2e-4], max=[6e-4, 2e-4, 0.0])                 # 20 um cubes
FLAT = dict(n=[30, 16, 9], min=[-1.5e-4, -2.4e-4, -1.8e-4], max=[6e-4, 2.4e-4, 0.0])           # 25 x 30 x 20 um cells
ISO = [1620.0, 1410.0, 800.0, 1.0e5]

# (mesh, patches xMin xMax yMin yMax zMin zMax, angle, spot ((centre), (radii), peak))
CONFIGS = [
    (BLOCK, ZG, 0.0, ((2e-4, 0.0, 0.0), (1.2e-4, 6e-5, 5e-5), 3000.0)),                       # a pool open to the top surface
    (BLOCK, ZG, 30.0, ((2e-4, 0.0, 0.0), (1.2e-4, 6e-5, 5e-5), 3000.0)),                      # the same, rotated
    (BLOCK, ZG, -67.0, ((2.3e-4, 3e-5, -4e-5), (9e-5, 7e-5, 3e-5), 2500.0)),                  # buried centre, negative angle
    (FLAT, ZG, 15.0, ((2e-4, 1e-5, 0.0), (1.1e-4, 8e-5, 6e-5), 2800.0)),                      # non-cubic cells
    (BLOCK, ["f 2000.0", "z", "z", "z", "z", "z"], 0.0, ((2e-4, 0.0, 0.0), (1.2e-4, 6e-5, 5e-5), 3000.0)),   # a hot wall: every face of xMin counts
    (BLOCK, ["z", "z", "z", "z", "z", "f 300.0"], 20.0, ((2e-4, 0.0, 0.0), (1.2e-4, 6e-5, 5e-5), 3000.0)),   # a cold top: max(cell, face) takes the cell
    (BLOCK, ZG, 0.0, ((-1.9e-4, -1.9e-4, -1e-4), (8e-5, 8e-5, 8e-5), 2600.0)),               # a pool in the corner: three walls involved
    (BLOCK, ZG, 45.0, ((2e-4, 0.0, 0.0), (1.2e-4, 6e-5, 5e-5), 700.0)),                       # nothing reaches any isotherm: all zero
]


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    assert os.path.exists(REF), "oracle/_ref/refFO was not built: /root/reference is needed"
    out = []
    for mesh, patches, angle, spot in CONFIGS:
        T = hot_spot(mesh["n"], mesh["min"], mesh["max"], spot)
        req = ["meltpool %d %d %d %r %r %r %r %r %r  0.001 %r  %d %s" % (*mesh["n"], *mesh["min"], *mesh["max"], angle, len(ISO), " ".join(repr(v) for v in ISO)),
               " ".join(patches), " ".join(repr(v) for v in T)]
        res = subprocess.run([REF], input="\n".join(req) + "\n", capture_output=True, text=True, check=True).stdout.split()
        assert len(res) == len(ISO)
        dims = [r.split(",")[1:] for r in res]
        out.append(dict(mesh=mesh, patches=patches, angle=angle, spot=spot, iso=ISO, dimensions=dims))
        print(angle, patches, [[float.fromhex(v) for v in d] for d in dims][:3])
    with open(os.path.join(HERE, "ref_fo.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
