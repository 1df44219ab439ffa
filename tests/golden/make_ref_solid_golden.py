"""Regenerates tests/golden/ref_solid.json  --  python tests/golden/make_ref_solid_golden.py   (needs /root/reference)

These ARE outputs of the reference's code: oracle/_ref/refFO also holds the reference's own functionObjects/solidificationData/
solidificationData.C with its real header (constructor, setOverlapCells, execute, end), compiled where it lies against the
stand-ins of oracle/ref_stubs/fo.  The function object is a function of the temperature fields of consecutive time steps; those
come from running the CPU oracle on a small single-track case (tests/test_ref_solid.py re-runs the same case, so only the rows
the function object writes to data_0.csv are stored, as hex floats: x, y, z, time, cooling rate, |grad T|, their ratio)."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "refFO")
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

SIDES = ["xMin", "xMax", "yMin", "yMax", "zMin", "zMax"]


def configs():
    """name -> (case, steps, boxMin, boxMax, isoValue, patch kinds xMin..zMax).  Patches are zeroGradient or fixedValue so that the
    face values the gradient needs are known to both sides (the mixed top patch is pinned by refBC)."""
    from additivefoam_b200 import abi, cases
    out = {}
    for name, box, iso in (("whole_track", ([0.0, -1e-4, -2e-4], [2e-3, 1e-4, 0.0]), 1620.0),          # the shipped box (controlDict:62-69), touches the top
                           ("inner_box_solidus", ([2.1e-4, -3.1e-5, -9e-5], [6.05e-4, 5e-5, -1e-5]), 1410.0)):   # faces between cell layers: the 1e-10 growth decides
        c = cases.amb2018_02_b(n=(75, 13, 8), writeInterval=0.0, endTime=0.0012)
        c["beams"][0]["path"] = [[1, 0.0, 0.0, 0.0, 0.0, 0.0], [0, 0.0008, 0.0, 0.0, 179.2, 0.8]]
        c["patches"] = [cases._patch(abi.BC_FIXED_VALUE, [abi.ZMIN], value=300.0), cases._patch(abi.BC_ZERO_GRADIENT, [abi.ZMAX]),
                        cases._patch(abi.BC_ZERO_GRADIENT, [abi.XMIN, abi.XMAX, abi.YMIN, abi.YMAX])]
        out[name] = (c, 400, box[0], box[1], iso, ["z", "z", "z", "z", "f 300.0", "z"])
    return out


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    assert os.path.exists(REF), "oracle/_ref/refFO was not built: /root/reference is needed"
    import numpy as np
    import oracle_lib as ol
    from additivefoam_b200 import abi
    out = {}
    for name, (case, steps, lo, hi, iso, kinds) in configs().items():
        o = ol.OracleCase(case)
        o.solidification_enable(lo, hi, iso)
        m = case["mesh"]
        fmt = lambda a: " ".join(repr(float(v)) for v in a)
        req = ["solid %d %d %d %s %s  %s %s %r  %s  %d" % (*m["n"], fmt(m["min"]), fmt(m["max"]), fmt(lo), fmt(hi), iso, " ".join(kinds), steps), fmt(o.field(abi.FIELD_T))]
        for _ in range(steps):
            rep = o.step()
            req.append("%r %r %s" % (rep.time, rep.deltaT, fmt(o.field(abi.FIELD_T))))
        text = subprocess.run([REF], input="\n".join(req) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
        assert text[0] == "x,y,z,ts,cr,g,v"
        rows = [ln.split(",") for ln in text[1:] if ln]
        out[name] = rows
        mine = o.solidification_events()
        ref = np.array([[float.fromhex(v) for v in r] for r in rows]).reshape(-1, 7)
        print(name, "events:", len(rows), "oracle:", len(mine), "identical:", mine.shape == ref.shape and bool(np.array_equal(mine, ref)))
    with open(os.path.join(HERE, "ref_solid.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
