"""Regenerates tests/golden/ref_mhs.json  --  python tests/golden/make_ref_mhs_golden.py   (needs /root/reference)

These ARE outputs of the reference's code: oracle/_ref/refMHS is the reference's own movingHeatSourceModel.C (adjustDeltaT, update:
each beam sub-cycled over the time step with its own deltaT, the time-weighted mean of qDot, the sum over beams) with its real
header, on top of its heatSourceModel.C, its three shapes, and its movingBeam.C / segment.C, compiled where they lie against
stand-in headers (oracle/ref_stubs/hsm, oracle/ref_stubs/beam, oracle/Makefile).  Per configuration: a few consecutive time
steps (proposed deltaT -> adjusted deltaT, qDot).  qDot is stored sparsely (cell, hex float); scan paths as their rows."""
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "refMHS")

MESH = dict(n=[48, 24, 10], min=[-2e-4, -2.4e-4, -2e-4], max=[7.6e-4, 2.4e-4, 0.0])          # 20 um cubes

TRACK = [[1, 0.0, 0.0, 0.0, 0.0, 0.0], [0, 5e-4, 0.0, 0.0, 179.2, 0.8]]
BACK = [[1, 5e-4, 1e-4, 0.0, 0.0, 0.0], [0, 0.0, 1e-4, 0.0, 150.0, 1.0]]
RASTER = [[1, 0.0, -1e-4, 0.0, 0.0, 0.0], [0, 4e-4, -1e-4, 0.0, 195.0, 0.8], [1, 4e-4, 0.0, 0.0, 0.0, 1e-4], [0, 0.0, 0.0, 0.0, 195.0, 0.8]]
LATE = [[1, 1e-4, -1e-4, 0.0, 0.0, 2e-4], [0, 4e-4, -1e-4, 0.0, 120.0, 0.5]]                   # powered only after a 200 us dwell


def beam(name, kind, path, dims, nPoints=(5, 5, 5), k=2.0, m=0.0, A=0.0, B=0.0, eta=0.33, deltaT=0.0, hit=1, kelly=None):
    """kelly = (geometry, eta0, etaMin): the reference's own KellyAbsorption.C decides eta from the beam's aspect ratio."""
    return dict(name=name, kind=kind, path=path, dims=list(dims), nPoints=list(nPoints), k=k, m=m, A=A, B=B, eta=eta, deltaT=deltaT, hit=hit, kelly=kelly)


CONFIGS = [
    # one beam, no beam deltaT (one sub-step per time step)
    dict(start=0.0, end=1e-3, beams=[beam("beam1", "super", TRACK, (85e-6, 85e-6, 30e-6))], steps=[1e-5, 2e-5, 4e-5]),
    # one beam sub-cycled: beam deltaT 3 us against time steps of 10-25 us (a ragged last sub-step)
    dict(start=0.0, end=1e-3, beams=[beam("beam1", "super", TRACK, (85e-6, 85e-6, 30e-6), deltaT=3e-6)], steps=[1e-5, 2.5e-5, 1.7e-5]),
    # two beams of different shapes and sub-cycle lengths, opposite directions; adjustDeltaT sees both paths
    dict(start=0.0, end=1e-3, beams=[beam("beam1", "super", TRACK, (85e-6, 85e-6, 30e-6), deltaT=4e-6),
                                     beam("beam2", "modified", BACK, (40e-6, 40e-6, 60e-6), k=7.95, m=2.72, eta=0.4, deltaT=7e-6)],
         steps=[2e-5, 3e-5, 3e-5]),
    # a raster path crossed with hitPathIntervals on: the time step is cut at the end of the first line, then the dwell (power 0),
    # then the return line; second beam starts late (inactive contribution 0 while it dwells unpowered)
    dict(start=4.6e-4, end=2e-3, beams=[beam("beam1", "projected", RASTER, (50e-6, 50e-6, 60e-6), A=2.0, B=1.0, eta=0.35, deltaT=5e-6),
                                        beam("beam2", "super", LATE, (60e-6, 60e-6, 40e-6), k=3.0, eta=0.3)],
         steps=[5e-5, 5e-5, 8e-5, 8e-5, 5e-5]),
    # hitPathIntervals off: the step straddles the end of the line; the sub-steps after the beam switches off add nothing
    dict(start=4.6e-4, end=2e-3, beams=[beam("beam1", "super", RASTER, (85e-6, 85e-6, 30e-6), deltaT=6e-6, hit=0)], steps=[6e-5, 6e-5, 9e-5]),
    # Kelly absorption through qDot(): eta from dimensions.z / min(dimensions.x, dimensions.y) (heatSourceModel.C:249-257) -- a cone of
    # aspect ratio 2.25, a cylinder whose x and y differ (ratio 2, not 1.2), and a shallow one (ratio < 1: etaMin)
    dict(start=0.0, end=1e-3, beams=[beam("beam1", "modified", TRACK, (40e-6, 40e-6, 90e-6), k=7.95, m=2.72, deltaT=5e-6, kelly=("cone", 0.28, 0.35)),
                                     beam("beam2", "super", BACK, (50e-6, 30e-6, 60e-6), k=2.0, kelly=("cylinder", 0.3, 0.2))], steps=[2e-5, 2e-5]),
    dict(start=0.0, end=1e-3, beams=[beam("beam1", "projected", TRACK, (60e-6, 80e-6, 45e-6), A=2.0, B=1.0, kelly=("cone", 0.28, 0.35))], steps=[1.5e-5, 1.5e-5]),
    # a time step that ends after the whole path: activePath() switches the beam off for the following step
    dict(start=5.9e-4, end=2e-3, beams=[beam("beam1", "super", TRACK, (85e-6, 85e-6, 30e-6), deltaT=1e-5, hit=0)], steps=[3e-5, 3e-5, 3e-5]),
]


def path_text(rows):
    return "Mode X Y Z Power Param\n" + "".join(" ".join(repr(float(v)) if i else str(int(v)) for i, v in enumerate(r)) + "\n" for r in rows)


def run(cfg, tmp, exe=REF, kind_prefix=""):
    """exe / kind_prefix: tests/test_gpu_shim_exec.py replays the same requests through oracle/_ref/refMHSb200 with the GPU shim's types."""
    m = MESH
    lines = ["mhs %d %d %d %r %r %r %r %r %r  %r %r  %d %d" % (*m["n"], *m["min"], *m["max"], cfg["start"], cfg["end"], len(cfg["beams"]), len(cfg["steps"]))]
    for b in cfg["beams"]:
        f = os.path.join(tmp, b["name"] + ".path")
        with open(f, "w") as fh:
            fh.write(path_text(b["path"]))
        lines.append("%s %s %r %r %r %r  %r %r %r  %d %d %d  %r  %r %d %s" % (b["name"], kind_prefix + b["kind"], b["k"], b["m"], b["A"], b["B"], *b["dims"], *b["nPoints"],
                                                                             b["eta"], b["deltaT"] if b["deltaT"] > 0 else -1.0, b["hit"], f))
        if b["kelly"]:
            lines[-1] = lines[-1].replace("  %r  " % b["eta"], "  kelly:%s:%r:%r  " % tuple(b["kelly"]), 1)
    lines += [repr(dt) for dt in cfg["steps"]]
    res = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True)
    assert res.returncode == 0, (res.returncode, res.stderr[-2000:])
    out = res.stdout.split("\n")
    out = [ln for ln in out if ln]
    steps, i = [], 0
    for _ in cfg["steps"]:
        assert out[i].startswith("dt ") and out[i + 1].startswith("n ")
        n = int(out[i + 1].split()[1])
        rows = [ln.split() for ln in out[i + 2:i + 2 + n]]
        steps.append(dict(dt=out[i].split()[1], cells=[int(r[0]) for r in rows], qDot=[r[1] for r in rows]))
        i += 2 + n
    assert i == len(out)
    return steps


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    assert os.path.exists(REF), "oracle/_ref/refMHS was not built: /root/reference is needed"
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        for cfg in CONFIGS:
            out.append(dict(mesh=MESH, start=cfg["start"], end=cfg["end"], beams=cfg["beams"], proposed=cfg["steps"], steps=run(cfg, tmp)))
    with open(os.path.join(HERE, "ref_mhs.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    for c in out:
        print(len(c["beams"]), "beams:", [(float.fromhex(s["dt"]), len(s["cells"])) for s in c["steps"]])


if __name__ == "__main__":
    main()
