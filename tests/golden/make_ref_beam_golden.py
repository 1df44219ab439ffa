"""Regenerates tests/golden/ref_beam.json  --  python tests/golden/make_ref_beam_golden.py   (needs /root/reference)

These ARE outputs of the reference's code: oracle/_ref/refBeam is the reference's own movingBeam.C and segment.C (scan-path
reading, path times, findIndex, move, activePath, adjustDeltaT) compiled where they lie against a stand-in header for the
OpenFOAM types they use (oracle/ref_stubs/beam/fvCFD.H, oracle/Makefile).  Values are stored as hex floats; the scan-path
texts are stored with them so that the tests rebuild the same beams."""
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "refBeam")

PATHS = {
    "AMB2018": "Mode    X       Y       Z   Power   Param\n1       0.000   0.000   0   0       0\n0       0.002   0.000   0   179.2   0.8\n",
    "raster": ("Mode\tX(m)\tY(m)\tZ(m)\tPower(W)\ttParam\n1\t0.0\t-0.0001\t0\t0\t0\n0\t0.002\t-0.0001\t0\t195.0\t0.8\n"
               "1\t0.002\t0.0\t0\t0\t0.0005\n0\t0.0\t0.0\t0\t195.0\t0.8\n1\t0.0\t0.0001\t0\t0\t0.0005\n0\t0.002\t0.0001\t0\t195.0\t0.8\n"),
    "awkward": ("header\n1 0 0 0 0 0\n\n0 1e-3 0 0 100 0.5\n1 1e-3 0 0 0 0\n1 1e-3 2e-4 0 50 3e-4\n0 1e-3 2e-4 0 80 1.0\n"      # zero dwell, dwell with power,
                "0 0 2e-4 -1e-4 120 0.25\n1 0 0 0 0 1e-4\n0 5e-4 5e-4 0 0 2.0\n0 1e-3 5e-4 0 60 0.4\n"),                       # a zero-length line, 3-D and unpowered moves
}
CONFIGS = [("AMB2018", -1.0, 1, 0.0, 0.004), ("raster", -1.0, 1, 0.0, 0.02), ("raster", 2e-6, 0, 0.0, 0.006), ("awkward", -1.0, 1, 0.0, 0.05),
           ("awkward", 5e-6, 1, 1.2e-3, 0.004)]


def ask(lines):
    assert not any("np." in ln for ln in lines), "numpy scalars print as np.float64(...): convert to float first"
    out = subprocess.run([REF], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    return [ln for ln in out if ln]


def times_for(text, start, end, rng):
    """A forward sweep, every segment boundary and its neighbours within a few eps (1e-10, movingBeam.C:32), and jumps back."""
    rows = [[float(v) for v in ln.split()] for ln in text.splitlines()[1:] if ln.strip()]
    t, bounds, prev = 0.0, [], (0.0, 0.0, 0.0)
    for mode, x, y, z, _, par in rows:
        t += par if mode == 1 else (np.sqrt((x - prev[0]) ** 2 + (y - prev[1]) ** 2 + (z - prev[2]) ** 2) / par)
        prev = (x, y, z)
        bounds.append(float(t))
    last = max(bounds[-1], start) * 1.15 + 1e-4
    sweep = np.sort(rng.uniform(start, last, 60)).tolist()
    near = [b + d for b in bounds for d in (-3e-10, -5e-11, 0.0, 5e-11, 3e-10) if b + d >= 0]
    forward = sorted(sweep + near)
    back = [forward[len(forward) // 2], forward[3], forward[-1], forward[len(forward) // 3], start]
    return [float(t) for t in forward + back]


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    assert os.path.exists(REF), "oracle/_ref/refBeam was not built: /root/reference is needed"
    rng = np.random.default_rng(99)
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        for name, deltaT, hit, start, end in CONFIGS:
            path = os.path.join(tmp, name)
            with open(path, "w") as f:
                f.write(PATHS[name])
            times = times_for(PATHS[name], start, end, rng)
            ans = ask([f"move {deltaT!r} {hit} {start!r} {end!r} {path} {len(times)}"] + [repr(t) for t in times])
            assert len(ans) == len(times)
            rec = dict(path=name, deltaT=deltaT, hitPathIntervals=hit, start=start, end=end, times=times, moves=[a.split() for a in ans], adjust=[])
            for _ in range(40):
                now = float(rng.choice(times[: len(times) - 5]))
                dt = float(10 ** rng.uniform(-7.5, -3.5))
                move_to = now if rng.integers(0, 4) else -1.0
                a = ask([f"adjust {deltaT!r} {hit} {start!r} {end!r} {path} {move_to!r} {now!r} {dt!r}"])
                rec["adjust"].append(dict(moveTo=move_to, now=now, dt=dt, result=a[0]))
            out.append(rec)
    with open(os.path.join(HERE, "ref_beam.json"), "w") as f:
        json.dump(dict(paths=PATHS, configs=out), f, indent=0)
    print("configs", len(out), os.path.getsize(os.path.join(HERE, "ref_beam.json")), "bytes")


if __name__ == "__main__":
    main()
