"""Regenerates tests/golden/scanpath/*  --  python tests/golden/make_scanpath_golden.py   (needs /root/reference)

These ARE outputs of the reference: oracle/_ref/createScanPath_ref is the reference's own scan-path geometry and writer
(/root/reference/applications/utilities/createScanPath/createFields.H, compiled where it lies by oracle/Makefile with
stand-ins for the two OpenFOAM includes it names) behind a main() that takes the createScanPathDict entries as arguments.
CASES[0] is the dictionary the reference ships (tutorials/multiLayerPBF/constant/createScanPathDict:18-34)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "scanpath")
REF = os.path.join(ROOT, "oracle", "_ref", "createScanPath_ref")

# name: (minPoint, maxPoint, angle, hatch, nRotations, power, speed, dwellTime, biDirection)
CASES = {
    "multiLayerPBF_tutorial": ((0.0, -1e-4), (0.002, 1e-4), 180.0, 1e-4, 2, 195.0, 0.8, 5e-4, True),
    "rot67_square": ((-1e-3, -1e-3), (1e-3, 1e-3), 67.0, 1.1e-4, 6, 285.0, 0.96, 2.5e-4, True),
    "unidirectional_rect": ((0.0, 0.0), (5e-3, 1.5e-3), 90.0, 7e-5, 3, 100.0, 1.2, 0.0, False),
    "rot45_offcentre": ((2.5e-4, -7.5e-4), (3.25e-3, 2.5e-4), 45.0, 8e-5, 4, 179.2, 0.8, 1e-4, True),
    "hatch_wider_than_box": ((0.0, 0.0), (1e-3, 1e-4), 30.0, 3e-4, 3, 50.0, 0.5, 1e-3, True),
    "degenerate_line_box": ((0.0, 0.0), (2e-3, 0.0), 15.0, 1e-4, 2, 120.0, 1.0, 2e-4, True),
}


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    assert os.path.exists(REF), "oracle/_ref/createScanPath_ref was not built: /root/reference is needed"
    os.makedirs(OUT, exist_ok=True)
    for name, (lo, hi, angle, hatch, nrot, power, speed, dwell, bidir) in CASES.items():
        prefix = os.path.join(OUT, name)
        args = [repr(v) for v in (lo[0], lo[1], hi[0], hi[1], angle, hatch)] + [str(nrot)] + [repr(v) for v in (power, speed, dwell)]
        subprocess.check_call([REF] + args + ["1" if bidir else "0", prefix])
        for i in range(nrot):
            os.replace(f"{prefix}_{i}", f"{prefix}_{i}.txt")
            print(f"{name}_{i}.txt", sum(1 for _ in open(f"{prefix}_{i}.txt")), "lines")


if __name__ == "__main__":
    sys.exit(main())
