"""b200_create_scan_path against scan paths written by the REFERENCE's own generator (tests/golden/scanpath, made by
tests/golden/make_scanpath_golden.py from oracle/_ref/createScanPath_ref = the reference's createFields.H compiled where it
lies).  Host code on both sides: no GPU needed.  Byte equality, including the sign of a printed -0.0000000000."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_scanpath_golden as mk  # noqa: E402

from additivefoam_b200 import lib  # noqa: E402

GOLD = os.path.join(HERE, "golden", "scanpath")


def golden(name, i):
    with open(os.path.join(GOLD, f"{name}_{i}.txt")) as f:
        return f.read()


@pytest.mark.parametrize("name", list(mk.CASES))
def test_scan_path_bytes_equal_reference(name):
    lo, hi, angle, hatch, nrot, power, speed, dwell, bidir = mk.CASES[name]
    texts = lib.create_scan_path(lo, hi, hatch, angle, nrot, power, speed, dwell, bidir)
    assert len(texts) == nrot
    for i, t in enumerate(texts):
        assert t == golden(name, i), f"{name} rotation {i}"


def test_generated_path_feeds_the_beam_reader():
    """createScanPath -> movingBeam::readPath: the tutorial's dictionary gives three bidirectional 2 mm lines at 195 W."""
    lo, hi, angle, hatch, nrot, power, speed, dwell, bidir = mk.CASES["multiLayerPBF_tutorial"]
    seg = lib.scan_path_parse(lib.create_scan_path(lo, hi, hatch, angle, nrot, power, speed, dwell, bidir)[0])
    assert seg.shape == (6, 6)
    assert list(seg[:, 0]) == [1, 0, 1, 0, 1, 0]
    assert np.allclose(seg[1::2, 4], 195.0) and np.allclose(seg[1::2, 5], 0.8)
    assert list(seg[0::2, 5]) == [0.0, 5e-4, 5e-4]                      # no dwell before the first line
    assert np.allclose(seg[:, 2], [-1e-4, -1e-4, 0, 0, 1e-4, 1e-4], atol=1e-15)
    assert np.allclose(seg[:, 1], [0, 2e-3, 2e-3, 0, 0, 2e-3], atol=1e-15)   # odd line reversed


def test_bad_arguments_are_refused():
    with pytest.raises(lib.B200Error):
        lib.create_scan_path((0, 0), (1e-3, 1e-3), 0.0, 90, 1, 100, 1, 0)          # hatch 0: the reference would loop forever
    with pytest.raises(lib.B200Error):
        lib.create_scan_path((1e-3, 0), (0, 1e-3), 1e-4, 90, 1, 100, 1, 0)


@pytest.mark.skipif(not os.path.exists("/root/reference/applications/utilities/createScanPath/createFields.H"),
                    reason="the reference tree is only present in the build container")
def test_golden_files_are_what_the_reference_writes(tmp_path):
    """Where the reference sources are present, rebuild its generator and check the committed files are its output."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    for name, (lo, hi, angle, hatch, nrot, power, speed, dwell, bidir) in mk.CASES.items():
        prefix = str(tmp_path / name)
        args = [repr(v) for v in (lo[0], lo[1], hi[0], hi[1], angle, hatch)] + [str(nrot)] + [repr(v) for v in (power, speed, dwell)]
        subprocess.check_call([mk.REF] + args + ["1" if bidir else "0", prefix])
        for i in range(nrot):
            assert open(f"{prefix}_{i}").read() == golden(name, i)
