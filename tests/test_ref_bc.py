"""mixedTemperatureFvPatchScalarField::updateCoeffs (derivedFvPatchFields/mixedTemperature/mixedTemperatureFvPatchScalarField.C:
157-188) against values computed by the REFERENCE's own code (tests/golden/ref_bc.json, made by
tests/golden/make_ref_bc_golden.py from oracle/_ref/refBC = the reference's boundary-condition source with its real header,
compiled where it lies): 6 patches x 40 faces -- the shipped top patch, radiation only, convection only, neither (the 1e-15 floor of
hEff), a strong film, a hot environment; patch temperatures from 250 K to 3500 K incl. exactly Tinf, Tsol, Tliq.  Bit for bit:
the oracle's restatement and the library's host mirror, which is the function k_side_coeffs evaluates on the device (the device
path itself is held to the oracle in every case test: all of them have the mixed top patch)."""
import json
import os

import pytest

from additivefoam_b200 import lib

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_bc.json")))
H = float.fromhex


def rows(g):
    return zip(map(H, g["Tp"]), map(H, g["kappa"]), map(H, g["deltaCoeffs"]), map(H, g["valueFraction"]), map(H, g["refValue"]), map(H, g["refGradient"]))


@pytest.mark.parametrize("i", range(len(GOLD)))
def test_oracle_value_fraction_matches_the_reference_code(oracle, i):
    g = GOLD[i]
    for Tp, kappa, delta, f, refValue, refGrad in rows(g):
        assert oracle.afo_mixed_value_fraction(g["h"], g["emissivity"], g["Tinf"], Tp, kappa, delta) == f
        assert refValue == g["Tinf"] and refGrad == 0.0                  # what the restatements assume (evaluate: f Tinf + (1 - f) Tc)


@pytest.mark.parametrize("i", range(len(GOLD)))
def test_library_value_fraction_matches_the_reference_code(i):
    g = GOLD[i]
    for Tp, kappa, delta, f, _, _ in rows(g):
        assert lib.mixed_temperature_value_fraction(g["h"], g["emissivity"], g["Tinf"], Tp, kappa, delta) == f


def test_fixture_reaches_the_floor_and_both_limits():
    fs = [H(v) for g in GOLD for v in g["valueFraction"]]
    assert min(fs) < 1e-18 and max(fs) > 0.5                             # hEff = 0 -> 1e-15 floor; strong film -> nearly fixedValue
