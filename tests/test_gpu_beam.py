"""GPU parity of heatSourceModel::qDot() (b200_beam_qdot) against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from additivefoam_b200 import abi, cases, lib

pytestmark = pytest.mark.gpu
P = ol.P

MESH = dict(n=[40, 30, 12], min=[-0.0002, -0.0003, -0.00024], max=[0.0006, 0.0003, 0.0])
PATH = [[1, 0, 0, 0, 0, 0], [0, 0.001, 0, 0, 100.0, 1.0]]


def oracle_qdot(L, mesh, beam, pos, dims, power):
    bm = abi.BlockMesh()
    bm.nx, bm.ny, bm.nz = mesh["n"]
    for q in range(3):
        bm.min[q], bm.max[q] = mesh["min"][q], mesh["max"][q]
    b = cases.make_beam(beam)
    n = bm.nx * bm.ny * bm.nz
    out = np.empty(n)
    pos, dims = np.array(pos, float), np.array(dims, float)
    sw, vol, ap = C.c_double(), C.c_double(), C.c_double()
    L.afo_beam_qdot(C.byref(bm), C.byref(b), P(pos), P(dims), power, P(out), C.cast(C.byref(sw), abi.c_double_p),
                    C.cast(C.byref(vol), abi.c_double_p), C.cast(C.byref(ap), abi.c_double_p))
    return out, sw.value, vol.value, ap.value


BEAMS = {
    "superGaussian": (cases.super_gaussian_beam(PATH), (85e-6, 85e-6, 30e-6)),
    "superGaussian_k4": (cases.super_gaussian_beam(PATH, k=4.2), (60e-6, 85e-6, 50e-6)),
    "modifiedSuperGaussian": (cases.modified_super_gaussian_beam(PATH), (40e-6, 40e-6, 30e-6)),
    "modifiedSuperGaussian_deep": (cases.modified_super_gaussian_beam(PATH), (40e-6, 40e-6, 95e-6)),
    "projectedGaussian": (cases.projected_gaussian_beam(PATH), (50e-6, 50e-6, 120e-6)),
}


@pytest.mark.parametrize("name", list(BEAMS))
@pytest.mark.parametrize("pos", [(1.234e-4, 1.7e-5, 0.0), (-0.00019, 0.00029, 0.0), (0.0003, 0.0, -0.0001), (0.01, 0.0, 0.0)])
def test_qdot_parity(ctx, oracle, name, pos):
    beam, dims = BEAMS[name]
    got, sw, vol, ap = lib.beam_qdot(ctx, MESH, beam, pos, dims, 150.0)
    ref, rsw, rvol, rap = oracle_qdot(oracle, MESH, beam, pos, dims, 150.0)
    assert ap == pytest.approx(rap, rel=1e-14)
    assert sw == pytest.approx(rsw, rel=1e-11, abs=1e-300)
    assert vol == pytest.approx(rvol, rel=1e-11, abs=1e-300)
    scale = max(np.abs(ref).max(), 1e-300)
    assert np.max(np.abs(got - ref)) / scale < 1e-11
    assert np.array_equal(got == 0.0, ref == 0.0)      # same support


def test_zero_power_is_zero_field(ctx, oracle):
    beam, dims = BEAMS["superGaussian"]
    got, sw, vol, ap = lib.beam_qdot(ctx, MESH, beam, (1e-4, 0, 0), dims, 0.0)
    assert not got.any() and ap == 0.0


def test_power_conservation(ctx):
    """Sum(V*qDot) == eta*P whenever the Riemann sum is within 5 % of V0 (heatSourceModel.C:359-367)."""
    beam, dims = BEAMS["superGaussian"]
    mesh = dict(MESH)
    got, sw, vol, ap = lib.beam_qdot(ctx, mesh, beam, (2e-4, 0.0, 0.0), dims, 179.2)
    V = np.prod([(mesh["max"][q] - mesh["min"][q]) / mesh["n"][q] for q in range(3)])
    assert ap == pytest.approx(0.33 * 179.2)
    if abs(1 - sw / vol) < 1e-12:          # normalised with the numerical volume
        assert got.sum() * V == pytest.approx(ap, rel=1e-12)


@pytest.mark.parametrize("name", ["superGaussian", "modifiedSuperGaussian_deep", "projectedGaussian"])
@pytest.mark.parametrize("split", [(2, 1, 1), (1, 3, 1), (1, 1, 2), (2, 3, 2)])
def test_qdot_of_a_decomposed_block_equals_the_serial_field(ctx, name, split):
    """b200_beam_qdot_sub, the per-rank form the OpenFOAM-side shim calls in a parallel run (heatSourceModel.C:287-353 loops over
    the rank's own cells, :359 sums V*w over all ranks): every sub-block of a `simple (nx ny nz)` decomposition of the block is
    evaluated on its own with the global sum handed in through the reduce callback; stitched together the pieces are the serial
    field bit for bit (same cell corners, same sample points, same normalisation), wherever the beam sits relative to the cuts."""
    beam, dims = BEAMS[name]
    n = MESH["n"]
    for pos in [(1.234e-4, 1.7e-5, 0.0), (0.0002, 0.0, 0.0), (0.0003, 0.0, -0.00012)]:
        ref, rsw, rvol, rap = lib.beam_qdot(ctx, MESH, beam, pos, dims, 150.0)
        cuts = [[n[a] * r // split[a] for r in range(split[a] + 1)] for a in range(3)]
        blocks = [((cuts[0][a], cuts[1][b], cuts[2][c]), (cuts[0][a + 1] - cuts[0][a], cuts[1][b + 1] - cuts[1][b], cuts[2][c + 1] - cuts[2][c]))
                  for c in range(split[2]) for b in range(split[1]) for a in range(split[0])]
        # pass 1: every rank's local sum (the callback sees it); pass 2: the same call with the global sum returned
        local = []
        for off, nl in blocks:
            lib.beam_qdot_sub(ctx, MESH, off, nl, beam, pos, dims, 150.0, lambda v: (local.append(v), v)[1])
        total = 0.0
        for v in local:            # rank order, as [UPSTREAM] gSum adds the partial sums
            total += v
        full = np.zeros(n[0] * n[1] * n[2]).reshape(n[2], n[1], n[0])
        for off, nl in blocks:
            q, sw, vol, ap = lib.beam_qdot_sub(ctx, MESH, off, nl, beam, pos, dims, 150.0, lambda v: total)
            full[off[2]:off[2] + nl[2], off[1]:off[1] + nl[1], off[0]:off[0] + nl[0]] = q.reshape(nl[2], nl[1], nl[0])
            assert ap == rap
        full = full.ravel()
        assert np.array_equal(full == 0.0, ref == 0.0)
        # the sum of the pieces' sums differs from the serial grid-wide sum by rounding only; with it the fields agree to that rounding
        assert total == pytest.approx(rsw, rel=1e-13)
        assert np.max(np.abs(full - ref)) <= 1e-13 * np.abs(ref).max()
