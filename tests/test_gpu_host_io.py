"""Host I/O of the resident step through the C ABI (include/additivefoam_b200.h): the blocking b200_case_set_field / get_field and the
copy-stream form set_field_async / apply_staged / get_field_async / transfers_wait that bench.py's `e2e` figure is measured through.
The fields of a step that make the round trip device -> pinned host buffer -> device must come back bit for bit, so a run whose every
step goes through the host equals the resident run exactly, and both stay within the bound of the oracle."""
import numpy as np
import pytest

from additivefoam_b200 import abi, cases

pytestmark = pytest.mark.gpu
IMPLICIT = dict(explicitSolve=0, maxDi=10.0, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC, writeInterval=0.0)


def _pinned(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory().numpy()


def test_async_upload_and_download_equal_the_blocking_calls(ctx):
    from additivefoam_b200 import lib
    case = cases.amb2018_02_b(n=(40, 10, 6), **IMPLICIT)
    a, b = lib.Case(ctx, case), lib.Case(ctx, case)
    rng = np.random.default_rng(11)
    T = _pinned(rng.uniform(300.0, 1800.0, a.n_local))
    P = _pinned((rng.uniform(0, 1, a.n_local) > 0.7).astype(np.float64))
    a.set_field(abi.FIELD_ALPHA_POWDER, P); a.set_field_keep_old(abi.FIELD_T, T)
    b.set_field_async(abi.FIELD_ALPHA_POWDER, P); b.set_field_async(abi.FIELD_T, T)
    assert not np.array_equal(b.field(abi.FIELD_T), T)               # staged, not applied yet
    b.apply_staged()
    out = _pinned(np.zeros(a.n_local))
    b.get_field_async(abi.FIELD_T, out)
    b.transfers_wait()
    assert np.array_equal(out, T) and np.array_equal(a.field(abi.FIELD_T), T)
    assert np.array_equal(a.field(abi.FIELD_T_OLD), b.field(abi.FIELD_T_OLD))          # upload-only: the old-time level is not touched
    for _ in range(5):
        ra, rb = a.step(), b.step()
    assert np.array_equal(a.field(abi.FIELD_T), b.field(abi.FIELD_T)) and ra.lastSolveIterations == rb.lastSolveIterations
    a.close(); b.close()


def test_two_cases_alternating_through_host_buffers(ctx, oracle):
    """The loop of bench.py's e2e figure: two different cases share the context; while one steps, the other's results go to its pinned
    buffers and come back as its next inputs (the upload is ordered after the download of the same field on the device, no host wait
    in between).  Each case must equal its own resident run bit for bit, and the oracle within the north-star bound."""
    import oracle_lib as ol
    from additivefoam_b200 import lib
    descs = [cases.amb2018_02_b(n=(60, 12, 8), **IMPLICIT),
             cases.multi_layer_pbf(n=(80, 16, 8), cell=20e-6, hatch=20e-6, nHatch=2, endTime=1.0, writeInterval=0.0)]
    through_host = [lib.Case(ctx, d) for d in descs]
    resident = [lib.Case(ctx, d) for d in descs]
    oracles = [ol.OracleCase(d, lib=oracle) for d in descs]
    F = (abi.FIELD_T, abi.FIELD_ALPHA_SOLID, abi.FIELD_ALPHA_POWDER)
    host = [{f: _pinned(g.field(f)) for f in F} for g in through_host]

    def out_and_back(q):
        g, h = through_host[q], host[q]
        g.get_field_async(abi.FIELD_T, h[abi.FIELD_T]); g.get_field_async(abi.FIELD_ALPHA_SOLID, h[abi.FIELD_ALPHA_SOLID])
        for f in (abi.FIELD_ALPHA_POWDER, abi.FIELD_T, abi.FIELD_ALPHA_SOLID):
            g.set_field_async(f, h[f])

    nsteps = [0, 0]
    for i in range(60):
        x, y = i & 1, 1 - (i & 1)
        if i > 0:
            out_and_back(y)
        through_host[x].apply_staged()
        through_host[x].step()
        nsteps[x] += 1
    for g in through_host:
        g.transfers_wait()
    for q in range(2):
        for _ in range(nsteps[q]):
            resident[q].step(); oracles[q].step()
        Th, Tr, To = through_host[q].field(abi.FIELD_T), resident[q].field(abi.FIELD_T), oracles[q].field(abi.FIELD_T)
        assert np.array_equal(Th, Tr), q
        assert np.array_equal(through_host[q].field(abi.FIELD_ALPHA_SOLID), resident[q].field(abi.FIELD_ALPHA_SOLID)), q
        assert float(np.max(np.abs(Th - To)) / np.max(np.abs(To))) < 1e-9
        assert To.max() > 400.0
    # case 0 was downloaded after its last step (during case 1's): its host buffer holds its final T
    assert np.array_equal(host[0][abi.FIELD_T], through_host[0].field(abi.FIELD_T))
    for g in through_host + resident:
        g.close()
    for o in oracles:
        o.close()
