"""Oracle parity of the sweep-kernel instantiations that the benchmark actually runs.

`Ldu::waveSweep` picks the kernel by tile count and environment: the three-role k_wave4 up to 600 tiles of 8x4 pencils, k_wave2 with
deep ring slots up to 1536, the many-tile schedule (k_wave2, ring 8 x 3) above that -- more than any other test reaches.  Here:

* a 64x400x124 block (1550 tiles, 3.2 M cells): factor + forward + backward sweep `array_equal` with the oracle's face loops
  ([UPSTREAM] DICPreconditioner / DILUPreconditioner::calcReciprocalD + precondition) and a PCG/DIC solve
  with the as-shipped controls (fvSolution:38-41) at the north-star tolerance -- this is the default schedule of the 100 M-cell
  benchmark (ticket order over thousands of tiles, ring reuse, sentinel re-arming at scale);
* a 64x400x60 block (750 tiles): the default between the two regimes;
* every selectable schedule -- `B200_WAVE=single` (k_wave2, one tile per CTA) with each ring geometry `B200_WAVE2_RING`,
  `B200_WAVE=multi` (k_wave3, several tiles per CTA exchanging edges through shared memory) with each `B200_WAVE3` geometry,
  `B200_WAVE=1warp` -- on small and ragged blocks (one tile, partial tiles, fewer tiles in J than a CTA holds, ...);
* the full time step on the thin slab shape of the 8-GPU decomposition, 1000x400x8, five steps against the oracle at 1e-9.
"""
import os

import numpy as np
import pytest

import helpers
import oracle_lib as ol
from additivefoam_b200 import abi, cases, lib
from test_gpu_ldu import oracle_precond, oracle_solve

pytestmark = pytest.mark.gpu


class env:
    """Set environment variables for the duration of a block (the library reads these knobs with getenv at every sweep)."""

    def __init__(self, **kw):
        self.kw = {k: v for k, v in kw.items() if v is not None}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def big_block(oracle, dims, seed, asym=False):
    n, nf, lo, up = ol.block_addressing(oracle, dims)
    rng = np.random.default_rng(seed)
    upper = -rng.uniform(0.5, 1.5, nf)
    lower = upper * rng.uniform(0.7, 1.3, nf) if asym else None
    diag = rng.uniform(0.1, 1.0, n)
    np.subtract.at(diag, lo, upper)
    np.subtract.at(diag, up, lower if asym else upper)
    return n, lo, up, diag, upper, lower, rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)


BIG = (64, 400, 124)          # 50 x 31 = 1550 tiles > 1536: the many-tile schedule of the 100 M-cell benchmark


@pytest.mark.parametrize("precond", [abi.PRECOND_DIC, abi.PRECOND_DILU])
def test_many_tile_schedule_sweeps_bit_exact(ctx, oracle, precond):
    asym = precond == abi.PRECOND_DILU
    n, lo, up, diag, upper, lower, x, b = big_block(oracle, BIG, 31, asym)
    s = lib.LduSolver(ctx, lo, up, n, block=BIG)
    got = [s.precondition(precond, b, diag, upper, lower) for _ in range(2)]       # the second call reuses the gathered coefficients
    ref = oracle_precond(oracle, n, lo, up, diag, upper, lower, precond, b)
    s.close()
    assert np.array_equal(got[0], ref) and np.array_equal(got[1], ref)


MID = (64, 400, 60)           # 50 x 15 = 750 tiles: between the thin-slab regime (k_wave4 up to 600 tiles) and the many-tile one: k_wave2, ring 8 x 6


@pytest.mark.parametrize("precond", [abi.PRECOND_DIC, abi.PRECOND_DILU])
def test_mid_tile_count_default_schedule_bit_exact(ctx, oracle, precond):
    """What a GPU gets of the 100 M-cell case on four GPUs (800 tiles) takes this default: factor + sweeps bit for bit, and a solve."""
    asym = precond == abi.PRECOND_DILU
    n, lo, up, diag, upper, lower, x0, b = big_block(oracle, MID, 41, asym)
    s = lib.LduSolver(ctx, lo, up, n, block=MID)
    got = s.precondition(precond, b, diag, upper, lower)
    assert np.array_equal(got, oracle_precond(oracle, n, lo, up, diag, upper, lower, precond, b))
    if not asym:
        got, perf = s.solve(x0, b, diag, upper, None, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC, tolerance=1e-15, relTol=0.0, minIter=1, maxIter=20)
        ref, rperf, _ = oracle_solve(oracle, n, lo, up, diag, upper, None, b, x0, abi.SOLVER_PCG, abi.PRECOND_DIC, 1e-15, 0.0, 1, 20)
        assert perf.nIterations == rperf.nIterations and helpers.rel_linf(got, ref) < 1e-9
    s.close()


def test_many_tile_schedule_pcg_dic_solve(ctx, oracle):
    """The benchmarked solve: PCG/DIC, tolerance 1e-15, maxIter 20 -> 21 iterations with wave-order vectors (k_wave_pcg_*)."""
    n, lo, up, diag, upper, lower, x0, b = big_block(oracle, BIG, 32)
    s = lib.LduSolver(ctx, lo, up, n, block=BIG)
    for _ in range(2):            # the second solve takes the look-ahead path of the host loop (lastIterations)
        got, perf = s.solve(x0, b, diag, upper, None, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC, tolerance=1e-15, relTol=0.0,
                            minIter=1, maxIter=20)
        ref, rperf, _ = oracle_solve(oracle, n, lo, up, diag, upper, None, b, x0, abi.SOLVER_PCG, abi.PRECOND_DIC, 1e-15, 0.0, 1, 20)
        assert perf.nIterations == rperf.nIterations == 21
        assert perf.initialResidual == pytest.approx(rperf.initialResidual, rel=1e-12)
        assert perf.finalResidual == pytest.approx(rperf.finalResidual, rel=1e-9)
        assert helpers.rel_linf(got, ref) < 1e-9
    s.close()


def test_many_tile_schedule_converging_solve(ctx, oracle):
    """Same block, a tolerance that is reached: iteration count and iterate against the oracle (the flag-look policy must not run past it)."""
    n, lo, up, diag, upper, lower, x0, b = big_block(oracle, BIG, 33)
    s = lib.LduSolver(ctx, lo, up, n, block=BIG)
    for _ in range(2):
        got, perf = s.solve(x0, b, diag, upper, None, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC, tolerance=1e-8, maxIter=200)
        ref, rperf, _ = oracle_solve(oracle, n, lo, up, diag, upper, None, b, x0, abi.SOLVER_PCG, abi.PRECOND_DIC, 1e-8, 0.0, 0, 200)
        assert perf.converged and abs(perf.nIterations - rperf.nIterations) <= 1
        assert helpers.rel_linf(got, ref) < 1e-9
    s.close()


SCHEDULES = [dict(B200_WAVE="single", B200_WAVE2_RING="83"), dict(B200_WAVE="single", B200_WAVE2_RING="82"),
             dict(B200_WAVE="single", B200_WAVE2_RING="44"), dict(B200_WAVE="single", B200_WAVE2_RING="86"),
             dict(B200_WAVE="single", B200_WAVE2_RING="84"), dict(B200_WAVE="single", B200_WAVE2_RING="412"),
             dict(B200_WAVE="single", B200_WAVE2_RING="48"), dict(B200_WAVE="tri", B200_WAVE2_RING="86"), dict(B200_WAVE="tri", B200_WAVE2_RING="83"),
             dict(B200_WAVE="tri", B200_WAVE2_RING="84"), dict(B200_WAVE="tri", B200_WAVE2_RING="48"), dict(B200_WAVE="tri", B200_WAVE2_RING="412"),
             dict(B200_WAVE="1warp"), dict(B200_WAVE="multi"),
             dict(B200_WAVE="multi", B200_WAVE3="444"), dict(B200_WAVE="multi", B200_WAVE3="248"), dict(B200_WAVE="multi", B200_WAVE3="183"),
             dict(B200_WAVE="multi", B200_WAVE3="146")]
SMALL = [(150, 25, 15), (33, 17, 9), (7, 5, 3), (300, 7, 2), (40, 70, 9), (1, 4, 3), (64, 33, 5)]


@pytest.mark.parametrize("sched", SCHEDULES, ids=lambda d: ",".join(f"{k[5:]}={v}" for k, v in d.items()))
@pytest.mark.parametrize("precond", [abi.PRECOND_DIC, abi.PRECOND_DILU])
def test_every_schedule_is_bit_exact_on_small_blocks(ctx, oracle, sched, precond):
    asym = precond == abi.PRECOND_DILU
    with env(**sched):
        for dims in SMALL:
            lo, up = helpers.block_addressing_np(dims)
            n = dims[0] * dims[1] * dims[2]
            diag, upper, lower, x, b = helpers.spd_coefficients(n, lo, up, seed=41, asym=asym)
            s = lib.LduSolver(ctx, lo, up, n, block=dims)
            got = s.precondition(precond, b, diag, upper, lower)
            ref = oracle_precond(oracle, n, lo, up, diag, upper, lower, precond, b)
            if not asym:
                xs, perf = s.solve(x, b, diag, upper, None, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC, tolerance=1e-10, maxIter=60)
                xr, rperf, _ = oracle_solve(oracle, n, lo, up, diag, upper, None, b, x, abi.SOLVER_PCG, abi.PRECOND_DIC, 1e-10, 0.0, 0, 60)
                assert abs(perf.nIterations - rperf.nIterations) <= 1 and helpers.rel_linf(xs, xr) < 1e-9, (sched, dims)
            s.close()
            assert np.array_equal(got, ref), (sched, dims)


@pytest.mark.parametrize("sched", [dict(B200_WAVE="single", B200_WAVE2_RING="83"), dict(B200_WAVE="single"), dict(B200_WAVE="multi"),
                                   dict(B200_WAVE="tri", B200_WAVE2_RING="83"), dict(B200_WAVE="tri", B200_WAVE2_RING="86"),
                                   dict(B200_WAVE="multi", B200_WAVE3="248")],
                         ids=lambda d: ",".join(f"{k[5:]}={v}" for k, v in d.items()))
def test_schedules_at_three_million_cells(ctx, oracle, sched):
    """The schedules that run at scale, on a block with enough tiles to recycle ring slots and tickets many times over."""
    dims = (200, 150, 100)
    n, lo, up, diag, upper, lower, x, b = big_block(oracle, dims, 34)
    with env(**sched):
        s = lib.LduSolver(ctx, lo, up, n, block=dims)
        got = s.precondition(abi.PRECOND_DIC, b, diag, upper, None)
        s.close()
    ref = oracle_precond(oracle, n, lo, up, diag, upper, None, abi.PRECOND_DIC, b)
    assert np.array_equal(got, ref)


def test_thin_slab_full_step_parity(ctx, oracle):
    """The per-GPU shape of the decomposed 100 M-cell run (1000 x 400 cells in plane, a few layers thick): five full time steps --
    properties, deltaT, heat source, assembly, correctors, PCG/DIC -- against the oracle at the north-star bar."""
    case = cases.synthetic_track((1000, 400, 8), deltaT=2e-5, explicitSolve=0, maxDi=10.0, solver=abi.SOLVER_PCG,
                                 preconditioner=abi.PRECOND_DIC)
    g, o = lib.Case(ctx, case), ol.OracleCase(case, lib=oracle, ranks=1)
    for s in range(5):
        rg, ro = g.step(), o.step()
        assert rg.nThermoCorr == ro.nThermoCorr and abs(rg.lastSolveIterations - ro.lastSolveIterations) <= 1, f"step {s}"
        assert rg.deltaT == pytest.approx(ro.deltaT, rel=1e-10)
        eT = helpers.rel_linf(g.field(abi.FIELD_T), o.field(abi.FIELD_T))
        eA = float(np.max(np.abs(g.field(abi.FIELD_ALPHA_SOLID) - o.field(abi.FIELD_ALPHA_SOLID))))
        assert eT < 1e-9 and eA < 1e-9, f"step {s}: T {eT:.3e} alpha {eA:.3e}"
    assert g.field(abi.FIELD_T).max() > 1000.0
    g.close(); o.close()
