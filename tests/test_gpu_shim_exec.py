"""The GPU heat-source shim of the drop-in seam B2 (openfoam/b200HeatSource/b200HeatSources.{H,C}) EXECUTED, not only parsed
(VERDICT r1, weak 7): oracle/_ref/refMHSb200 is the reference's own movingHeatSourceModel.C, heatSourceModel.C and
heatSourceModelNew.C compiled where they lie (as in oracle/_ref/refMHS) with the shim compiled into the same stand-in world and the
product library linked in.  The requests of tests/golden/make_ref_mhs_golden.py are replayed with the `heatSourceModel` entries
b200SuperGaussian / b200ModifiedSuperGaussian / b200ProjectedGaussian, so

  * the reference's run-time selection code (heatSourceModelNew.C:43-63) finds the shim's types in its table and constructs them,
  * the reference's movingHeatSourceModel::update() (:100-150) sub-cycles the shim objects, whose qDot() goes through the C ABI
    (b200_beam_qdot_sub) to the CUDA kernels while V0() / weight() forward to the library's host mirror,
  * eta comes from the reference's own absorption models (constant and Kelly),

and the time-averaged qDot must be the stock models' (tests/golden/ref_mhs.json, outputs of the reference's code): same cells,
1e-11 relative (device exp/pow), the adjusted time step bit for bit."""
import json
import os
import sys
import tempfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_ref_mhs_golden as mk  # noqa: E402

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "oracle", "_ref", "refMHSb200")
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_mhs.json")))
H = float.fromhex


@pytest.mark.parametrize("i", range(len(mk.CONFIGS)))
def test_reference_update_through_the_gpu_shim(i):
    assert os.path.exists(EXE), "oracle/_ref/refMHSb200 is built by __graft_entry__.build() where /root/reference is present and travels with the repo"
    cfg, want = mk.CONFIGS[i], GOLD[i]["steps"]
    with tempfile.TemporaryDirectory() as tmp:
        got = mk.run(cfg, tmp, exe=EXE, kind_prefix="b200")
    assert len(got) == len(want)
    for s, (g, w) in enumerate(zip(got, want)):
        assert g["dt"] == w["dt"], f"step {s}: adjusted deltaT"
        assert g["cells"] == w["cells"], f"step {s}: cells touched"
        scale = max((abs(H(v)) for v in w["qDot"]), default=1.0)
        for c, a, b in zip(w["cells"], g["qDot"], w["qDot"]):
            assert abs(H(a) - H(b)) <= 1e-11 * scale, f"step {s} cell {c}: {H(a)} vs {H(b)}"


def test_shim_library_is_the_product_library():
    """The binary resolves the C ABI in the in-tree product library (rpath), not in a copy."""
    import subprocess
    out = subprocess.run(["ldd", EXE], capture_output=True, text=True).stdout
    line = [ln for ln in out.splitlines() if "libadditivefoam_b200.so" in ln]
    assert line and os.path.realpath(line[0].split("=>")[1].split()[0]) == os.path.realpath(os.path.join(ROOT, "additivefoam_b200", "libadditivefoam_b200.so"))


# ---------------------------------------------------------------------------------------------------------------------------
# B1: the lduMatrix::solver shim (openfoam/bPCG/bPCG.{H,C}) compiled against the executable lduMatrix stand-in of
# oracle/ref_stubs/ldu and linked against the product library (oracle/_ref/refBPCG): [UPSTREAM] lduMatrix::solver::New reads `solver`
# from the dictionary and finds bPCG / bPBiCGStab in the symmetric or asymmetric run-time selection table the shim's own
# add...ConstructorToTable objects filled; solve() goes through b200_ldu_solve to the CUDA kernels.  Checked against the oracle's
# restatement of the stock solvers on the same matrix.
BPCG = os.path.join(ROOT, "oracle", "_ref", "refBPCG")


def run_bpcg(field, n, lo, up, diag, upper, lower, b, x0, entries):
    import subprocess
    nf = len(lo)
    vec = lambda a: " ".join(repr(float(v)) for v in a)
    ints = lambda a: " ".join(str(int(v)) for v in a)
    req = ["bpcg %s %d %d %d %d %s" % (field, n, nf, 1 if lower is not None else 0, len(entries), " ".join("%s %s" % kv for kv in entries.items())),
           ints(lo), ints(up), vec(diag), vec(upper)]
    if lower is not None:
        req.append(vec(lower))
    req += [vec(b), vec(x0)]
    res = subprocess.run([BPCG], input="\n".join(req) + "\n", capture_output=True, text=True)
    return res.returncode, res.stdout.split("\n")


@pytest.mark.parametrize("solver,precond,asym,dims", [("bPCG", "DIC", False, (12, 9, 7)), ("bPCG", "diagonal", False, (12, 9, 7)),
                                                    ("bPCG", "DIC", "alloc", (10, 6, 5)), ("bPBiCGStab", "DILU", True, (12, 9, 7)),
                                                    ("bPBiCGStab", "none", True, (6, 5, 4))])
def test_solver_selected_at_run_time_solves_like_the_stock_solver(oracle, solver, precond, asym, dims):
    import numpy as np
    import helpers
    from additivefoam_b200 import abi
    from test_gpu_ldu import oracle_solve
    assert os.path.exists(BPCG), "oracle/_ref/refBPCG is built by __graft_entry__.build() and travels with the repo"
    lo, up = helpers.block_addressing_np(dims)
    n = dims[0] * dims[1] * dims[2]
    diag, upper, lower, x0, b = helpers.spd_coefficients(n, lo, up, seed=7, asym=asym is True)
    if asym == "alloc":
        lower = upper.copy()           # AdditiveFOAM's TEqn: both triangles allocated (fvm::div(phi, T)), numerically symmetric (finding F4)
    entries = dict(solver=solver, preconditioner=precond, tolerance="1e-10", relTol="0", maxIter="200")
    rc, out = run_bpcg("T", n, lo, up, diag, upper, lower, b, x0, entries)
    assert rc == 0, out
    name, its, r0, r1, conv = out[0].split()
    assert name == precond + solver and int(conv) == 1
    got = np.array([H(v) for v in out[1].split()])
    sid = abi.SOLVER_PCG if solver == "bPCG" else abi.SOLVER_PBICGSTAB
    pid = dict(DIC=abi.PRECOND_DIC, DILU=abi.PRECOND_DILU, diagonal=abi.PRECOND_DIAGONAL, none=abi.PRECOND_NONE)[precond]
    ref, rperf, _ = oracle_solve(oracle, n, lo, up, diag, upper, lower if asym is True else None, b, x0, sid, pid, 1e-10, 0.0, 0, 200)
    assert abs(int(its) - rperf.nIterations) <= 1
    assert H(r0) == pytest.approx(rperf.initialResidual, rel=1e-10)
    assert helpers.rel_linf(got, ref) < (1e-9 if solver == "bPCG" else 1e-7)


def test_unknown_solver_and_asymmetric_matrix_are_refused():
    import numpy as np
    import helpers
    dims = (5, 4, 3)
    lo, up = helpers.block_addressing_np(dims)
    n = 60
    diag, upper, lower, x0, b = helpers.spd_coefficients(n, lo, up, seed=3, asym=True)
    rc, out = run_bpcg("T", n, lo, up, diag, upper, lower, b, x0, dict(solver="PCG", preconditioner="DIC"))
    assert rc == 4 and "Unknown asymmetric matrix solver PCG" in out[0]        # the stock PCG is not in the asymmetric table: why bPCG registers in both
    rc, out = run_bpcg("T", n, lo, up, diag, upper, lower, b, x0, dict(solver="bPCG", preconditioner="DIC"))
    assert rc == 4 and "not symmetric" in out[0]                               # bPCG checks lower == upper before taking the PCG path
