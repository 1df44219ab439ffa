"""ctypes binding of the CUDA library (``libadditivefoam_b200.so``) -- the product path.

No fallback: if the shared library is missing or no CUDA device is present, loading / context
creation raises.  The classes mirror the reference interfaces the library replaces:

* :class:`LduSolver`  -- ``Foam::lduMatrix::solver`` ([UPSTREAM OpenFOAM-10]; constructed from the LDU
  addressing, ``solve(psi, source)`` per call; ``solver``/``preconditioner``/``tolerance``/``relTol``/
  ``minIter``/``maxIter`` as in ``system/fvSolution``, tutorials/AMB2018-02-B/system/fvSolution:34-42)
* :func:`beam_qdot`   -- ``heatSourceModel::qDot()`` (movingHeatSource/heatSourceModels/heatSourceModel/heatSourceModel.C:221-371)
* :class:`Case`       -- the time-loop body of applications/solvers/additiveFoam/additiveFoam.C:66-95
"""
import ctypes as C
import os

import numpy as np

from . import abi, cases

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("B200_LIBRARY") or os.path.join(_HERE, "libadditivefoam_b200.so")

dp, ip = abi.c_double_p, abi.c_int32_p
REDUCE_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)      # b200_reduce_sum_fn

#: every symbol include/additivefoam_b200.h declares
SYMBOLS = [
    "b200_last_error", "b200_abi_version", "b200_device_count", "b200_nccl_unique_id", "b200_ctx_create", "b200_ctx_destroy",
    "b200_ctx_peer_export", "b200_ctx_peer_export_halo", "b200_ctx_peer_import", "b200_ctx_synchronize", "b200_ctx_mark", "b200_ctx_elapsed_ms", "b200_profile_enable", "b200_profile_groups",
    "b200_profile_name", "b200_profile_read", "b200_ldu_create", "b200_ldu_destroy", "b200_ldu_set_block", "b200_ldu_solve",
    "b200_ldu_solve_dev", "b200_ldu_amul", "b200_ldu_amul_dev", "b200_ldu_precondition", "b200_ldu_levels",
    "b200_beam_qdot", "b200_beam_qdot_sub", "b200_beam_shape_v0", "b200_beam_shape_weight", "b200_beam_absorption_eta", "b200_mixed_temperature_value_fraction", "b200_case_create", "b200_case_destroy", "b200_case_running", "b200_case_step",
    "b200_case_n_cells", "b200_case_get_field", "b200_case_set_field", "b200_case_reset_alpha_solid", "b200_case_set_field_async", "b200_case_apply_staged",
    "b200_case_get_field_async", "b200_case_transfers_wait", "b200_case_melt_pool_dimensions", "b200_case_isotherm_depth", "b200_case_sources_update", "b200_case_solidification_enable",
    "b200_case_solidification_events", "b200_selftest_division", "b200_case_kernel_launches",
    "b200_scan_path_parse", "b200_create_scan_path",
]

_lib = None


class B200Error(RuntimeError):
    pass


def _preload_nccl():
    """libadditivefoam_b200.so links libnccl.so.2 by soname.  PyTorch bundles a newer NCCL under the same soname; if ours
    pulled the system copy in first, a later ``import torch`` would bind to it and fail.  Load the bundled one first."""
    import glob
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
        roots = list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []
    except Exception:
        roots = []
    for r in roots:
        for so in glob.glob(os.path.join(r, "lib", "libnccl.so*")):
            try:
                C.CDLL(so, mode=C.RTLD_GLOBAL)
                return so
            except OSError:
                continue
    return None


def load():
    """Load the CUDA library; raises if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise B200Error(f"{SO_PATH} is missing: build it with additivefoam_b200/csrc/Makefile; there is no CPU fallback")
    _preload_nccl()
    L = C.CDLL(SO_PATH)
    L.b200_last_error.restype = C.c_char_p
    vp = C.c_void_p
    L.b200_nccl_unique_id.argtypes = [vp]
    L.b200_ctx_create.argtypes = [C.c_int, vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.b200_ctx_destroy.argtypes = [vp]
    L.b200_ctx_synchronize.argtypes = [vp]
    L.b200_ctx_peer_export.argtypes = [vp, vp]
    L.b200_ctx_peer_import.argtypes = [vp, vp]
    L.b200_ctx_peer_export_halo.argtypes = [vp, C.c_int64, vp]
    L.b200_ctx_mark.argtypes = [vp, C.c_int]
    L.b200_ctx_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, dp]
    L.b200_profile_enable.argtypes = [vp, C.c_int]
    L.b200_profile_name.restype = C.c_char_p
    L.b200_profile_name.argtypes = [C.c_int]
    L.b200_profile_read.argtypes = [vp, dp, C.POINTER(C.c_int64), C.c_int]
    L.b200_ldu_create.argtypes = [vp, C.c_int32, C.c_int32, ip, ip, C.c_int32, ip, ip, C.POINTER(ip), C.POINTER(vp)]
    L.b200_ldu_destroy.argtypes = [vp]
    L.b200_ldu_set_block.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
    L.b200_ldu_solve.argtypes = [vp, C.c_int32, C.c_int32, dp, dp, dp, C.POINTER(dp), dp, dp, C.c_double, C.c_double,
                                 C.c_int32, C.c_int32, C.POINTER(abi.SolverPerf)]
    L.b200_ldu_solve_dev.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp, C.c_double, C.c_double,
                                     C.c_int32, C.c_int32, C.POINTER(abi.SolverPerf)]
    L.b200_ldu_amul.argtypes = [vp, dp, dp, dp, C.POINTER(dp), dp, dp]
    L.b200_ldu_amul_dev.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.b200_ldu_precondition.argtypes = [vp, C.c_int32, dp, dp, dp, dp, dp]
    L.b200_ldu_levels.argtypes = [vp, ip, ip]
    L.b200_beam_qdot.argtypes = [vp, C.POINTER(abi.BlockMesh), C.POINTER(abi.Beam), dp, dp, C.c_double, dp, dp, dp, dp]
    L.b200_beam_qdot_sub.argtypes = [vp, C.POINTER(abi.BlockMesh), ip, ip, C.POINTER(abi.Beam), dp, dp, C.c_double, REDUCE_FN, vp, dp, dp, dp, dp]
    L.b200_beam_shape_v0.argtypes = [C.POINTER(abi.Beam), dp, dp, dp]
    L.b200_beam_shape_weight.argtypes = [C.POINTER(abi.Beam), dp, dp, dp]
    L.b200_beam_absorption_eta.argtypes = [C.POINTER(abi.Beam), C.c_double, dp]
    L.b200_mixed_temperature_value_fraction.argtypes = [C.c_double] * 6 + [dp]
    L.b200_case_create.argtypes = [vp, C.POINTER(abi.CaseDesc), C.POINTER(vp)]
    L.b200_case_destroy.argtypes = [vp]
    L.b200_case_running.argtypes = [vp]
    L.b200_case_step.argtypes = [vp, C.POINTER(abi.StepReport)]
    L.b200_case_n_cells.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.b200_case_get_field.argtypes = [vp, C.c_int32, dp]
    L.b200_case_set_field.argtypes = [vp, C.c_int32, dp]
    L.b200_case_set_field_async.argtypes = [vp, C.c_int32, dp]
    L.b200_case_get_field_async.argtypes = [vp, C.c_int32, dp]
    L.b200_case_apply_staged.argtypes = [vp]
    L.b200_case_transfers_wait.argtypes = [vp]
    L.b200_case_melt_pool_dimensions.argtypes = [vp, C.c_int32, dp, C.c_double, dp]
    L.b200_case_isotherm_depth.argtypes = [vp, C.c_int32, dp, dp]
    L.b200_case_sources_update.argtypes = [vp, C.c_double, dp]
    L.b200_case_solidification_enable.argtypes = [vp, dp, dp, C.c_double, C.c_int64]
    L.b200_case_solidification_events.argtypes = [vp, dp, C.c_int64, C.POINTER(C.c_int64)]
    L.b200_selftest_division.argtypes = [vp, C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_int64)]
    L.b200_case_kernel_launches.argtypes = [vp, C.POINTER(C.c_int64)]
    L.b200_scan_path_parse.argtypes = [C.c_char_p, dp, C.c_int32]
    L.b200_create_scan_path.argtypes = [dp, dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32,
                                        C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise B200Error(load().b200_last_error().decode())


def _P(a):
    if a is None:
        return None
    if a.dtype == np.float64:
        return a.ctypes.data_as(dp)
    if a.dtype == np.int32:
        return a.ctypes.data_as(ip)
    raise TypeError(a.dtype)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def nccl_unique_id():
    buf = (C.c_char * 128)()
    _check(load().b200_nccl_unique_id(buf))
    return bytes(buf)


class Context:
    """One per process = one GPU = one subdomain (the reference's MPI rank)."""

    def __init__(self, device=0, unique_id=None, rank=0, nranks=1):
        self.L = load()
        self.h = C.c_void_p()
        uid = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        _check(self.L.b200_ctx_create(device, uid, rank, nranks, C.byref(self.h)))
        self.rank, self.nranks, self.device = rank, nranks, device

    def synchronize(self):
        _check(self.L.b200_ctx_synchronize(self.h))

    def peer_export(self, halo_doubles=0):
        """64-byte IPC handle of this rank's all-reduce mailbox; with `halo_doubles` > 0 the allocation also carries the window
        the slab neighbours store their halo planes into (>= nx*ny of the largest case on this context)."""
        buf = (C.c_char * 64)()
        if halo_doubles > 0:
            _check(self.L.b200_ctx_peer_export_halo(self.h, int(halo_doubles), buf))
        else:
            _check(self.L.b200_ctx_peer_export(self.h, buf))
        return bytes(buf)

    def peer_import(self, handles):
        """`handles`: the 64-byte handles of all ranks, concatenated in rank order."""
        assert len(handles) == 64 * self.nranks
        _check(self.L.b200_ctx_peer_import(self.h, C.create_string_buffer(handles, len(handles))))

    def mark(self, i):
        """Record CUDA event ``i`` (0..3) on the library's stream."""
        _check(self.L.b200_ctx_mark(self.h, i))

    def elapsed_ms(self, i, j):
        ms = C.c_double()
        _check(self.L.b200_ctx_elapsed_ms(self.h, i, j, C.cast(C.byref(ms), dp)))
        return ms.value

    def profile_enable(self, on=True):
        _check(self.L.b200_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        """{group: (total_ms, launches)} since the last read; synchronises."""
        n = self.L.b200_profile_groups()
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        _check(self.L.b200_profile_read(self.h, ms, cnt, n))
        return {self.L.b200_profile_name(i).decode(): (ms[i], cnt[i]) for i in range(n)}

    def selftest_division(self, n_pairs, seed=1, exp_range=250):
        bad = C.c_int64()
        _check(self.L.b200_selftest_division(self.h, int(n_pairs), int(seed), int(exp_range), C.byref(bad)))
        return bad.value

    def kernel_launches(self):
        n = C.c_int64()
        self.L.b200_case_kernel_launches(None, C.byref(n))
        return n.value

    def close(self):
        if self.h:
            self.L.b200_ctx_destroy(self.h)
            self.h = C.c_void_p()


class LduSolver:
    """``lduMatrix::solver`` on the GPU: addressing at construction, coefficients per ``solve``."""

    def __init__(self, ctx, lower, upper, n_cells, block=None, solver=abi.SOLVER_PCG, preconditioner=abi.PRECOND_DIC,
                 tolerance=1e-6, relTol=0.0, minIter=0, maxIter=1000, interfaces=None):
        """``interfaces``: processor patches of a decomposed case as ``[(neighbour rank, faceCells), ...]``
        ([UPSTREAM] processorLduInterface::neighbProcNo(), lduAddressing::patchAddr(p)); the matching
        ``interfaceBouCoeffs`` go to ``solve`` / ``amul`` as ``bou_coeffs``."""
        self.ctx, self.L = ctx, ctx.L
        self.lower = np.ascontiguousarray(lower, dtype=np.int32)
        self.upper = np.ascontiguousarray(upper, dtype=np.int32)
        self.n, self.nf = int(n_cells), len(self.lower)
        self.h = C.c_void_p()
        self.face_cells = [np.ascontiguousarray(fc, dtype=np.int32) for _, fc in (interfaces or [])]
        ni = len(self.face_cells)
        ranks = np.array([r for r, _ in (interfaces or [])], dtype=np.int32)
        sizes = np.array([len(fc) for fc in self.face_cells], dtype=np.int32)
        ptrs = (ip * max(ni, 1))(*[_P(fc) for fc in self.face_cells])
        _check(self.L.b200_ldu_create(ctx.h, self.n, self.nf, _P(self.lower), _P(self.upper), ni, _P(ranks) if ni else None,
                                      _P(sizes) if ni else None, ptrs if ni else None, C.byref(self.h)))
        if block is not None:
            _check(self.L.b200_ldu_set_block(self.h, *block))
        self.controls = dict(solver=solver, preconditioner=preconditioner, tolerance=tolerance, relTol=relTol,
                             minIter=minIter, maxIter=maxIter)

    def _bou(self, bou_coeffs):
        if not self.face_cells:
            return None, None
        keep = [_f64(b) for b in bou_coeffs]
        assert [len(b) for b in keep] == [len(fc) for fc in self.face_cells], "one coefficient per interface face"
        return (dp * len(keep))(*[_P(b) for b in keep]), keep

    def solve(self, psi, source, diag, upper, lower=None, bou_coeffs=None, **kw):
        c = dict(self.controls)
        c.update(kw)
        psi = _f64(psi).copy()
        perf = abi.SolverPerf()
        diag, upper, source = _f64(diag), _f64(upper), _f64(source)
        lower = None if lower is None else _f64(lower)
        bou, _keep = self._bou(bou_coeffs)
        _check(self.L.b200_ldu_solve(self.h, c["solver"], c["preconditioner"], _P(diag), _P(upper), _P(lower), bou, _P(source),
                                     _P(psi), c["tolerance"], c["relTol"], c["minIter"], c["maxIter"], C.byref(perf)))
        return psi, perf

    def amul(self, psi, diag, upper, lower=None, bou_coeffs=None):
        psi, diag, upper = _f64(psi), _f64(diag), _f64(upper)
        lower = None if lower is None else _f64(lower)
        out = np.empty(self.n)
        bou, _keep = self._bou(bou_coeffs)
        _check(self.L.b200_ldu_amul(self.h, _P(diag), _P(upper), _P(lower), bou, _P(psi), _P(out)))
        return out

    def precondition(self, kind, rA, diag, upper, lower=None):
        rA, diag, upper = _f64(rA), _f64(diag), _f64(upper)
        lower = None if lower is None else _f64(lower)
        out = np.empty(self.n)
        _check(self.L.b200_ldu_precondition(self.h, kind, _P(diag), _P(upper), _P(lower), _P(rA), _P(out)))
        return out

    def levels(self):
        lv = np.empty(self.n, dtype=np.int32)
        nl = C.c_int32()
        _check(self.L.b200_ldu_levels(self.h, _P(lv), C.byref(nl)))
        return lv, nl.value

    def close(self):
        if self.h:
            self.L.b200_ldu_destroy(self.h)
            self.h = C.c_void_p()


def beam_qdot(ctx, mesh, beam, position, dimensions, power):
    """heatSourceModel::qDot() for one beam state; returns (qDot[nCells], sumWeights, volumeUsed, absorbedPower)."""
    L = ctx.L
    bm = abi.BlockMesh()
    bm.nx, bm.ny, bm.nz = mesh["n"]
    for q in range(3):
        bm.min[q], bm.max[q] = mesh["min"][q], mesh["max"][q]
    b = cases.make_beam(beam)
    n = bm.nx * bm.ny * bm.nz
    out = np.empty(n)
    pos, dims = _f64(position), _f64(dimensions)
    sw, vol, ap = C.c_double(), C.c_double(), C.c_double()
    _check(L.b200_beam_qdot(ctx.h, C.byref(bm), C.byref(b), _P(pos), _P(dims), power, _P(out),
                            C.cast(C.byref(sw), dp), C.cast(C.byref(vol), dp), C.cast(C.byref(ap), dp)))
    return out, sw.value, vol.value, ap.value


def beam_qdot_sub(ctx, mesh, offset, n_local, beam, position, dimensions, power, reduce_sum):
    """heatSourceModel::qDot() of one rank's sub-block (b200_beam_qdot_sub): `reduce_sum(local) -> global` plays the part of
    the reference's parallel reduction of sum(V*w).  Returns (qDot[local cells], sumWeights, volumeUsed, absorbedPower)."""
    L = ctx.L
    bm = abi.BlockMesh()
    bm.nx, bm.ny, bm.nz = mesh["n"]
    for q in range(3):
        bm.min[q], bm.max[q] = mesh["min"][q], mesh["max"][q]
    b = cases.make_beam(beam)
    off = np.ascontiguousarray(offset, dtype=np.int32)
    nl = np.ascontiguousarray(n_local, dtype=np.int32)
    out = np.empty(int(nl[0]) * int(nl[1]) * int(nl[2]))
    pos, dims = _f64(position), _f64(dimensions)
    sw, vol, ap = C.c_double(), C.c_double(), C.c_double()
    cb = REDUCE_FN(lambda v, user: float(reduce_sum(v)))
    _check(L.b200_beam_qdot_sub(ctx.h, C.byref(bm), off.ctypes.data_as(ip), nl.ctypes.data_as(ip), C.byref(b), _P(pos), _P(dims), power,
                                cb, None, _P(out), C.cast(C.byref(sw), dp), C.cast(C.byref(vol), dp), C.cast(C.byref(ap), dp)))
    return out, sw.value, vol.value, ap.value


class Case:
    """GPU-resident thermal case; ``step()`` is one pass of additiveFoam's ``while (runTime.run())`` body."""

    def __init__(self, ctx, case):
        self.ctx, self.L = ctx, ctx.L
        self.built = cases.BuiltCase(case)
        self.h = C.c_void_p()
        _check(self.L.b200_case_create(ctx.h, C.byref(self.built.desc), C.byref(self.h)))
        nl, ng = C.c_int64(), C.c_int64()
        _check(self.L.b200_case_n_cells(self.h, C.byref(nl), C.byref(ng)))
        self.n_local, self.n_global = nl.value, ng.value

    @classmethod
    def from_case_dir(cls, ctx, case_dir, **overrides):
        """A case straight from an AdditiveFOAM case directory (system/blockMeshDict, 0/T, constant/transportProperties,
        thermoPath, heatSourceDict + scan paths, system/controlDict, fvSolution, setFieldsDict): see ``foamcase.read_case``.
        ``overrides`` update the controls (e.g. ``explicitSolve=0, solver=abi.SOLVER_PCG``)."""
        from . import foamcase
        case = foamcase.read_case(case_dir)
        case["controls"].update(overrides)
        return cls(ctx, case)

    def running(self):
        return bool(self.L.b200_case_running(self.h))

    def step(self):
        rep = abi.StepReport()
        _check(self.L.b200_case_step(self.h, C.byref(rep)))
        return rep

    def field(self, f):
        out = np.empty(self.n_local)
        _check(self.L.b200_case_get_field(self.h, f, _P(out)))
        return out

    def set_field(self, f, a):
        a = _f64(a)
        assert a.size == self.n_local
        _check(self.L.b200_case_set_field(self.h, f, _P(a)))

    def reset_alpha_solid(self):
        """Restart: alpha.solid (and its old-time level) rebuilt from T through the thermoPath table, properties updated (createFields.H:31-44,74-78,112)."""
        _check(self.L.b200_case_reset_alpha_solid(self.h))

    def restart(self, T, alpha_powder=None):
        """Continue from fields of a time directory (foamcase.read_field): T and alpha.powder as read, alpha.solid rebuilt from T -- what
        the reference does when started from a written time (createFields.H:2-44,74-78)."""
        if alpha_powder is not None:
            self.set_field(abi.FIELD_ALPHA_POWDER, alpha_powder)
        self.set_field(abi.FIELD_T, T)
        self.reset_alpha_solid()

    def field_into(self, f, out):
        """get_field into an existing (e.g. pinned) array."""
        _check(self.L.b200_case_get_field(self.h, f, _P(out)))

    def set_field_keep_old(self, f, a):
        """set_field without touching the old-time copy (host application owns the field between steps)."""
        _check(self.L.b200_case_set_field(self.h, f | 0x100, _P(a)))

    def set_field_async(self, f, pinned):
        """Streamed upload from PINNED host memory (see include/additivefoam_b200.h); takes effect at apply_staged()."""
        assert pinned.size == self.n_local and pinned.dtype == np.float64
        _check(self.L.b200_case_set_field_async(self.h, f, _P(pinned)))

    def apply_staged(self):
        _check(self.L.b200_case_apply_staged(self.h))

    def get_field_async(self, f, pinned_out):
        assert pinned_out.size == self.n_local and pinned_out.dtype == np.float64
        _check(self.L.b200_case_get_field_async(self.h, f, _P(pinned_out)))

    def transfers_wait(self):
        _check(self.L.b200_case_transfers_wait(self.h))

    def melt_pool_dimensions(self, iso_values, angle_deg=0.0):
        """meltPoolDimensions function object (meltPoolDimensions.C:123-294): (nIso, 3) spans, computed on the device."""
        iso = np.ascontiguousarray(iso_values, dtype=np.float64)
        out = np.empty(3 * len(iso))
        _check(self.L.b200_case_melt_pool_dimensions(self.h, len(iso), _P(iso), float(angle_deg), _P(out)))
        return out.reshape(-1, 3)

    def isotherm_depth(self, beam, position):
        """heatSourceModel::updateDimensions of beam `beam` on the resident T with the beam centre at `position` (collective)."""
        pos, out = _f64(position), np.zeros(3)
        _check(self.L.b200_case_isotherm_depth(self.h, beam, _P(pos), _P(out)))
        return out

    def sources_update(self, proposed_dt):
        """movingHeatSourceModel::adjustDeltaT + ::update over [time, time + dt], then time += dt; returns the adjusted dt.
        qDot() is field(abi.FIELD_QDOT) (collective)."""
        dt = C.c_double(0.0)
        _check(self.L.b200_case_sources_update(self.h, float(proposed_dt), C.cast(C.byref(dt), dp)))
        return dt.value

    def solidification_enable(self, box_min, box_max, iso_value, max_events=1 << 20):
        """solidificationData function object (solidificationData.C): record cooling events on the device from the next step on."""
        lo, hi = np.asarray(box_min, dtype=np.float64), np.asarray(box_max, dtype=np.float64)
        self._solid_cap = int(max_events)
        _check(self.L.b200_case_solidification_enable(self.h, _P(lo), _P(hi), float(iso_value), self._solid_cap))

    def solidification_events(self):
        """(n, 7) rows x,y,z,ts,cr,g,v of this rank in the reference's order; raises if the device buffer overflowed."""
        n = C.c_int64()
        _check(self.L.b200_case_solidification_events(self.h, None, 0, C.byref(n)))
        if n.value > self._solid_cap:
            raise RuntimeError(f"solidificationData: {n.value} events exceed the capacity {self._solid_cap} given at enable")
        out = np.empty((max(n.value, 1), 7))
        _check(self.L.b200_case_solidification_events(self.h, _P(out), n.value, C.byref(n)))
        return out[:n.value]

    def close(self):
        if self.h:
            self.L.b200_case_destroy(self.h)
            self.h = C.c_void_p()


def beam_shape_V0(beam, dimensions):
    """heatSourceModel::V0() for the current dimensions (host only); returns (V0, k) -- k is the shape exponent, which
    projectedGaussian re-derives from the depth (projectedGaussian.C:96-106)."""
    b, d = cases.make_beam(beam), _f64(dimensions)
    v, k = C.c_double(), C.c_double()
    _check(load().b200_beam_shape_v0(C.byref(b), _P(d), C.cast(C.byref(v), dp), C.cast(C.byref(k), dp)))
    return v.value, k.value


def beam_shape_weight(beam, dimensions, distance):
    """heatSourceModel::weight(d) for a component-wise distance from the beam centre (host only)."""
    b, d, x = cases.make_beam(beam), _f64(dimensions), _f64(distance)
    w = C.c_double()
    _check(load().b200_beam_shape_weight(C.byref(b), _P(d), _P(x), C.cast(C.byref(w), dp)))
    return w.value


def beam_absorption_eta(beam, aspect_ratio):
    """absorptionModel::eta(aspectRatio) (host only)."""
    b = cases.make_beam(beam)
    e = C.c_double()
    _check(load().b200_beam_absorption_eta(C.byref(b), aspect_ratio, C.cast(C.byref(e), dp)))
    return e.value


def create_scan_path(min_point, max_point, hatch, angle, n_rotations, power, speed, dwell_time, bi_direction=True):
    """applications/utilities/createScanPath: the texts of constant/scanPath_0 .. scanPath_<nRotations-1>
    (createScanPath.C:74-101: file i is turned by i accumulated additions of ``angle`` degrees)."""
    L = load()
    lo, hi = _f64(min_point), _f64(max_point)
    out, rotation = [], 0.0
    for _ in range(n_rotations):
        n = C.c_int64()
        _check(L.b200_create_scan_path(_P(lo), _P(hi), hatch, rotation, power, speed, dwell_time, int(bool(bi_direction)), None, 0,
                                       C.byref(n)))
        buf = C.create_string_buffer(n.value + 1)
        _check(L.b200_create_scan_path(_P(lo), _P(hi), hatch, rotation, power, speed, dwell_time, int(bool(bi_direction)), buf,
                                       n.value + 1, C.byref(n)))
        out.append(buf.value.decode())
        rotation += angle
    return out


def scan_path_parse(text, max_segments=4096):
    out = np.empty(6 * max_segments)
    n = load().b200_scan_path_parse(text.encode(), _P(out), max_segments)
    if n < 0:
        raise B200Error("scan path has more segments than max_segments")
    return out[: 6 * n].reshape(n, 6).copy()


def mixed_temperature_value_fraction(h, emissivity, Tinf, Tp, kappa, delta_coeff):
    """mixedTemperatureFvPatchScalarField::updateCoeffs: the value fraction of one face (host only)."""
    f = C.c_double(0.0)
    _check(load().b200_mixed_temperature_value_fraction(h, emissivity, Tinf, Tp, kappa, delta_coeff, C.cast(C.byref(f), dp)))
    return f.value
