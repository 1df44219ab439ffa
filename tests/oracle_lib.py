"""ctypes binding of the CPU oracle (oracle/libaf_oracle.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import numpy as np

from additivefoam_b200 import abi, cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "libaf_oracle.so")

dp = abi.c_double_p
ip = abi.c_int32_p


def _ensure_built():
    srcs = [os.path.join(ROOT, "oracle", f) for f in ("af_oracle_capi.cpp", "af_thermal.hpp", "af_beam.hpp", "af_ldu.hpp")]
    srcs.append(os.path.join(ROOT, "include", "additivefoam_b200.h"))
    if (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)


def load():
    _ensure_built()
    L = C.CDLL(_SO)
    L.afo_case_create.restype = C.c_void_p
    L.afo_case_create.argtypes = [C.POINTER(abi.CaseDesc), C.c_int]
    L.afo_case_destroy.argtypes = [C.c_void_p]
    L.afo_case_running.argtypes = [C.c_void_p]
    L.afo_case_step.argtypes = [C.c_void_p, C.POINTER(abi.StepReport)]
    L.afo_case_n_cells.restype = C.c_int64
    L.afo_case_n_cells.argtypes = [C.c_void_p]
    L.afo_case_get_field.argtypes = [C.c_void_p, C.c_int, dp]
    L.afo_case_set_field.argtypes = [C.c_void_p, C.c_int, dp]
    L.afo_case_residual_log.argtypes = [C.c_void_p, dp, C.c_int32]
    L.afo_case_melt_pool_dimensions.argtypes = [C.c_void_p, C.c_int32, dp, C.c_double, dp]
    L.afo_case_solidification_enable.argtypes = [C.c_void_p, dp, dp, C.c_double]
    L.afo_case_solidification_events.argtypes = [C.c_void_p, dp, C.c_int64]
    L.afo_case_solidification_events.restype = C.c_int64
    L.afo_case_assemble_probe.argtypes = [C.c_void_p, C.c_int, dp, dp, dp]
    L.afo_case_isotherm_depth.argtypes = [C.c_void_p, C.c_int, dp, dp]
    L.afo_case_sources_probe.argtypes = [C.c_void_p, C.c_double, dp]
    L.afo_case_pre_step_state.argtypes = [C.c_void_p, dp]
    L.afo_case_reset_alpha_solid.argtypes = [C.c_void_p]
    L.afo_case_time_scheme_state.argtypes = [C.c_void_p, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.afo_case_cn_snapshot.argtypes = [C.c_void_p, C.POINTER(C.c_int), dp, dp]
    L.afo_case_n_patches.argtypes = [C.c_void_p]
    L.afo_case_patch_size.restype = C.c_int64
    L.afo_case_patch_size.argtypes = [C.c_void_p, C.c_int]
    L.afo_case_patch_state.argtypes = [C.c_void_p, C.c_int, dp, dp]
    L.afo_mixed_value_fraction.restype = C.c_double
    L.afo_mixed_value_fraction.argtypes = [C.c_double] * 6
    for f in ("afo_case_time", "afo_case_Tliq", "afo_case_Tsol"):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [C.c_void_p]
    L.afo_absorption_eta.restype = C.c_double
    L.afo_absorption_eta.argtypes = [C.POINTER(abi.Beam), C.c_double]
    L.afo_shape_V0.restype = C.c_double
    L.afo_shape_V0.argtypes = [C.POINTER(abi.Beam), dp]
    L.afo_shape_weight.restype = C.c_double
    L.afo_shape_weight.argtypes = [C.POINTER(abi.Beam), dp, dp]
    L.afo_interpolateXY.restype = C.c_double
    L.afo_interpolateXY.argtypes = [C.c_double, C.c_int32, dp, dp]
    L.afo_beam_adjust_deltaT.restype = C.c_double
    L.afo_beam_adjust_deltaT.argtypes = [C.POINTER(abi.Beam), C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
    L.afo_beam_move.argtypes = [C.POINTER(abi.Beam), C.c_double, C.c_double, C.c_int32, dp, dp, dp, ip, dp]
    L.afo_beam_qdot.argtypes = [C.POINTER(abi.BlockMesh), C.POINTER(abi.Beam), dp, dp, C.c_double, dp, dp, dp, dp]
    L.afo_scan_path_parse.argtypes = [C.c_char_p, dp, C.c_int32]
    L.afo_block_addressing.argtypes = [C.POINTER(abi.BlockMesh), C.c_int, C.c_int, ip, ip, ip, ip]
    L.afo_ldu_amul.argtypes = [C.c_int32, C.c_int32, ip, ip, dp, dp, dp, dp, dp]
    L.afo_ldu_precondition.argtypes = [C.c_int32, C.c_int32, ip, ip, dp, dp, dp, C.c_int, dp, dp, dp]
    L.afo_ldu_solve.argtypes = [C.c_int32, C.c_int32, ip, ip, dp, dp, dp, dp, dp, C.c_int, C.c_int, C.c_double,
                                C.c_double, C.c_int32, C.c_int32, C.POINTER(abi.SolverPerf), dp, C.c_int32, ip]
    L.afo_ldu_levels.argtypes = [C.c_int32, C.c_int32, ip, ip, ip, ip]
    return L


def P(a):
    """numpy array -> typed pointer (None passes NULL)."""
    if a is None:
        return None
    if a.dtype == np.float64:
        return a.ctypes.data_as(dp)
    if a.dtype == np.int32:
        return a.ctypes.data_as(ip)
    raise TypeError(a.dtype)


class OracleCase:
    def __init__(self, case, ranks=1, lib=None):
        self.L = lib or load()
        self.built = cases.BuiltCase(case)
        self.h = self.L.afo_case_create(C.byref(self.built.desc), ranks)
        assert self.h, "oracle case creation failed"
        self.n = self.L.afo_case_n_cells(self.h)

    def running(self):
        return bool(self.L.afo_case_running(self.h))

    def step(self):
        rep = abi.StepReport()
        self.L.afo_case_step(self.h, C.byref(rep))
        return rep

    def field(self, f):
        out = np.empty(self.n, dtype=np.float64)
        assert self.L.afo_case_get_field(self.h, f, P(out)) == 0
        return out

    def set_field(self, f, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert self.L.afo_case_set_field(self.h, f, P(a)) == 0

    def restart(self, T, alpha_powder=None):
        if alpha_powder is not None:
            self.set_field(abi.FIELD_ALPHA_POWDER, alpha_powder)
        self.set_field(abi.FIELD_T, T)
        self.reset_alpha_solid()

    def reset_alpha_solid(self):
        assert self.L.afo_case_reset_alpha_solid(self.h) == 0

    def time_scheme_state(self):
        """(deltaT0 the next step's TEqn will see, old-time levels of T, of alpha.solid)."""
        dts, a, b = C.c_double(0.0), C.c_int(0), C.c_int(0)
        self.L.afo_case_time_scheme_state(self.h, C.byref(dts), C.byref(a), C.byref(b))
        return dts.value, a.value, b.value

    def cn_snapshot(self):
        """CrankNicolson: (state dict, ddt0(T), ddt0(alpha.solid)) as the last step's TEqn.H found them, right after runTime++."""
        st = (C.c_int * 9)()
        dT, dA = np.zeros(self.n), np.zeros(self.n)
        assert self.L.afo_case_cn_snapshot(self.h, st, P(dT), P(dA)) == 0
        keys = ("timeIndex", "TnOld", "a1nOld", "existsT", "startT", "tiT", "existsA", "startA", "tiA")
        return dict(zip(keys, [int(v) for v in st])), dT, dA

    def pre_step_state(self):
        """updateProperties() (so that kappa, Cp are what the next step's controls see) and T.oldTime()."""
        Told = np.empty(self.n, dtype=np.float64)
        assert self.L.afo_case_pre_step_state(self.h, P(Told)) == 0
        return Told

    def patch_state(self):
        """[(Tb, valueFraction)] of the physical patches of T, in the oracle's face order (single-rank cases)."""
        out = []
        for p in range(self.L.afo_case_n_patches(self.h)):
            n = self.L.afo_case_patch_size(self.h, p)
            Tb, vf = np.empty(n), np.empty(n)
            self.L.afo_case_patch_state(self.h, p, P(Tb), P(vf))
            out.append((Tb, vf))
        return out

    def sources_probe(self, proposed_dt):
        """adjustDeltaT + update of the sources over [time, time + dt], time += dt; returns the adjusted dt (qDot: field(FIELD_QDOT))."""
        dt = C.c_double(0.0)
        assert self.L.afo_case_sources_probe(self.h, float(proposed_dt), C.byref(dt)) == 0
        return dt.value

    def melt_pool_dimensions(self, iso_values, angle_deg=0.0):
        iso = np.ascontiguousarray(iso_values, dtype=np.float64)
        out = np.empty(3 * len(iso))
        self.L.afo_case_melt_pool_dimensions(self.h, len(iso), P(iso), angle_deg, P(out))
        return out.reshape(-1, 3)

    def solidification_enable(self, box_min, box_max, iso_value):
        lo, hi = np.asarray(box_min, dtype=np.float64), np.asarray(box_max, dtype=np.float64)
        self.L.afo_case_solidification_enable(self.h, P(lo), P(hi), iso_value)

    def solidification_events(self):
        n = self.L.afo_case_solidification_events(self.h, None, 0)
        out = np.empty((max(n, 1), 7))
        self.L.afo_case_solidification_events(self.h, P(out), n)
        return out[:n]

    def residual_log(self):
        out = np.empty(4096)
        n = self.L.afo_case_residual_log(self.h, P(out), 4096)
        return out[:n].copy()

    def close(self):
        if self.h:
            self.L.afo_case_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def block_addressing(L, n, r=0, R=1):
    bm = abi.BlockMesh()
    bm.nx, bm.ny, bm.nz = n
    for q in range(3):
        bm.min[q], bm.max[q] = 0.0, float(n[q])
    nc, nf = C.c_int32(), C.c_int32()
    L.afo_block_addressing(C.byref(bm), r, R, C.byref(nc), C.byref(nf), None, None)
    lo = np.empty(nf.value, dtype=np.int32)
    up = np.empty(nf.value, dtype=np.int32)
    L.afo_block_addressing(C.byref(bm), r, R, C.byref(nc), C.byref(nf), P(lo), P(up))
    return nc.value, nf.value, lo, up
