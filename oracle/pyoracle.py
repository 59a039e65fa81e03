"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE — never imported by nyx_b200).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs use it.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from nyx_b200 import abi

_DIR = Path(__file__).resolve().parent
_LIB = None
_SPEED = None


def build(force: bool = False) -> Path:
    so = _DIR / "libnyx_oracle.so"
    srcs = [_DIR / "nyx_oracle.c", _DIR / "nyx_oracle_od.c", _DIR / "nyx_oracle_mvn.c", _DIR / "nyx_oracle.h", _DIR / "nyx_oracle_priv.h"]
    so2 = _DIR / "libnyx_oracle_speed.so"
    if force or not so.exists() or not so2.exists() or min(so.stat().st_mtime, so2.stat().st_mtime) < max(f.stat().st_mtime for f in srcs):
        subprocess.run(["make", "-C", str(_DIR), "-B" if force else "-s"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(str(so))
        vp = C.c_void_p
        L.nyx_oracle_propagate_batch.restype = C.c_int
        L.nyx_oracle_propagate_batch.argtypes = [
            C.POINTER(abi.DynamicsC), C.POINTER(abi.IntegOpts), C.c_size_t, vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp, C.c_int]
        L.nyx_oracle_propagate_batch_traj.restype = C.c_int
        L.nyx_oracle_propagate_batch_traj.argtypes = [
            C.POINTER(abi.DynamicsC), C.POINTER(abi.IntegOpts), C.c_size_t, vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp,
            C.POINTER(abi.TrajSink), C.c_int]
        L.nyx_oracle_propagate_batch_event.restype = C.c_int
        L.nyx_oracle_propagate_batch_event.argtypes = [
            C.POINTER(abi.DynamicsC), C.POINTER(abi.IntegOpts), C.c_size_t, vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp,
            C.POINTER(abi.TrajSink), C.POINTER(abi.EventC), C.c_int]
        L.nyx_oracle_dur_to_seconds.restype = C.c_double
        L.nyx_oracle_dur_to_seconds.argtypes = [C.c_int64]
        L.nyx_oracle_dur_from_seconds.restype = C.c_int64
        L.nyx_oracle_dur_from_seconds.argtypes = [C.c_double]
        L.nyx_oracle_sincos.restype = None
        L.nyx_oracle_sincos.argtypes = [C.c_double, abi.c_double_p, abi.c_double_p]
        L.nyx_oracle_rotation.restype = None
        L.nyx_oracle_rotation.argtypes = [C.POINTER(abi.Rotation), C.c_int64, abi.c_double_p, abi.c_double_p]
        L.nyx_oracle_body_position.restype = C.c_int
        L.nyx_oracle_body_position.argtypes = [C.POINTER(abi.BodyC), C.c_int64, abi.c_double_p]
        L.nyx_oracle_occultation.restype = C.c_double
        L.nyx_oracle_occultation.argtypes = [abi.c_double_p, abi.c_double_p, C.c_double, C.c_double]
        L.nyx_oracle_error_estimate.restype = C.c_double
        L.nyx_oracle_error_estimate.argtypes = [C.c_int, abi.c_double_p, abi.c_double_p, abi.c_double_p]
        L.nyx_oracle_eom.restype = C.c_int
        L.nyx_oracle_eom.argtypes = [C.POINTER(abi.DynamicsC), C.c_int64, C.c_double, abi.c_double_p, abi.c_double_p, abi.c_double_p]
        L.nyx_oracle_tableau.restype = C.c_int
        L.nyx_oracle_tableau.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(abi.c_double_p), C.POINTER(abi.c_double_p)]
        L.nyx_oracle_num_threads.restype = C.c_int
        L.nyx_oracle_set_error_scale.restype = None
        L.nyx_oracle_set_error_scale.argtypes = [C.c_double]
        # STM path (nyx_oracle_od.c)
        L.nyx_oracle_dual_eom.restype = C.c_int
        L.nyx_oracle_dual_eom.argtypes = [C.POINTER(abi.DynamicsC), C.c_int64, abi.c_double_p, abi.c_double_p, abi.c_double_p, abi.c_double_p]
        L.nyx_oracle_propagate_batch_stm.restype = C.c_int
        L.nyx_oracle_propagate_batch_stm.argtypes = [
            C.POINTER(abi.DynamicsC), C.POINTER(abi.IntegOpts), C.c_size_t, vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp, C.c_int]
        L.nyx_oracle_inst_new.restype = vp
        L.nyx_oracle_inst_new.argtypes = [C.POINTER(abi.DynamicsC), C.POINTER(abi.IntegOpts), abi.c_double_p, abi.c_double_p, C.c_int64]
        L.nyx_oracle_inst_free.restype = None
        L.nyx_oracle_inst_free.argtypes = [vp]
        L.nyx_oracle_inst_for_duration.restype = C.c_int
        L.nyx_oracle_inst_for_duration.argtypes = [vp, C.c_int64]
        L.nyx_oracle_inst_get.restype = None
        L.nyx_oracle_inst_get.argtypes = [vp, abi.c_double_p, abi.c_int64_p, abi.c_int64_p, C.POINTER(C.c_int), vp]
        L.nyx_oracle_inst_set.restype = None
        L.nyx_oracle_inst_set.argtypes = [vp, abi.c_double_p, C.c_int64]
        L.nyx_oracle_inst_set_step.restype = None
        L.nyx_oracle_inst_set_step.argtypes = [vp, C.c_int64, C.c_int]
        L.nyx_oracle_philox4x32_10.restype = None
        L.nyx_oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.nyx_oracle_mvn_sample.restype = C.c_int
        L.nyx_oracle_mvn_sample.argtypes = [C.c_uint64, C.c_uint64, C.c_size_t, abi.c_double_p, abi.c_double_p, abi.c_double_p, vp, vp]
        _LIB = L
    return _LIB


def speed_lib():
    """The -O3 -march=x86-64-v3 -ffp-contract=fast build of the same sources (timing / sensitivity only, never the checker)."""
    global _SPEED
    if _SPEED is None:
        build()
        L = C.CDLL(str(_DIR / "libnyx_oracle_speed.so"))
        vp = C.c_void_p
        L.nyx_oracle_propagate_batch_event.restype = C.c_int
        L.nyx_oracle_propagate_batch_event.argtypes = [
            C.POINTER(abi.DynamicsC), C.POINTER(abi.IntegOpts), C.c_size_t, vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp,
            C.POINTER(abi.TrajSink), C.POINTER(abi.EventC), C.c_int]
        _SPEED = L
    return _SPEED


def set_error_scale(scale: float = 1.0):
    """Sensitivity probe of the parity build: every adaptive error norm is multiplied by `scale` (1.0 restores the restatement)."""
    lib().nyx_oracle_set_error_scale(float(scale))


def propagate_batch(dyn_c, opts_c, state_soa, consts_soa, epoch0_ns, end_epoch_ns, step_ns=None, n_threads=0, traj_capacity=0,
                    event=None, speed_build=False):
    """Same contract as nyxb_propagate_batch[_traj] (include/nyxb.h) but on the CPU oracle.
    With traj_capacity > 0 returns a fifth element (epochs[cap][n], states[6][cap][n], count[n]); with
    event=(kind, value, trigger) the stop condition of until_nth_event applies and crossings[n] is appended."""
    L = speed_lib() if speed_build else lib()
    state_soa = np.ascontiguousarray(state_soa, dtype=np.float64)
    consts_soa = np.ascontiguousarray(consts_soa, dtype=np.float64)
    epoch0_ns = np.ascontiguousarray(epoch0_ns, dtype=np.int64)
    n = state_soa.shape[1]
    assert state_soa.shape == (9, n) and consts_soa.shape == (4, n) and epoch0_ns.shape == (n,)
    out_state = np.empty((9, n), dtype=np.float64)
    out_epoch = np.empty(n, dtype=np.int64)
    details = np.zeros(n, dtype=abi.DETAILS_DTYPE)
    status = np.zeros(n, dtype=np.int32)
    step_ptr = None
    if step_ns is not None:
        assert step_ns.dtype == np.int64 and step_ns.shape == (n,)
        step_ptr = step_ns.ctypes.data
    sink = None
    if traj_capacity:
        t_ep = np.zeros((traj_capacity, n), dtype=np.int64)
        t_st = np.zeros((6, traj_capacity, n), dtype=np.float64)
        t_cnt = np.zeros(n, dtype=np.int64)
        sink = abi.TrajSink(int(traj_capacity), t_ep.ctypes.data, t_st.ctypes.data, t_cnt.ctypes.data)
    ev = None
    if event is not None:
        crossings = np.zeros(n, dtype=np.int32)
        ev = abi.EventC(int(event[0]), int(event[2]), float(event[1]), crossings.ctypes.data)
    rc = L.nyx_oracle_propagate_batch_event(
        C.byref(dyn_c), C.byref(opts_c), n, state_soa.ctypes.data, consts_soa.ctypes.data, epoch0_ns.ctypes.data,
        int(end_epoch_ns), step_ptr, out_state.ctypes.data, out_epoch.ctypes.data, details.ctypes.data,
        status.ctypes.data, C.byref(sink) if sink is not None else None, C.byref(ev) if ev is not None else None, int(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle rejected the configuration (rc={rc})")
    ret = (out_state, out_epoch, details, status)
    if traj_capacity:
        ret = ret + ((t_ep, t_st, t_cnt),)
    if ev is not None:
        ret = ret + (crossings,)
    return ret


def num_threads() -> int:
    return lib().nyx_oracle_num_threads()


# --------------------------------------------------------------------------- STM path
def dual_eom(dyn_c, t_ns, y9, consts4):
    """One evaluation of SpacecraftDynamics::dual_eom (spacecraft.rs:312-363): returns (dx[9], A[9][9])."""
    L = lib()
    y9 = np.ascontiguousarray(y9, dtype=np.float64)
    consts4 = np.ascontiguousarray(consts4, dtype=np.float64)
    dx = np.zeros(9)
    grad = np.zeros(81)
    rc = L.nyx_oracle_dual_eom(C.byref(dyn_c), int(t_ns), abi.as_double_p(y9), abi.as_double_p(consts4), abi.as_double_p(dx), abi.as_double_p(grad))
    if rc != 0:
        raise RuntimeError(f"oracle dual_eom rc={rc}")
    return dx, grad.reshape(9, 9)


def propagate_batch_stm(dyn_c, opts_c, state_soa, consts_soa, epoch0_ns, end_epoch_ns, stm_in=None, step_ns=None, n_threads=0):
    """Same contract as nyxb_propagate_batch_stm: returns (state[9][n], epoch[n], stm[81][n], details, status)."""
    L = lib()
    state_soa = np.ascontiguousarray(state_soa, dtype=np.float64)
    consts_soa = np.ascontiguousarray(consts_soa, dtype=np.float64)
    epoch0_ns = np.ascontiguousarray(epoch0_ns, dtype=np.int64)
    n = state_soa.shape[1]
    out_state = np.empty((9, n)); out_epoch = np.empty(n, dtype=np.int64); out_stm = np.empty((81, n))
    details = np.zeros(n, dtype=abi.DETAILS_DTYPE); status = np.zeros(n, dtype=np.int32)
    if stm_in is not None:
        stm_in = np.ascontiguousarray(stm_in, dtype=np.float64)
        assert stm_in.shape == (81, n)
    rc = L.nyx_oracle_propagate_batch_stm(
        C.byref(dyn_c), C.byref(opts_c), n, state_soa.ctypes.data, consts_soa.ctypes.data, epoch0_ns.ctypes.data, int(end_epoch_ns),
        step_ns.ctypes.data if step_ns is not None else None, stm_in.ctypes.data if stm_in is not None else None,
        out_state.ctypes.data, out_epoch.ctypes.data, out_stm.ctypes.data, details.ctypes.data, status.ctypes.data, int(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle rejected the STM configuration (rc={rc})")
    return out_state, out_epoch, out_stm, details, status


class Inst:
    """PropInstance over the 90-vector (state + STM) for the numpy Kalman-filter restatement (pyoracle_od.py)."""

    def __init__(self, dyn_c, opts_c, y9, consts4, epoch_ns):
        self._L = lib()
        self._keep = (dyn_c, opts_c)
        y9 = np.ascontiguousarray(y9, dtype=np.float64)
        consts4 = np.ascontiguousarray(consts4, dtype=np.float64)
        self._h = self._L.nyx_oracle_inst_new(C.byref(dyn_c), C.byref(opts_c), abi.as_double_p(y9), abi.as_double_p(consts4), int(epoch_ns))
        if not self._h:
            raise RuntimeError("oracle rejected the STM configuration")

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.nyx_oracle_inst_free(self._h)
            self._h = None

    def for_duration(self, duration_ns) -> int:
        return self._L.nyx_oracle_inst_for_duration(self._h, int(duration_ns))

    def get(self):
        y = np.zeros(90)
        ep, st = C.c_int64(), C.c_int64()
        fx = C.c_int()
        det = np.zeros(1, dtype=abi.DETAILS_DTYPE)
        self._L.nyx_oracle_inst_get(self._h, abi.as_double_p(y), C.byref(ep), C.byref(st), C.byref(fx), det.ctypes.data)
        return y, ep.value, st.value, fx.value, det[0]

    def set(self, y90, epoch_ns):
        y90 = np.ascontiguousarray(y90, dtype=np.float64)
        self._L.nyx_oracle_inst_set(self._h, abi.as_double_p(y90), int(epoch_ns))

    def set_step(self, step_ns, fixed):
        self._L.nyx_oracle_inst_set_step(self._h, int(step_ns), int(bool(fixed)))


def mvn_sample(seed, first_index, n, template9, mean9, sqrt_s_v):
    """Same contract as nyxb_mvn_sample: returns (state[9][n], dispersion[9][n])."""
    L = lib()
    t = np.ascontiguousarray(template9, dtype=np.float64); m = np.ascontiguousarray(mean9, dtype=np.float64)
    sq = np.ascontiguousarray(np.asarray(sqrt_s_v, dtype=np.float64).reshape(81))
    out = np.empty((9, n)); disp = np.empty((9, n))
    rc = L.nyx_oracle_mvn_sample(int(seed), int(first_index), n, abi.as_double_p(t), abi.as_double_p(m), abi.as_double_p(sq), out.ctypes.data, disp.ctypes.data)
    assert rc == 0
    return out, disp
