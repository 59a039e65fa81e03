"""Shared helpers for the parity tests."""
import json
from pathlib import Path

import numpy as np

import nyx_b200 as nb
from nyx_b200 import abi

GOLDEN = json.loads((Path(__file__).parent / "golden" / "reference_vectors.json").read_text())
S = 10**9


def opts_from_json(o) -> nb.IntegratorOptions:
    if o["kind"] == "fixed":
        return nb.IntegratorOptions.with_fixed_step_s(o["step_s"])
    if o["kind"] == "default":
        return nb.IntegratorOptions.default()
    return nb.IntegratorOptions.with_adaptive_step_s(o["min_step_s"], o["max_step_s"], o["tolerance"],
                                                     nb.ErrorControl[o["error_ctrl"]])


def leo_state(frame, epoch_ns=0) -> nb.Spacecraft:
    x = GOLDEN["initial_state"]
    return nb.Spacecraft.from_orbit(nb.Orbit.cartesian(*x, epoch_ns, frame))


def leo_ensemble(n, seed=0, sma=6678.0, pos_std=1.0, vel_std=1e-3, frame=None, mass=None, srp=None, drag=None):
    """Example-01 orbit (examples/01_orbit_prop/main.rs:52-53) + N(0, diag(1 km, 1 m/s)) dispersions."""
    frame = frame or nb.EARTH_J2000
    orbit = nb.Orbit.keplerian(sma, 0.015, 68.5, 65.2, 75.0, 0.0, 0, frame)
    template = nb.Spacecraft(orbit=orbit, mass=mass or nb.Mass(100.0, 20.0, 0.0), srp=srp or nb.SRPData(),
                             drag=drag or nb.DragData())
    mvn = nb.MvnSpacecraft.from_cartesian_std(template, pos_std, vel_std)
    mc = nb.MonteCarlo(template, mvn, "test", seed=seed)
    return mc, nb.pack_spacecraft(ds.state for _, ds in mc.generate_states(0, n))


def oracle_run(oracle, prop: nb.Propagator, frame, almanac, st, cs, ep, end_ns, step_ns=None):
    packed = prop.dynamics.pack(frame, almanac)
    return oracle.propagate_batch(packed.c, prop.opts.to_c(prop.method), st, cs, ep, end_ns, step_ns)


def max_dr_dv(a, b):
    d = a - b
    return float(np.sqrt((d[:3] ** 2).sum(0)).max()), float(np.sqrt((d[3:6] ** 2).sum(0)).max())


# ---- batched resampling (nyxb_traj_resample)
def resample_reference(sc, t_ep, t_st, t_cnt, queries):
    """Traj.at per trajectory and query -> (states[6][m][n], status[m][n]); NaN + 1 where TrajError."""
    from nyx_b200.trajectory import Traj, TrajError
    n, m = t_ep.shape[1], len(queries)
    out = np.full((6, m, n), np.nan)
    status = np.ones((m, n), dtype=np.int32)
    for i in range(n):
        k = int(t_cnt[i])
        tr = Traj(sc, t_ep[:k, i].copy(), np.ascontiguousarray(t_st[:, :k, i].T)).finalize()
        for j, q in enumerate(queries):
            try:
                out[:, j, i] = tr.at(int(q)).orbit.to_cartesian_pos_vel()
                status[j, i] = 0
            except TrajError:
                pass
    return out, status


def hermite_shim(tmp_path):
    """The resampling kernel's per-(query, trajectory) function (nyx_b200/csrc/nyxb_hermite.h) compiled for the host:
    run(t_ep, t_st, t_cnt, queries) -> (states[6][m][n], status[m][n])."""
    import ctypes
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    so = tmp_path / "hermite_core_shim.so"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC",
                    str(root / "tests" / "cpp" / "hermite_core_shim.cpp"), "-o", str(so)], check=True, capture_output=True)
    lib = ctypes.CDLL(str(so))
    lib.shim_event_locate.restype = None
    lib.shim_event_locate.argtypes = ([ctypes.c_longlong] + [ctypes.c_void_p] * 3 + [ctypes.c_size_t, ctypes.c_int, ctypes.c_double,
                                      ctypes.c_longlong] + [ctypes.c_void_p] * 4)
    lib.shim_traj_resample.restype = None
    lib.shim_traj_resample.argtypes = [ctypes.c_longlong] + [ctypes.c_void_p] * 3 + [ctypes.c_size_t] * 2 + [ctypes.c_void_p] * 3
    def run(t_ep, t_st, t_cnt, queries):
        cap, n = t_ep.shape
        q = np.ascontiguousarray(queries, dtype=np.int64)
        out = np.empty((6, len(q), n))
        status = np.empty((len(q), n), dtype=np.int32)
        t_ep, t_st, t_cnt = (np.ascontiguousarray(a) for a in (t_ep, t_st, t_cnt))
        lib.shim_traj_resample(cap, t_ep.ctypes.data, t_st.ctypes.data, t_cnt.ctypes.data, n, len(q), q.ctypes.data,
                               out.ctypes.data, status.ctypes.data)
        return out, status

    def locate(t_ep, t_st, t_cnt, kind, value, precision_ns, run_status=None):
        cap, n = t_ep.shape
        t_ep, t_st, t_cnt = (np.ascontiguousarray(a) for a in (t_ep, t_st, t_cnt))
        rs = None if run_status is None else np.ascontiguousarray(run_status, dtype=np.int32)
        ev_ep = np.zeros(n, dtype=np.int64)
        ev_st = np.empty((6, n))
        status = np.empty(n, dtype=np.int32)
        lib.shim_event_locate(cap, t_ep.ctypes.data, t_st.ctypes.data, t_cnt.ctypes.data, n, int(kind), float(value), int(precision_ns),
                              None if rs is None else rs.ctypes.data, ev_ep.ctypes.data, ev_st.ctypes.data, status.ctypes.data)
        return ev_ep, ev_st, status

    run.locate = locate
    return run


def resample_queries(t_ep, t_cnt, end):
    """exact hits, both window edges, out-of-span epochs, a regular grid"""
    k0 = int(t_cnt[0])
    return np.concatenate([
        np.array([-5, 0, 1, int(t_ep[1, 0]) - 1, int(t_ep[3, 0]), int(t_ep[k0 - 2, 0]) + 1, end - 1, end, end + 1], dtype=np.int64),
        np.arange(0, end, 600 * S, dtype=np.int64) + 123_456_789,
    ])


# ---- CPU stand-in for `nyx_b200.Engine`: the same host-facing methods on the oracle, so that the host mirror above the C ABI
# (PropInstance, Propagator.many_*, MonteCarlo, Results) can be exercised without a device.  Test infrastructure only.
class OracleEngine:
    def __init__(self, oracle, prop, frame, almanac, tmp_path=None):
        self.oracle = oracle
        self.packed, self.opts = prop.lower(frame, almanac)
        self.launches = 0
        self._resident = None
        self._shim = hermite_shim(tmp_path) if tmp_path is not None else None

    def propagate_batch(self, st, cs, ep, end_ns, step_ns=None, traj_capacity=0, event=None):
        self.launches += 1
        ret = self.oracle.propagate_batch(self.packed.c, self.opts, st, cs, ep, int(end_ns), step_ns, traj_capacity=traj_capacity, event=event)
        if traj_capacity:
            self._resident = ret[4]
        return ret

    def resample(self, q, recording=None, n=None):
        if recording is not None:
            self._resident = recording
        return self._shim(*self._resident, q)

    def locate_events(self, kind, value, precision_ns, recording=None, n=None, run_status=None):
        if recording is not None:
            self._resident = recording
        return self._shim.locate(*self._resident, kind, value, precision_ns, run_status)

    def od_ekf_batch(self, cfg_c, n_stations, stations_c, msr_epoch_ns, msr_tracker, obs, state_soa, consts_soa, epoch0_ns, covar0_soa,
                     record_estimates=False):
        """`Engine.od_ekf_batch` on the numpy / C oracle filter, one filter after the other."""
        from nyx_b200.od import ODSolution
        from oracle import pyoracle_od

        self.launches += 1
        n, m = state_soa.shape[1], len(msr_epoch_ns)
        out_state = np.empty((9, n)); out_epoch = np.empty(n, dtype=np.int64); covar = np.empty((n, 9, 9)); dev = np.empty((9, n))
        ratio = np.full((m, 2, n), np.nan); prefit = np.full((m, 2, n), np.nan); postfit = np.full((m, 2, n), np.nan)
        flags = np.zeros((m, n), dtype=np.int32)
        est_state = np.full((m, 9, n), np.nan) if record_estimates else None
        est_cov = np.full((m, 9, n), np.nan) if record_estimates else None
        details = np.zeros(n, dtype=abi.DETAILS_DTYPE); status = np.zeros(n, dtype=np.int32)
        for i in range(n):
            cov0 = covar0_soa[:, i].reshape(9, 9).T          # (c*9 + r) -> [r][c]
            r = pyoracle_od.process_arc(self.packed.c, self.opts, cfg_c, stations_c, msr_epoch_ns, np.asarray(msr_tracker, dtype=np.int32),
                                        np.ascontiguousarray(obs[:, :, i]), state_soa[:, i].copy(), consts_soa[:, i].copy(), int(epoch0_ns[i]), cov0)
            out_state[:, i], out_epoch[i], covar[i], dev[:, i] = r["state"], r["epoch"], r["covar"], r["state_dev"]
            ratio[:, :, i], prefit[:, :, i], postfit[:, :, i], flags[:, i] = r["resid_ratio"], r["prefit"], r["postfit"], r["msr_flags"]
            if record_estimates:
                est_state[:, :, i], est_cov[:, :, i] = r["est_state"], r["est_covar_diag"]
            details["n_steps"][i], status[i] = r["n_steps"], r["status"]
        return ODSolution(out_state, out_epoch, covar, dev, ratio, prefit, postfit, flags, est_state, est_cov, details, status)

    def launch_count(self):
        return self.launches


def use_oracle_engine(monkeypatch, oracle, prop, tmp_path=None):
    """Route `prop.engine(frame, almanac)` to an OracleEngine (one per (frame, almanac), like the real cache)."""
    cache = {}

    def engine(frame, almanac):
        key = (id(almanac), frame)
        if key not in cache:
            cache[key] = OracleEngine(oracle, prop, frame, almanac, tmp_path)
        return cache[key]

    monkeypatch.setattr(prop, "engine", engine)
    return cache
