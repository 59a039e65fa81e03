"""TEST INFRASTRUCTURE — numpy restatement of the reference's sequential Kalman orbit determination for ONE filter
(never imported by nyx_b200).  It drives the C oracle's 90-vector `PropInstance` (oracle/nyx_oracle_od.c) and follows,
line by line (paths relative to /root/reference/nyx-core/src):

  KalmanODProcess::process_arc           od/process/mod.rs:128-497   (the loop is :211-426)
  KalmanFilter::time_update              od/kalman/filtering.rs:59-102
  KalmanFilter::measurement_update       od/kalman/filtering.rs:107-316
  ProcessNoise::{to_matrix, propagate}   od/snc.rs:175-286
  GroundStation::measure_instantaneous   od/ground_station/trk_device.rs:154-200
  ScalarSensitivity::new (h_tilde)       od/msr/sensitivity.rs:118-239
  MeasurementType::compute_one_way       od/msr/types.rs:104-121

PARITY UNPINNED: the reference's OD tests (tests/orbit_determination/*.rs) assert estimation errors against truth
trajectories built from anise data that is absent here (DE440s, pck08), not intermediate filter values.  The geometry
that anise supplies (`azimuth_elevation_range_sez_from_location`, `line_of_sight_obstructed`) is restated from its
published algorithms: range = |rho|, range rate = rho.rho_dot/|rho|, elevation from the local zenith, Vallado's SIGHT
test for the obstruction.  The matrix algebra uses numpy (`@`), the reference uses nalgebra: tolerance-level parity.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from nyx_b200 import abi

from . import pyoracle

MSRF_PROCESSED, MSRF_REJECTED, MSRF_NOT_VISIBLE, MSRF_ABSENT = 1, 2, 4, 8


def _rotation(rot_c, t_ns):
    R = np.zeros(9)
    wdot = C.c_double()
    pyoracle.lib().nyx_oracle_rotation(C.byref(rot_c), int(t_ns), abi.as_double_p(R), C.byref(wdot))
    return R.reshape(3, 3), wdot.value


def _body_pos_vel(body_c, t_ns):
    """Chebyshev position (C oracle) and velocity (derivative of the same series) of a body w.r.t. the centre."""
    pos = np.zeros(3)
    if pyoracle.lib().nyx_oracle_body_position(C.byref(body_c), int(t_ns), abi.as_double_p(pos)) != 0:
        raise RuntimeError("epoch outside ephemeris coverage")
    nc = body_c.n_coeffs
    coeffs = np.ctypeslib.as_array(body_c.coeffs, shape=(body_c.n_intervals, 3, nc))
    dt = int(t_ns) - body_c.t0_ns
    idx, off = divmod(dt, body_c.interval_ns)
    tau = 2.0 * (off / body_c.interval_ns) - 1.0
    scale = 2.0 / (body_c.interval_ns * 1e-9)
    vel = np.array([np.polynomial.chebyshev.chebval(tau, np.polynomial.chebyshev.chebder(coeffs[idx, ax])) for ax in range(3)]) * scale
    return pos, vel


def station_state(gs_c, dyn_c, t_ns):
    """trk_device.rs:150-152 `location`: antenna position/velocity in the integration frame + inertial zenith."""
    R, wdot = _rotation(gs_c.rot, t_ns)
    p = np.array(list(gs_c.pos_fixed_km))
    up = np.array(list(gs_c.up_fixed))
    r = R.T @ p
    v = R.T @ np.cross(np.array([0.0, 0.0, wdot]), p)
    if gs_c.body != abi.NYXB_CENTRAL_BODY:
        bp, bv = _body_pos_vel(dyn_c.bodies[gs_c.body], t_ns)
        r, v = r + bp, v + bv
    return r, v, R.T @ up


def measure(gs_c, dyn_c, t_ns, y):
    """measure_instantaneous without noise (rng = None): returns None when not visible, else dict type -> value."""
    r_tx, v_tx, up = station_state(gs_c, dyn_c, t_ns)
    rho = y[:3] - r_tx
    rng = math.sqrt(rho @ rho)
    rr = float(rho @ (y[3:6] - v_tx)) / rng
    elev = math.degrees(math.asin(float(rho @ up) / rng))
    if elev - gs_c.elevation_mask_deg < 0.0:
        return None, (rng, rr)
    if gs_c.body != abi.NYXB_CENTRAL_BODY and gs_c.body_radius_km > 0.0:
        r1, r2 = y[:3], r_tx
        r1sq, r2sq, r12 = float(r1 @ r1), float(r2 @ r2), float(r1 @ r2)
        tau = (r1sq - r12) / (r1sq + r2sq - 2.0 * r12)
        if 0.0 <= tau <= 1.0 and (1.0 - tau) * r1sq + r12 * tau <= gs_c.body_radius_km ** 2:
            return None, (rng, rr)
    return {abi.MSR_RANGE: rng, abi.MSR_DOPPLER: rr}, (rng, rr)


def _snc(cfg, y, epoch_ns, prev_epoch_ns, delta_t_ns):
    """ProcessNoise::propagate (snc.rs:211-286) for one 3x3 SNC; returns a 9x9 matrix or None."""
    if not cfg.snc_enabled:
        return None
    if epoch_ns - prev_epoch_ns > cfg.snc_disable_time_ns:      # to_matrix :188-196 (prev_epoch is set by the filter)
        return None
    snc = np.diag(np.array(list(cfg.snc_diag)))
    if cfg.snc_frame == 1:                                      # :219-262 RIC: keep the rotated diagonal only
        r, v = y[:3], y[3:6]
        rh = r / math.sqrt(r @ r)
        h = np.cross(r, v)
        ch = h / math.sqrt(h @ h)
        ih = np.cross(ch, rh)
        D = np.column_stack([rh, ih, ch])
        snc = np.diag(np.diag(D @ snc @ D.T))
    if delta_t_ns > cfg.snc_disable_time_ns:                    # :264-266
        return None
    dt = pyoracle.lib().nyx_oracle_dur_to_seconds(int(delta_t_ns))
    gamma = np.zeros((9, 3))
    for i in range(3):
        gamma[i, i] = dt ** 2 / 2.0
        gamma[i + 3, i] = dt
    return gamma @ snc @ gamma.T


def process_arc(dyn_c, opts_c, cfg, stations_c, msr_epoch_ns, msr_tracker, obs, y9, consts4, epoch0_ns, covar0):
    """One `KalmanODProcess::process_arc`.  obs: [m][2] (NaN = type absent).  Returns a dict of the quantities the
    C ABI's nyxb_od_outputs carries."""
    m = len(msr_epoch_ns)
    inst = pyoracle.Inst(dyn_c, opts_c, y9, consts4, epoch0_ns)            # prop.with(nominal.with_stm()) :167
    y, ep, step, fixed, _ = inst.get()
    if not fixed:
        inst.set_step(cfg.max_step_ns, False)                               # :170-172
    # filter state (KalmanFilter, kalman/mod.rs)
    P = np.array(covar0, dtype=np.float64).reshape(9, 9).copy()
    xdev = np.zeros(9)
    prev_epoch = int(epoch0_ns)
    epoch = int(epoch0_ns)
    ratio = np.full((m, 2), np.nan); prefit_o = np.full((m, 2), np.nan); postfit_o = np.full((m, 2), np.nan)
    flags = np.zeros(m, dtype=np.int32)
    est_state = np.full((m, 9), np.nan); est_cov = np.full((m, 9), np.nan)
    status = 0
    ekf = cfg.variant == abi.KF_REFERENCE_UPDATE
    reject = cfg.reject_num_sigmas if cfg.reject_num_sigmas >= 0.0 else None

    def reset_stm():
        y, ep, *_ = inst.get()
        y[9:] = np.eye(9).reshape(81)
        inst.set(y, ep)

    def time_update(y, ep):
        nonlocal P, xdev, prev_epoch
        stm = y[9:].reshape(9, 9).T                                          # column-major tail
        P_bar = stm @ P @ stm.T                                              # filtering.rs:61
        q = _snc(cfg, y, ep, prev_epoch, ep - prev_epoch)
        if q is not None:
            P_bar = P_bar + q
        xdev = stm @ xdev if not ekf else np.zeros(9)                        # :81-85
        P = P_bar
        prev_epoch = ep
        return P_bar

    for k in range(m):
        t_k = int(msr_epoch_ns[k])
        o = obs[k]
        if np.isnan(o[0]) and np.isnan(o[1]):
            flags[k] = MSRF_ABSENT
            continue
        while True:
            delta_t = t_k - epoch
            y, ep, step, fixed, _ = inst.get()
            next_step = min(delta_t, step, cfg.max_step_ns)                  # :218
            rc = inst.for_duration(next_step)                                # :232-234
            if rc:
                status = rc
                break
            y, ep, step, fixed, _ = inst.get()
            epoch = ep
            if abs(ep - t_k) < cfg.epoch_precision_ns:                       # :250
                inst.set(y, t_k)                                             # :254 set_epoch
                ep = epoch_for_msr = t_k
                trk = int(msr_tracker[k])
                if trk < 0:                                                  # unknown tracker :400-410
                    break
                gs = stations_c[trk]
                n_types = gs.n_types
                windows = n_types // cfg.msr_size
                for wno in range(windows + 1):                               # :270-398
                    y, ep_now, *_ = inst.get()                               # nominal_state = prop_instance.state
                    cur = [gs.types[q] for q in range(wno * cfg.msr_size, min((wno + 1) * cfg.msr_size, n_types))]
                    if not cur:
                        break
                    avail = [not np.isnan(o[t]) for t in cur]
                    if not any(avail):
                        continue
                    M = cfg.msr_size
                    real_obs = np.zeros(M)
                    for i, t in enumerate(cur):
                        if avail[i]:
                            real_obs[i] = o[t]
                    # h_tilde (sensitivity.rs:88-115): identity rows unless the type is in msr.data
                    H = np.eye(M, 9)
                    r_tx, v_tx, _up = station_state(gs, dyn_c, epoch_for_msr)
                    dr = y[:3] - r_tx
                    dv = y[3:6] - v_tx
                    computed, (rng_now, _rr) = measure(gs, dyn_c, epoch_for_msr, y)
                    for i, t in enumerate(cur):
                        if not avail[i]:
                            continue
                        if t == abi.MSR_DOPPLER:                             # :145-175 (range recomputed, rho_dot = the observation)
                            rho = rng_now
                            rho_dot = o[abi.MSR_DOPPLER]
                            H[i] = [dv[0] / rho - rho_dot * dr[0] / rho ** 2, dv[1] / rho - rho_dot * dr[1] / rho ** 2,
                                    dv[2] / rho - rho_dot * dr[2] / rho ** 2, dr[0] / rho, dr[1] / rho, dr[2] / rho, 0, 0, 0]
                        else:                                                # :176-192 (rho = the observation)
                            rho = o[abi.MSR_RANGE]
                            H[i] = [dr[0] / rho, dr[1] / rho, dr[2] / rho, 0, 0, 0, 0, 0, 0]
                    Rk = np.zeros((M, M))
                    bias = np.zeros(M)
                    for i, t in enumerate(cur):                              # trackdata.rs:84-122
                        q = [gs.types[j] for j in range(n_types)].index(t)
                        Rk[i, i] = gs.noise_var[q]
                        bias[i] = gs.bias[q]
                    if computed is None:                                     # :386-392
                        flags[k] |= MSRF_NOT_VISIBLE
                        continue
                    comp = np.zeros(M)
                    for i, t in enumerate(cur):
                        comp[i] = computed[t]
                    comp = comp - bias
                    # ---- measurement_update (filtering.rs:107-316)
                    stm = y[9:].reshape(9, 9).T
                    P_bar = stm @ P @ stm.T
                    q = _snc(cfg, y, ep_now, prev_epoch, ep_now - prev_epoch)
                    if q is not None:
                        P_bar = P_bar + q
                    PHt = P_bar @ H.T
                    S = H @ PHt + Rk
                    pre = real_obs - comp
                    try:
                        L = np.linalg.cholesky(S)
                    except np.linalg.LinAlgError:
                        L = np.linalg.cholesky(Rk)
                    white = np.linalg.solve(L, pre)
                    rat = math.sqrt(float(white @ white) / M)
                    slot = wno if cfg.msr_size == 1 else 0
                    ratio[k, slot] = rat
                    for i, t in enumerate(cur):
                        prefit_o[k, wno * cfg.msr_size + i] = pre[i]
                    flags[k] |= MSRF_PROCESSED
                    if reject is not None and rat > reject:                  # :169-184
                        time_update(y, ep_now)
                        flags[k] |= MSRF_REJECTED
                        rejected = True
                    else:
                        K = np.linalg.solve(S, PHt.T).T                      # :206-231
                        if ekf:
                            x_hat = K @ pre
                            post = pre - H @ x_hat
                        else:
                            x_bar = stm @ xdev
                            post = pre - H @ x_bar
                            x_hat = x_bar + K @ post
                        first = np.eye(9) - K @ H
                        cov = first @ P_bar @ first.T + K @ Rk @ K.T          # Joseph :292-297
                        P = 0.5 * (cov + cov.T)
                        xdev = x_hat
                        prev_epoch = ep_now
                        for i, t in enumerate(cur):
                            postfit_o[k, wno * cfg.msr_size + i] = post[i]
                        rejected = False
                        if ekf:                                              # :364-369 replace the state
                            ynew = y.copy()
                            ynew[:9] = y[:9] + x_hat
                            ynew[6] = min(max(ynew[6], 0.0), 2.0)            # Spacecraft + OVector clamps Cr
                            inst.set(ynew, ep_now)
                    reset_stm()                                              # :371
                y, ep_now, *_ = inst.get()
                est_state[k] = y[:9]
                est_cov[k] = np.diag(P)
                break
            else:
                time_update(y, ep)                                           # :417-421
                reset_stm()
        if status:
            break
    y, ep, step, fixed, det = inst.get()
    return dict(state=y[:9].copy(), epoch=ep, covar=P, state_dev=xdev, resid_ratio=ratio, prefit=prefit_o, postfit=postfit_o,
                msr_flags=flags, est_state=est_state, est_covar_diag=est_cov, n_steps=int(det["n_steps"]), status=status)
