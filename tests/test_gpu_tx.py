"""GPU parity of the TRANSPOSED kernel (csrc/nyxb_tx.cu: lane = trajectory, walker warp = column position, helper warps for the serial
phases, two sets in flight per persistent CTA, (set, time-slice) tickets) against the CPU oracle, through the C ABI with `nyxb_engine_set_kernel(NYXB_KERNEL_TRANSPOSED)`.
Tolerances as for the lane-cooperative FAST kernel: adaptive runs differ from the oracle by the integrator's own step-sequence
sensitivity (5e-7 km over 4-6 h, 1e-6 km being the north-star bound); a FIXED step pins the regrouped harmonic sum to round-off
(5e-9 km).  Time slicing must not change a single bit (same arithmetic, state parked and restored exactly)."""
import numpy as np
import pytest

import nyx_b200 as nb
from tests.util import S, leo_ensemble, max_dr_dv, oracle_run

pytestmark = pytest.mark.gpu
DAY = 86400 * S


def _leo_dyn(degree=21, order=None, extras=False, days=1.0):
    gd = nb.GravityFieldData.from_fixture("jgm3_70x70", degree, degree if order is None else order, nb.IAU_EARTH_FRAME)
    if not extras:
        return nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(nb.GravityField.new(gd))), None
    almanac = nb.Almanac.synthetic(nb.EARTH_J2000, 0, days)
    orb = nb.OrbitalDynamics.new([nb.PointMasses.new([nb.MOON, nb.SUN]), nb.GravityField.new(gd)])
    return nb.SpacecraftDynamics.new(orb), almanac


def _tx_engine(prop, almanac=None, frame=None):
    eng = prop.engine(frame or nb.EARTH_J2000, almanac)
    eng.set_kernel(nb.KERNEL_TRANSPOSED)
    return eng


@pytest.mark.parametrize("n", [1, 33, 70, 160])
def test_transposed_kernel_ragged_sizes_and_rejections(oracle, n):
    """Sets with absent lanes, trajectories of a set ending at different step counts, forced rejections, fixed steps."""
    mc, (st, cs, ep) = leo_ensemble(n, seed=13)
    ep = ep + (np.arange(n, dtype=np.int64) % 5) * 600 * S   # different start epochs => different step counts
    dyn, _ = _leo_dyn()
    end = 4 * 3600 * S
    for opts, tol in ((nb.IntegratorOptions.default(), 5e-7), (nb.IntegratorOptions.with_fixed_step_s(45.0), 5e-9),
                      (nb.IntegratorOptions(init_step=600 * nb.Unit.Second, tolerance=1e-13), 5e-7)):
        prop = nb.Propagator.rk89(dyn, opts, mode=nb.MODE_FAST)
        eng = _tx_engine(prop)
        out, out_ep, det, status = eng.propagate_batch(st, cs, ep, end)
        assert eng.last_kernel() == nb.KERNEL_TRANSPOSED
        ref, ref_ep, ref_det, ref_status = oracle_run(oracle, prop, nb.EARTH_J2000, None, st, cs, ep, end)
        assert np.array_equal(status, ref_status) and np.array_equal(out_ep, ref_ep)
        assert max_dr_dv(out, ref)[0] < tol, (opts, max_dr_dv(out, ref))
        assert np.abs(det["n_steps"] - ref_det["n_steps"]).max() <= 1
        assert np.abs(det["n_rejected"] - ref_det["n_rejected"]).max() <= 1
        assert np.array_equal(det["n_rhs"], 16 * (det["n_steps"] + det["n_rejected"]))


@pytest.mark.parametrize("degree,order", [(21, 21), (12, 7), (8, 8), (40, 40), (70, 70)])
def test_transposed_kernel_vs_oracle_and_cooperative(oracle, degree, order):
    n = 96 if degree <= 40 else 40
    mc, (st, cs, ep) = leo_ensemble(n, seed=11)
    dyn, _ = _leo_dyn(degree, order)
    prop = nb.Propagator.default(dyn, mode=nb.MODE_FAST)
    eng = _tx_engine(prop)
    out, out_ep, det, status = eng.propagate_batch(st, cs, ep, 6 * 3600 * S)
    ref, ref_ep, ref_det, ref_status = oracle_run(oracle, prop, nb.EARTH_J2000, None, st, cs, ep, 6 * 3600 * S)
    assert (status == 0).all() and np.array_equal(out_ep, ref_ep)
    dr, dv = max_dr_dv(out, ref)
    assert dr < 5e-7 and dv < 1e-9, (dr, dv)
    assert np.abs(det["n_steps"] - ref_det["n_steps"]).max() <= 1
    eng.set_kernel(nb.KERNEL_COOP)   # same engine, lane-cooperative kernel: same algorithm class
    out2 = eng.propagate_batch(st, cs, ep, 6 * 3600 * S)[0]
    assert eng.last_kernel() == nb.KERNEL_COOP
    assert max_dr_dv(out, out2)[0] < 5e-7


def test_transposed_kernel_fixed_step_tight_with_third_bodies(oracle):
    """Fixed step: no step-sequence divergence, so agreement to round-off pins the harmonic sum, the DCM update and the
    point-mass path of this kernel (6 h, 5e-9 km)."""
    mc, (st, cs, ep) = leo_ensemble(64, seed=12)
    dyn, almanac = _leo_dyn(21, extras=True)
    for method in (nb.IntegratorMethod.RungeKutta89, nb.IntegratorMethod.DormandPrince78, nb.IntegratorMethod.RungeKutta4):
        prop = nb.Propagator.new(dyn, method, nb.IntegratorOptions.with_fixed_step_s(60.0 if method != nb.IntegratorMethod.RungeKutta4 else 10.0),
                                 mode=nb.MODE_FAST)
        eng = _tx_engine(prop, almanac)
        out, out_ep, det, status = eng.propagate_batch(st, cs, ep, 6 * 3600 * S)
        ref, ref_ep, ref_det, _ = oracle_run(oracle, prop, nb.EARTH_J2000, almanac, st, cs, ep, 6 * 3600 * S)
        assert (status == 0).all() and np.array_equal(det["n_steps"], ref_det["n_steps"])
        dr, dv = max_dr_dv(out, ref)
        assert dr < (5e-9 if method != nb.IntegratorMethod.RungeKutta4 else 5e-8) and dv < 5e-11, (method, dr, dv)


@pytest.mark.parametrize("positions,degree", [(10, 21), (10, 16), (8, 21), (16, 21)])
def test_transposed_kernel_walker_positions(oracle, positions, degree):
    """The same field walked by 8, 10 or 16 warps per set (different column schedules, published or assembled start powers):
    fixed steps, 6 h, agreement with the oracle to round-off."""
    mc, (st, cs, ep) = leo_ensemble(70, seed=14)
    dyn, _ = _leo_dyn(degree)
    prop = nb.Propagator.new(dyn, nb.IntegratorMethod.RungeKutta89, nb.IntegratorOptions.with_fixed_step_s(60.0), mode=nb.MODE_FAST)
    eng = _tx_engine(prop)
    eng.set_tx_positions(positions)
    out, out_ep, det, status = eng.propagate_batch(st, cs, ep, 6 * 3600 * S)
    assert eng.last_kernel() == nb.KERNEL_TRANSPOSED
    ref, ref_ep, ref_det, _ = oracle_run(oracle, prop, nb.EARTH_J2000, None, st, cs, ep, 6 * 3600 * S)
    assert (status == 0).all() and np.array_equal(det["n_steps"], ref_det["n_steps"])
    dr, dv = max_dr_dv(out, ref)
    assert dr < 5e-9 and dv < 5e-11, (positions, dr, dv)


def test_transposed_kernel_time_slicing_is_bit_invisible(oracle):
    """More sets than persistent CTAs: sets are parked after every slice and resumed by whichever CTA draws their next ticket.
    Results, details and the recorded trajectories must equal the all-resident run bit for bit."""
    n = 200   # 7 sets
    mc, (st, cs, ep) = leo_ensemble(n, seed=21)
    ep = ep + (np.arange(n, dtype=np.int64) % 3) * 900 * S
    dyn, _ = _leo_dyn()
    prop = nb.Propagator.rk89(dyn, nb.IntegratorOptions(init_step=600 * nb.Unit.Second, tolerance=1e-13), mode=nb.MODE_FAST)   # forced rejections
    eng = _tx_engine(prop)
    end = 5 * 3600 * S
    step0 = np.full(n, 600 * S, dtype=np.int64)   # the PropInstance step carried in and out (instance.rs:56)
    base = eng.propagate_batch(st, cs, ep, end, step_ns=step0.copy(), traj_capacity=400)
    assert (base[3] == 0).all() and base[2]["n_rejected"].sum() > 0
    for slice_attempts, max_ctas in ((7, 2), (1, 3), (64, 1), (5, 6)):   # CTAs of two set contexts each: 4, 6, 2 contexts for 7 sets; all resident
        eng.set_tx_tuning(slice_attempts, max_ctas)
        got = eng.propagate_batch(st, cs, ep, end, step_ns=step0.copy(), traj_capacity=400)
        assert np.array_equal(got[0], base[0]) and np.array_equal(got[1], base[1]) and np.array_equal(got[3], base[3])
        for f in ("step_ns", "error", "attempts", "n_steps", "n_rejected", "n_rhs"):
            assert np.array_equal(got[2][f], base[2][f]), f
        assert np.array_equal(got[4][2], base[4][2])
        for i in range(n):
            k = int(base[4][2][i])
            assert np.array_equal(got[4][0][:k, i], base[4][0][:k, i]) and np.array_equal(got[4][1][:, :k, i], base[4][1][:, :k, i])
    eng.set_tx_tuning(64, 0)
    # the recording against the oracle's stream
    ref = oracle_run(oracle, prop, nb.EARTH_J2000, None, st, cs, ep, end, step_ns=step0.copy())
    assert max_dr_dv(base[0], ref[0])[0] < 5e-7


def test_transposed_kernel_backward_events_and_statuses(oracle):
    n = 96
    mc, (st, cs, ep) = leo_ensemble(n, seed=5)
    dyn, _ = _leo_dyn(12)
    prop = nb.Propagator.default(dyn, mode=nb.MODE_FAST)
    eng = _tx_engine(prop)
    eng.set_tx_tuning(9, 1)   # one CTA (two set contexts), three sets: every slice boundary is exercised as well
    # forward, then backward to the start
    fwd, fep, _, fst = eng.propagate_batch(st, cs, ep, 3 * 3600 * S)
    back, bep, _, bst = eng.propagate_batch(fwd, cs, fep, 0)
    assert (fst == 0).all() and (bst == 0).all() and (bep == 0).all()
    assert np.sqrt(((back[:3] - st[:3]) ** 2).sum(0)).max() < 1e-6
    rb = oracle_run(oracle, prop, nb.EARTH_J2000, None, fwd, cs, fep, 0)
    assert max_dr_dv(back, rb[0])[0] < 5e-7
    # stop condition: second apoapsis/periapsis crossing (r.v = 0), recorded; same stop step as the oracle
    ev = (nb.abi.EVENT_RDOTV if hasattr(nb.abi, "EVENT_RDOTV") else 2, 0.0, 2)
    got = eng.propagate_batch(st, cs, ep, 6 * 3600 * S, traj_capacity=300, event=ev)
    packed = dyn.pack(nb.EARTH_J2000, None)
    want = oracle.propagate_batch(packed.c, prop.opts.to_c(prop.method), st, cs, ep, 6 * 3600 * S, traj_capacity=300, event=ev)
    assert np.array_equal(got[3], want[3]) and np.array_equal(got[5], want[5])
    assert np.abs(got[2]["n_steps"] - want[2]["n_steps"]).max() <= 1
    # the run stops at the end of the bracketing STEP; FAST adapts its steps ~1e-5 differently, so the stop epochs differ by tens of
    # milliseconds after ~60 steps and the states by v dt: compare along the orbit
    dt_s = (got[1] - want[1]) * 1e-9
    assert np.abs(dt_s).max() < 0.5
    drift = np.linalg.norm(got[0][:3] - (want[0][:3] + want[0][3:6] * dt_s[None, :]), axis=0)
    assert drift.max() < 1e-3, drift.max()   # second order in dt: |a| dt^2 / 2 ~ 1e-5 km
    # zero-length span and negative propellant mass (FuelExhausted, spacecraft.rs:163-168)
    st2 = st.copy(); st2[8, 3] = -1.0
    ep2 = ep.copy(); ep2[5] = 3600 * S
    o2, e2, d2, s2 = eng.propagate_batch(st2, cs, ep2, 3600 * S)
    r2 = oracle_run(oracle, prop, nb.EARTH_J2000, None, st2, cs, ep2, 3600 * S)
    assert np.array_equal(s2, r2[3]) and s2[3] == 2 and d2["n_steps"][5] == 0 and np.array_equal(o2[:, 5], st2[:, 5])


def test_automatic_dispatch_and_set_spreading(oracle):
    """The library's own choice: FAST ensembles with a field of degree 8-40 go to the transposed kernel from 1 024 trajectories
    (below: the lane-cooperative kernel; STRICT never).  1 100 trajectories = 35 sets on 35 CTAs (one set per CTA, the second set
    context of every CTA finds nothing and leaves): same bits as the kernel forced explicitly, parity with the oracle."""
    n = 1100
    mc, (st, cs, ep) = leo_ensemble(n, seed=17)
    dyn, _ = _leo_dyn()
    end = 2 * 3600 * S
    prop = nb.Propagator.default(dyn, mode=nb.MODE_FAST)
    eng = prop.engine(nb.EARTH_J2000, None)
    out, out_ep, det, status = eng.propagate_batch(st, cs, ep, end)
    assert eng.last_kernel() == nb.KERNEL_TRANSPOSED
    small = eng.propagate_batch(st[:, :1000].copy(), cs[:, :1000].copy(), ep[:1000].copy(), end)[0]
    assert eng.last_kernel() == nb.KERNEL_COOP
    assert max_dr_dv(small, out[:, :1000])[0] < 1e-6
    eng.set_kernel(nb.KERNEL_TRANSPOSED)
    forced = eng.propagate_batch(st, cs, ep, end)[0]
    assert np.array_equal(forced, out)
    ref, ref_ep, ref_det, ref_status = oracle_run(oracle, prop, nb.EARTH_J2000, None, st, cs, ep, end)
    assert np.array_equal(status, ref_status) and np.array_equal(out_ep, ref_ep)
    assert max_dr_dv(out, ref)[0] < 5e-7
    strict = nb.Propagator.default(dyn, mode=nb.MODE_STRICT).engine(nb.EARTH_J2000, None)
    strict.propagate_batch(st, cs, ep, 600 * S)
    assert strict.last_kernel() != nb.KERNEL_TRANSPOSED
