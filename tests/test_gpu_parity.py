"""GPU parity tests proper: every call goes through the C ABI (libnyxb.so) and is compared with the
CPU oracle on the same seeded inputs, or with the reference's golden vectors."""
import numpy as np
import pytest

import nyx_b200 as nb
from tests.util import GOLDEN, S, leo_ensemble, leo_state, max_dr_dv, opts_from_json, oracle_run

pytestmark = pytest.mark.gpu
DAY = 86400 * S


def two_body():
    return nb.SpacecraftDynamics.new(nb.OrbitalDynamics.two_body())


@pytest.mark.parametrize("case", GOLDEN["two_body"], ids=lambda c: c["id"])
def test_golden_vectors_strict(case):
    """The reference's golden vectors through the CUDA path, strict mode: bit-exact where the reference is."""
    frame = nb.EARTH_J2000.with_mu_km3_s2(case["mu"])
    prop = nb.Propagator.new(two_body(), nb.IntegratorMethod[case["method"]], opts_from_json(case["opts"]), mode=nb.MODE_STRICT)
    inst = prop.with_(leo_state(frame))
    final = inst.for_duration(1 * nb.Unit.Day)
    got = final.orbit.to_cartesian_pos_vel()
    gold = np.array(case["final"])
    if case["tol_km"] == 0.0:
        assert np.array_equal(got, gold), (case["id"], got - gold)
    else:
        assert np.abs(got - gold).max() < case["tol_km"]
    if "n_steps" in case:
        assert inst.latest_details().n_steps == case["n_steps"]


@pytest.mark.parametrize("case", GOLDEN["two_body"], ids=lambda c: c["id"])
def test_golden_vectors_fast(case):
    """Fast (FMA) mode: within the reference's own tolerance class (1e-7 km) of the golden vectors."""
    frame = nb.EARTH_J2000.with_mu_km3_s2(case["mu"])
    prop = nb.Propagator.new(two_body(), nb.IntegratorMethod[case["method"]], opts_from_json(case["opts"]), mode=nb.MODE_FAST)
    got = prop.with_(leo_state(frame)).for_duration(1 * nb.Unit.Day).orbit.to_cartesian_pos_vel()
    # RK4 @1 s: 86 400 steps of round-off; low-order adaptive methods (CK45/DP45) are chaotic in the step
    # sequence at the 1e-6 km level, the 8(9)/7(8) pairs are not.
    tol = {"G7": 2e-5, "G8": 5e-6, "G5a": 5e-6, "G6a": 1e-6}.get(case["id"], 1e-7)
    assert np.abs(got - np.array(case["final"])).max() < tol


def _ensemble_vs_oracle(oracle, prop, frame, almanac, st, cs, ep, end):
    eng = prop.engine(frame, almanac)
    out, out_ep, det, status = eng.propagate_batch(st, cs, ep, end)
    ref, ref_ep, ref_det, ref_status = oracle_run(oracle, prop, frame, almanac, st, cs, ep, end)
    assert np.array_equal(status, ref_status)
    assert np.array_equal(out_ep, ref_ep)
    return out, det, ref, ref_det


def test_two_body_ensemble_strict_bitexact(oracle):
    mc, (st, cs, ep) = leo_ensemble(512, seed=1)
    prop = nb.Propagator.default(two_body(), mode=nb.MODE_STRICT)
    out, det, ref, ref_det = _ensemble_vs_oracle(oracle, prop, nb.EARTH_J2000, None, st, cs, ep, DAY // 4)
    same = (out == ref).all(axis=0)
    # CUDA pow vs glibc pow may flip an ns truncation of the adapted step on rare trajectories
    assert same.mean() >= 0.98, same.mean()
    dr, dv = max_dr_dv(out, ref)
    assert dr < 1e-9 and dv < 1e-12
    assert np.array_equal(det["n_steps"], ref_det["n_steps"])


@pytest.mark.parametrize("degree,order,lanes", [(2, 2, 1), (8, 8, 1), (21, 21, 1), (8, 8, 8), (21, 21, 8), (21, 21, 16), (21, 21, 32),
                                                (12, 7, 8), (40, 40, 16)])
def test_harmonics_ensemble_strict_bitexact(oracle, degree, order, lanes):
    """Strict mode reproduces the oracle's harmonic sums bit for bit (deterministic sin/cos, no FMA), both with the
    per-thread kernel (lanes = 1) and with the lane-cooperative STRICT kernel (columns -> shared triangle -> rows)."""
    mc, (st, cs, ep) = leo_ensemble(128 if degree <= 21 else 32, seed=2)
    gd = nb.GravityFieldData.from_fixture("jgm3_70x70", degree, order, nb.IAU_EARTH_FRAME)
    dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(nb.GravityField.new(gd)))
    prop = nb.Propagator.default(dyn, mode=nb.MODE_STRICT)
    prop.engine(nb.EARTH_J2000, None).set_lanes(lanes)
    assert prop.engine(nb.EARTH_J2000, None).lanes() == lanes
    out, det, ref, ref_det = _ensemble_vs_oracle(oracle, prop, nb.EARTH_J2000, None, st, cs, ep, 3 * 3600 * S)
    same = (out == ref).all(axis=0)
    assert same.mean() >= 0.98, same.mean()
    dr, dv = max_dr_dv(out, ref)
    assert dr < 1e-9, dr


@pytest.mark.parametrize("degree", [2, 21])
def test_harmonics_ensemble_fast_tolerance(oracle, degree):
    """Fast mode: sub-mm (north-star tolerance 1e-6 km).  FMA contraction perturbs the error estimate by
    ~1e-4 relative, hence the adapted step sequence, hence the result at the integrator's own
    truncation-error level (~1e-7 km after 6 h): we assert 5e-7 km."""
    mc, (st, cs, ep) = leo_ensemble(256, seed=3)
    gd = nb.GravityFieldData.from_fixture("jgm3_70x70", degree, degree, nb.IAU_EARTH_FRAME)
    dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(nb.GravityField.new(gd)))
    prop = nb.Propagator.default(dyn, mode=nb.MODE_FAST)
    out, det, ref, ref_det = _ensemble_vs_oracle(oracle, prop, nb.EARTH_J2000, None, st, cs, ep, 6 * 3600 * S)
    dr, dv = max_dr_dv(out, ref)
    assert dr < 5e-7 and dv < 1e-9, (dr, dv)
    assert np.abs(det["n_steps"] - ref_det["n_steps"]).max() <= 1


@pytest.mark.parametrize("lanes,degree,order", [(8, 21, 21), (16, 21, 21), (32, 21, 21), (8, 12, 7), (16, 40, 40), (32, 70, 70)])
def test_cooperative_kernel_vs_oracle(oracle, lanes, degree, order):
    """Lane-cooperative kernel (G lanes per trajectory, column-split harmonic sum) against the oracle.
    6 h of adaptive RK89: differences are bounded by the integrator's own step-sequence sensitivity
    (~1e-7 km); the north-star bound is 1e-6 km."""
    n = 96 if degree <= 40 else 24
    mc, (st, cs, ep) = leo_ensemble(n, seed=11)
    gd = nb.GravityFieldData.from_fixture("jgm3_70x70", degree, order, nb.IAU_EARTH_FRAME)
    dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(nb.GravityField.new(gd)))
    prop = nb.Propagator.default(dyn, mode=nb.MODE_FAST)
    eng = prop.engine(nb.EARTH_J2000, None)
    eng.set_lanes(lanes)
    assert eng.lanes() == lanes
    out, out_ep, det, status = eng.propagate_batch(st, cs, ep, 6 * 3600 * S)
    ref, ref_ep, ref_det, ref_status = oracle_run(oracle, prop, nb.EARTH_J2000, None, st, cs, ep, 6 * 3600 * S)
    assert (status == 0).all() and np.array_equal(out_ep, ref_ep)
    dr, dv = max_dr_dv(out, ref)
    assert dr < 5e-7 and dv < 1e-9, (dr, dv)
    assert np.abs(det["n_steps"] - ref_det["n_steps"]).max() <= 1
    # cooperative vs per-thread fast kernel on the same engine: same algorithm class
    eng.set_lanes(1)
    out1, _, det1, _ = eng.propagate_batch(st, cs, ep, 6 * 3600 * S)
    assert max_dr_dv(out, out1)[0] < 5e-7


@pytest.mark.parametrize("n", [1, 2, 33, 64])
def test_cooperative_kernel_ragged_sizes_and_rejections(oracle, n):
    """Ensemble sizes that leave lane groups of a warp without a trajectory; trajectories of a warp end at different step counts
    and retry rejected attempts independently."""
    mc, (st, cs, ep) = leo_ensemble(n, seed=13)
    ep = ep + (np.arange(n, dtype=np.int64) % 5) * 600 * S  # different start epochs => different step counts
    gd = nb.GravityFieldData.from_fixture("jgm3_70x70", 21, 21, nb.IAU_EARTH_FRAME)
    dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(nb.GravityField.new(gd)))
    end = 4 * 3600 * S
    for opts, tol in ((nb.IntegratorOptions.default(), 5e-7), (nb.IntegratorOptions.with_fixed_step_s(45.0), 5e-9),
                      # forced rejections: a too-large initial step with a tight tolerance
                      (nb.IntegratorOptions(init_step=600 * nb.Unit.Second, tolerance=1e-13), 5e-7)):
        prop = nb.Propagator.rk89(dyn, opts, mode=nb.MODE_FAST)
        eng = prop.engine(nb.EARTH_J2000, None)
        eng.set_lanes(8)
        out, out_ep, det, status = eng.propagate_batch(st, cs, ep, end)
        ref, ref_ep, ref_det, ref_status = oracle_run(oracle, prop, nb.EARTH_J2000, None, st, cs, ep, end)
        assert np.array_equal(status, ref_status) and np.array_equal(out_ep, ref_ep)
        assert max_dr_dv(out, ref)[0] < tol, (opts, max_dr_dv(out, ref))
        assert np.abs(det["n_steps"] - ref_det["n_steps"]).max() <= 1
        assert np.abs(det["n_rejected"] - ref_det["n_rejected"]).max() <= 1


def test_cooperative_kernel_fixed_step_tight(oracle):
    """With a FIXED step the step sequence cannot diverge, so the cooperative kernel must agree with the
    oracle to round-off (5e-9 km over 6 h), which pins the regrouped harmonic sum itself."""
    mc, (st, cs, ep) = leo_ensemble(64, seed=12)
    gd = nb.GravityFieldData.from_fixture("jgm3_70x70", 21, 21, nb.IAU_EARTH_FRAME)
    almanac = nb.Almanac.synthetic(nb.EARTH_J2000, 0, 1.0)
    orb = nb.OrbitalDynamics.new([nb.PointMasses.new([nb.MOON, nb.SUN]), nb.GravityField.new(gd)])
    dyn = nb.SpacecraftDynamics.new(orb)
    for lanes in (8, 16, 32):
        prop = nb.Propagator.rk89(dyn, nb.IntegratorOptions.with_fixed_step_s(60.0), mode=nb.MODE_FAST)
        eng = prop.engine(nb.EARTH_J2000, almanac)
        eng.set_lanes(lanes)
        out, out_ep, det, status = eng.propagate_batch(st, cs, ep, 6 * 3600 * S)
        ref, ref_ep, ref_det, _ = oracle_run(oracle, prop, nb.EARTH_J2000, almanac, st, cs, ep, 6 * 3600 * S)
        assert (status == 0).all() and np.array_equal(det["n_steps"], ref_det["n_steps"])
        dr, dv = max_dr_dv(out, ref)
        assert dr < 5e-9 and dv < 5e-12, (lanes, dr, dv)


@pytest.mark.parametrize("mode", [nb.MODE_STRICT, nb.MODE_FAST])
def test_third_body_srp_ensemble(oracle, mode):
    """JWST-like config: Sun+Moon point masses + SRP with Earth & Moon shadows (examples/02_jwst…/main.rs:99-146)."""
    almanac = nb.Almanac.synthetic(nb.EARTH_J2000, 0, 12.0)
    frame = nb.EARTH_J2000
    orbit = nb.Orbit.cartesian(119901.070276, -1389299.665421, -1041369.150539, 0.045956, -0.013168, 0.034535, 0, frame)
    template = nb.Spacecraft(orbit=orbit, mass=nb.Mass(6200.0, 0.0, 0.0), srp=nb.SRPData(21.197 * 14.162, 1.56))
    mvn = nb.MvnSpacecraft.from_cartesian_std(template, 0.5, 1e-4)
    mc = nb.MonteCarlo(template, mvn, "jwst", seed=4)
    st, cs, ep = nb.pack_spacecraft(ds.state for _, ds in mc.generate_states(0, 256))
    srp = nb.SolarPressure.new([nb.EARTH_J2000, nb.MOON_J2000], almanac)
    dyn = nb.SpacecraftDynamics.from_model(nb.OrbitalDynamics.point_masses([nb.MOON, nb.SUN]), srp)
    prop = nb.Propagator.default(dyn, mode=mode)
    out, det, ref, ref_det = _ensemble_vs_oracle(oracle, prop, frame, almanac, st, cs, ep, 10 * DAY)
    dr, dv = max_dr_dv(out, ref)
    assert dr < 1e-6 and dv < 1e-11, (dr, dv)  # |r| ~ 1.7e6 km: 1e-6 km is < 1e-12 relative


@pytest.mark.parametrize("mode,lanes", [(nb.MODE_FAST, 32), (nb.MODE_FAST, 16), (nb.MODE_STRICT, 1)])
def test_cislunar_70x70_third_body(oracle, mode, lanes):
    """BASELINE config 4 shape: low lunar orbit, GRAIL 70x70 + Earth/Sun point masses, Moon-centred, IAU Moon rotation."""
    from nyx_b200.frames import EARTH

    moon = nb.MOON_J2000
    almanac = nb.Almanac.synthetic(moon, 0, 2.0, bodies=(EARTH, nb.SUN))
    gd = nb.GravityFieldData.from_fixture("luna_jggrx_80x80", 70, 70, nb.IAU_MOON_FRAME)
    orb = nb.OrbitalDynamics.new([nb.PointMasses.new([EARTH, nb.SUN]), nb.GravityField.new(gd)])
    dyn = nb.SpacecraftDynamics.new(orb)
    orbit = nb.Orbit.keplerian(1737.4 + 100.0, 0.001, 90.0, 10.0, 0.0, 0.0, 0, moon)
    template = nb.Spacecraft(orbit=orbit, mass=nb.Mass(1000.0, 0.0, 0.0))
    mc = nb.MonteCarlo(template, nb.MvnSpacecraft.from_cartesian_std(template, 0.1, 1e-4), "llo", seed=31)
    st, cs, ep = nb.pack_spacecraft(ds.state for _, ds in mc.generate_states(0, 12))
    prop = nb.Propagator.default(dyn, mode=mode)
    eng = prop.engine(moon, almanac)
    if lanes > 1:
        eng.set_lanes(lanes)
    end = 4 * 3600 * S
    out, out_ep, det, status = eng.propagate_batch(st, cs, ep, end)
    ref, ref_ep, ref_det, ref_status = oracle_run(oracle, prop, moon, almanac, st, cs, ep, end)
    assert (status == 0).all() and np.array_equal(status, ref_status) and np.array_equal(out_ep, ref_ep)
    dr, dv = max_dr_dv(out, ref)
    if mode == nb.MODE_STRICT:
        assert (out == ref).all(axis=0).mean() >= 0.9 and dr < 1e-9, dr
    else:
        assert dr < 5e-7 and dv < 1e-9, (dr, dv)


@pytest.mark.parametrize("density", ["constant", "exponential", "stdatm"])
def test_drag_and_leo_eclipse_ensemble(oracle, density):
    """LEO with drag (3 density models, drag.rs:181-284) + SRP through Earth umbra/penumbra."""
    almanac = nb.Almanac.synthetic(nb.EARTH_J2000, 0, 2.0)
    dens = {"constant": nb.AtmDensity.Constant(1e-12), "exponential": nb.AtmDensity.earth_exponential(),
            "stdatm": nb.AtmDensity.StdAtm(1_000_000.0)}[density]
    drag = nb.Drag(dens, nb.IAU_EARTH_FRAME)
    srp = nb.SolarPressure.new([nb.EARTH_J2000], almanac)
    gd = nb.GravityFieldData.from_fixture("jgm3_70x70", 4, 4, nb.IAU_EARTH_FRAME)
    orb = nb.OrbitalDynamics.new([nb.PointMasses.new([nb.MOON, nb.SUN]), nb.GravityField.new(gd)])
    dyn = nb.SpacecraftDynamics.from_models(orb, [srp, drag])
    mc, (st, cs, ep) = leo_ensemble(128, seed=5, srp=nb.SRPData(16.0, 1.8), drag=nb.DragData(16.0, 2.2),
                                    mass=nb.Mass(300.0, 0.0, 0.0))
    prop = nb.Propagator.default(dyn, mode=nb.MODE_FAST)
    out, det, ref, ref_det = _ensemble_vs_oracle(oracle, prop, nb.EARTH_J2000, almanac, st, cs, ep, 4 * 3600 * S)
    dr, dv = max_dr_dv(out, ref)
    assert dr < 1e-6 and dv < 2e-9, (dr, dv)  # north-star bound; see DESIGN.md §3 on step-sequence sensitivity


@pytest.mark.parametrize("ctrl", list(nb.ErrorControl))
def test_all_error_controls_strict(oracle, ctrl):
    mc, (st, cs, ep) = leo_ensemble(64, seed=6)
    opts = nb.IntegratorOptions.with_adaptive_step_s(0.1, 120.0, 1e-10, ctrl)
    prop = nb.Propagator.new(two_body(), nb.IntegratorMethod.DormandPrince78, opts, mode=nb.MODE_STRICT)
    out, det, ref, ref_det = _ensemble_vs_oracle(oracle, prop, nb.EARTH_J2000, None, st, cs, ep, 2 * 3600 * S)
    assert (out == ref).all(axis=0).mean() >= 0.95
    assert max_dr_dv(out, ref)[0] < 1e-8


def test_backward_and_repeated_calls_match_oracle(oracle):
    """PropInstance semantics: adapted step carried between calls, back-propagation (instance.rs:112-115,198-200)."""
    frame = nb.EARTH_J2000
    mc, (st, cs, ep) = leo_ensemble(32, seed=7)
    prop = nb.Propagator.default(two_body(), mode=nb.MODE_STRICT)
    eng = prop.engine(frame, None)
    step_g = np.full(32, prop.opts.init_step, dtype=np.int64)
    step_o = step_g.copy()
    cur_g, ep_g, cur_o, ep_o = st, ep, st, ep
    for target in (3600 * S, -1800 * S, 7200 * S, 7200 * S, 0):
        cur_g, ep_g, _, sg = eng.propagate_batch(cur_g, cs, ep_g, target, step_g)
        cur_o, ep_o, _, so = oracle_run(oracle, prop, frame, None, cur_o, cs, ep_o, target, step_o)
        assert np.array_equal(sg, so) and np.array_equal(ep_g, ep_o) and (ep_g == target).all()
        assert np.array_equal(step_g, step_o)
        assert max_dr_dv(cur_g, cur_o)[0] < 1e-9
    assert max_dr_dv(cur_g, st)[0] < 1e-5  # round trip (orbitaldyn.rs:139-151)


def test_edge_cases_and_error_statuses(oracle):
    frame = nb.EARTH_J2000
    almanac = nb.Almanac.synthetic(frame, 0, 1.0, pad_days=0.5)
    srp = nb.SolarPressure.new([frame], almanac)
    dyn = nb.SpacecraftDynamics.from_model(nb.OrbitalDynamics.point_masses([nb.SUN]), srp)
    prop = nb.Propagator.default(dyn, mode=nb.MODE_STRICT)
    eng = prop.engine(frame, almanac)
    # empty batch
    out, out_ep, det, status = eng.propagate_batch(np.empty((9, 0)), np.empty((4, 0)), np.empty(0, dtype=np.int64), DAY)
    assert out.shape == (9, 0) and status.shape == (0,)
    base = nb.Spacecraft(orbit=nb.Orbit.keplerian(7000.0, 0.01, 30.0, 0, 0, 0, 0, frame), mass=nb.Mass(100.0, 5.0, 0.0),
                         srp=nb.SRPData(10.0, 1.8))
    import dataclasses as dc

    cases = [
        base,                                                              # 0 ok
        dc.replace(base, mass=nb.Mass(100.0, -1.0, 0.0)),                  # 1 FuelExhausted (spacecraft.rs:163-168)
        dc.replace(base, mass=nb.Mass(0.0, 0.0, 0.0)),                     # 2 MasslessSpacecraft (:201-203)
        # 3: NaN state.  With force models the reference reports MasslessSpacecraft: the retry step is NaN, and
        #    prop_mass + h*0 = NaN fails `mass_kg() > 0` (spacecraft.rs:201-203) before the NaN check (instance.rs:432-439)
        dc.replace(base, orbit=dc.replace(base.orbit, x_km=float("nan"))),
        dc.replace(base, srp=nb.SRPData(10.0, 5.0)),                       # 4 Cr clamped to 2 (cosmic/spacecraft.rs:494)
        dc.replace(base, orbit=dc.replace(base.orbit, epoch_ns=3600 * S)), # 5 zero duration: returned untouched
    ]
    st, cs, ep = nb.pack_spacecraft(cases)
    end = 3600 * S
    out, out_ep, det, status = eng.propagate_batch(st, cs, ep, end)
    ref, ref_ep, ref_det, ref_status = oracle_run(oracle, prop, frame, almanac, st, cs, ep, end)
    assert list(status) == [0, nb.abi.ERR_FUEL_EXHAUSTED, nb.abi.ERR_MASSLESS, nb.abi.ERR_MASSLESS, 0, 0]
    assert np.array_equal(status, ref_status) and np.array_equal(out_ep, ref_ep)
    assert out[6, 4] == 2.0 and out[6, 0] == 1.8
    assert np.array_equal(out[:, 5], st[:, 5]) and det["n_steps"][5] == 0
    # SRP shadow geometry goes through libm acos/asin (CUDA vs glibc differ in the last ulp): step-sequence sensitivity applies
    assert np.abs(out[:, 0] - ref[:, 0]).max() < 5e-7
    # NaN state without force models -> PropMathError after the attempts are exhausted (instance.rs:432-439)
    prop2 = nb.Propagator.default(two_body(), mode=nb.MODE_STRICT)
    o2, _, _, s2 = prop2.engine(frame, None).propagate_batch(st[:, 3:4], cs[:, 3:4], ep[3:4], end)
    r2 = oracle_run(oracle, prop2, frame, None, st[:, 3:4], cs[:, 3:4], ep[3:4], end)
    assert s2[0] == nb.abi.ERR_PROP_MATH == r2[3][0]  # NaN check precedes the max-attempts warning (instance.rs:432-445)
    # outside ephemeris coverage -> almanac error status, not a crash
    out, out_ep, det, status = eng.propagate_batch(st[:, :1], cs[:, :1], ep[:1], 30 * DAY)
    assert status[0] == nb.abi.ERR_EPHEMERIS
    # the reference API raises per-run errors from PropInstance::until_epoch
    with pytest.raises(nb.PropagationError, match="FuelExhausted"):
        prop.with_(cases[1], almanac).until_epoch(end)
    # many_until_epoch drops failed runs (py_md.rs:251-254)
    assert len(prop.many_until_epoch(cases, end, almanac)) == 3


def test_monte_carlo_api_and_resume(oracle):
    """MonteCarlo::run_until_epoch / resume_run_until_epoch (montecarlo.rs:188-273): resume(skip) reproduces the tail."""
    mc, _ = leo_ensemble(8, seed=9)
    prop = nb.Propagator.default(two_body(), mode=nb.MODE_FAST)
    full = mc.run_until_epoch(prop, None, 1800 * S, 48)
    tail = mc.resume_run_until_epoch(prop, None, 40, 1800 * S, 8)
    assert len(full.runs) == 48 and [r.index for r in full.runs] == list(range(48))
    assert np.array_equal(full.final_state_soa[:, 40:], tail.final_state_soa)
    assert all(isinstance(r.result, nb.Spacecraft) for r in full.runs)
    assert full.total_steps() == int(full.details["n_steps"].sum()) > 48 * 20


def test_multi_device_call_shards_and_gathers(oracle):
    """`nyxb_propagate_batch_multi`: one ensemble over several engines (one per GPU; two engines on device 0 when the box has a
    single GPU — the sharding, strided uploads and the gather into the caller's [9][n] arrays are the same code)."""
    import torch
    from nyx_b200.dist import propagate_batch_multi

    n = 777   # not a multiple of anything
    mc, (st, cs, ep) = leo_ensemble(n, seed=31)
    gd = nb.GravityFieldData.from_fixture("jgm3_70x70", 12, 12, nb.IAU_EARTH_FRAME)
    dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(nb.GravityField.new(gd)))
    prop = nb.Propagator.default(dyn, mode=nb.MODE_STRICT)
    ndev = torch.cuda.device_count()
    for devices in ([0, 0, 0], list(range(ndev)) if ndev > 1 else [0, 0]):
        engs = prop.engines(nb.EARTH_J2000, None, devices)
        step = np.full(n, 60 * S, dtype=np.int64)
        out, oep, det, status = propagate_batch_multi(engs, st, cs, ep, 2 * 3600 * S, step_ns=step)
        one = prop.engine(nb.EARTH_J2000, None).propagate_batch(st, cs, ep, 2 * 3600 * S, step_ns=np.full(n, 60 * S, dtype=np.int64))
        assert np.array_equal(out, one[0]) and np.array_equal(oep, one[1]) and np.array_equal(status, one[3])
        assert np.array_equal(det["n_steps"], one[2]["n_steps"])
        for e in engs:
            e.close()
    res = nb.MonteCarlo(mc.nominal_state, mc.random_state, "multi", seed=31).run_until_epoch(prop, None, 3600 * S, 100, devices=[0, 0])
    ref = nb.MonteCarlo(mc.nominal_state, mc.random_state, "multi", seed=31).run_until_epoch(prop, None, 3600 * S, 100)
    assert all(np.array_equal(a.result.orbit.to_cartesian_pos_vel(), b.result.orbit.to_cartesian_pos_vel()) for a, b in zip(res.runs, ref.runs))
