"""`IntegratorOptions.integration_frame` (options.rs:60; instance.rs:117-142, 167-176, 211-220) and harmonic fields of another
body than the integration centre / several fields at once (gravity_field.rs:149-154, orbital.rs:44-46), on the CPU oracle and
through the host mirror (the GPU parity of both is in tests/test_gpu_frames_fields.py).

The reference's own check (tests/mission_design/force_models.rs:181-212) is restated: a state handed over in the Moon frame with
`integration_frame = EME2000` must give the same trajectory as the Earth-frame propagation of the same spacecraft."""
import ctypes as C

import numpy as np
import pytest

import nyx_b200 as nb
from nyx_b200 import abi
from nyx_b200.frames import EARTH, MOON, SUN
from tests.util import S, use_oracle_engine

DAY = 86400 * S


def _eom(oracle, packed, t_ns, y9):
    dy = np.zeros(9)
    consts = np.array([100.0, 0.0, 1.0, 1.0])
    y9 = np.ascontiguousarray(y9, dtype=np.float64)
    assert oracle.lib().nyx_oracle_eom(C.byref(packed.c), int(t_ns), 0.0, abi.as_double_p(y9), abi.as_double_p(consts), abi.as_double_p(dy)) == 0
    return dy[3:6]


def _body_state(oracle, body_c, t_ns):
    p, v = np.zeros(3), np.zeros(3)
    L = oracle.lib()
    L.nyx_oracle_body_velocity.restype = C.c_int
    L.nyx_oracle_body_velocity.argtypes = [C.POINTER(abi.BodyC), C.c_int64, abi.c_double_p]
    assert L.nyx_oracle_body_position(C.byref(body_c), int(t_ns), abi.as_double_p(p)) == 0
    assert L.nyx_oracle_body_velocity(C.byref(body_c), int(t_ns), abi.as_double_p(v)) == 0
    return p, v


def test_field_of_another_body_is_evaluated_about_that_body(oracle):
    """Earth-centred state near the Moon, lunar 20x20 field: the field's acceleration equals what the SAME field gives as the
    central field of a Moon-centred state at the Moon-relative position (the vector is not transformed, gravity_field.rs:258-267);
    two fields add up."""
    alm_e = nb.Almanac.synthetic(nb.EARTH_J2000, 0, 4.0, bodies=(MOON, SUN))
    luna = nb.GravityField.new(nb.GravityFieldData.from_fixture("luna_jggrx_80x80", 20, 20, nb.IAU_MOON_FRAME))
    terra = nb.GravityField.new(nb.GravityFieldData.from_fixture("jgm3_70x70", 8, 8, nb.IAU_EARTH_FRAME))
    t = 3 * 3600 * S
    moon_c = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.two_body()).pack(nb.EARTH_J2000, alm_e).c.bodies[alm_e.body_index(MOON)]
    p_moon, _ = _body_state(oracle, moon_c, t)
    rel = np.array([1200.0, -900.0, 1100.0])
    y_e = np.concatenate([p_moon + rel, [0.1, 1.0, 0.2], [1.8, 2.2, 0.0]])
    two_e = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.two_body()).pack(nb.EARTH_J2000, alm_e)
    luna_e = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(luna)).pack(nb.EARTH_J2000, alm_e)
    assert luna_e.c.gravity[0].body == alm_e.body_index(MOON) and luna_e.c.n_gravity == 1
    both_e = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.new([terra, luna])).pack(nb.EARTH_J2000, alm_e)
    assert both_e.c.n_gravity == 2 and both_e.c.gravity[0].body == abi.NYXB_CENTRAL_BODY and both_e.c.gravity[1].body == alm_e.body_index(MOON)
    terra_e = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(terra)).pack(nb.EARTH_J2000, alm_e)
    a_two = _eom(oracle, two_e, t, y_e)
    a_luna = _eom(oracle, luna_e, t, y_e) - a_two
    a_terra = _eom(oracle, terra_e, t, y_e) - a_two
    a_both = _eom(oracle, both_e, t, y_e) - a_two
    # the same field as the central one of a Moon-centred state
    y_m = np.concatenate([(p_moon + rel) - p_moon, y_e[3:]])
    two_m = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.two_body()).pack(nb.MOON_J2000, None)
    luna_m = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(luna)).pack(nb.MOON_J2000, None)
    a_ref = _eom(oracle, luna_m, t, y_m) - _eom(oracle, two_m, t, y_m)
    scale = np.abs(_eom(oracle, two_m, t, y_m)).max()
    assert np.abs(a_luna - a_ref).max() < 1e-13 * scale     # differences of accelerations ~1e3 x larger: rounding of the subtraction
    assert np.abs(a_both - (a_terra + a_luna)).max() < 1e-13 * np.abs(a_two).max() + 1e-13 * scale
    assert np.linalg.norm(a_luna) > 1e-4 * np.linalg.norm(_eom(oracle, two_m, t, y_m))   # the field matters at 1 860 km from the centre of the Moon


def test_point_mass_order_is_the_callers(oracle):
    alm = nb.Almanac.synthetic(nb.EARTH_J2000, 0, 2.0, bodies=(SUN, MOON))
    a = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.point_masses([MOON, SUN])).pack(nb.EARTH_J2000, alm)
    b = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.point_masses([SUN, MOON])).pack(nb.EARTH_J2000, alm)
    assert list(a.c.point_mass_order[:2]) == [alm.body_index(MOON), alm.body_index(SUN)] and a.c.n_point_masses == 2
    assert list(b.c.point_mass_order[:2]) == [alm.body_index(SUN), alm.body_index(MOON)] and a.c.point_mass_mask == b.c.point_mass_mask
    y = np.array([42000.0, 1000.0, -500.0, 0.1, 3.0, 0.2, 1.8, 2.2, 0.0])
    da, db = _eom(oracle, a, 0, y), _eom(oracle, b, 0, y)
    assert np.abs(da - db).max() < 1e-18 and np.abs(da - db).max() >= 0.0   # same physics; summation order may change the last bit


def test_almanac_centre_must_be_the_integration_frame():
    alm_e = nb.Almanac.synthetic(nb.EARTH_J2000, 0, 2.0)
    dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.point_masses([EARTH, SUN]))
    with pytest.raises(nb.DynamicsError, match="centred on"):
        dyn.pack(nb.MOON_J2000, alm_e)


def _jwst_like(frame, alm):
    srp = nb.SolarPressure.new([nb.EARTH_J2000, nb.MOON_J2000], alm)
    return nb.SpacecraftDynamics.from_model(nb.OrbitalDynamics.point_masses([MOON, SUN]), srp)


def test_integration_frame_swap_reproduces_the_earth_frame_run(oracle, monkeypatch):
    """force_models.rs:181-212: `setup_moon.opts.integration_frame = Some(eme2k)` on a Moon-frame state gives a result in the Moon
    frame which, transformed to the Earth frame, equals the plain Earth-frame run."""
    alm = nb.Almanac.synthetic(nb.EARTH_J2000, 0, 6.0, bodies=(MOON, SUN))
    dyn = _jwst_like(nb.EARTH_J2000, alm)
    orbit = nb.Orbit.keplerian(24396.0, 0.35, 30.0, 60.0, 60.0, 180.0, 0, nb.EARTH_J2000)   # the MEO of force_models.rs:128-133
    sc_e = nb.Spacecraft(orbit=orbit, mass=nb.Mass(300.0, 0.0, 0.0), srp=nb.SRPData(16.0, 1.8))
    prop_e = nb.Propagator.default(dyn)
    use_oracle_engine(monkeypatch, oracle, prop_e)
    fin_e = prop_e.with_(sc_e, alm).for_duration(2 * DAY)
    # the same spacecraft expressed in the Moon frame at t0
    moon_c = dyn.pack(nb.EARTH_J2000, alm).c.bodies[alm.body_index(MOON)]
    p0, v0 = _body_state(oracle, moon_c, 0)
    x = orbit.to_cartesian_pos_vel()
    sc_m = nb.Spacecraft(orbit=nb.Orbit.cartesian(*(x[:3] - p0), *(x[3:] - v0), 0, nb.MOON_J2000), mass=sc_e.mass, srp=sc_e.srp)
    prop_m = nb.Propagator.default(dyn)
    prop_m.opts.integration_frame = nb.EARTH_J2000
    use_oracle_engine(monkeypatch, oracle, prop_m)
    packed, opts_c = prop_m.lower(nb.MOON_J2000, alm)
    assert opts_c.state_center == alm.body_index(MOON) + 1 and packed.c.mu_central_km3_s2 == nb.EARTH_J2000.mu_km3_s2()
    fin_m = prop_m.with_(sc_m, alm).for_duration(2 * DAY)
    assert fin_m.orbit.frame.ephemeris_id == MOON and fin_m.epoch() == 2 * DAY        # "expected a result in the Moon frame"
    p1, v1 = _body_state(oracle, moon_c, 2 * DAY)
    back = fin_m.orbit.to_cartesian_pos_vel() + np.concatenate([p1, v1])
    want = fin_e.orbit.to_cartesian_pos_vel()
    # the translations round at |r_moon| ~ 4e5 km (ulp 6e-11 km); two days of adaptive MEO propagation turn that into a slightly
    # different step sequence (the sensitivity of profiles/r02_oracle_sensitivity_*.json): sub-mm
    assert np.abs(back[:3] - want[:3]).max() < 2e-6 and np.abs(back[3:] - want[3:]).max() < 1e-9
    # with a fixed step the two runs see the same step sequence: round-off only
    for pr in (prop_e, prop_m):
        pr.opts = nb.IntegratorOptions.with_fixed_step_s(120.0)
        pr.opts.integration_frame = nb.EARTH_J2000 if pr is prop_m else None
        use_oracle_engine(monkeypatch, oracle, pr)
    fe = prop_e.with_(sc_e, alm).for_duration(2 * DAY).orbit.to_cartesian_pos_vel()
    fm = prop_m.with_(sc_m, alm).for_duration(2 * DAY).orbit.to_cartesian_pos_vel() + np.concatenate([p1, v1])
    assert np.abs(fm[:3] - fe[:3]).max() < 1e-8 and np.abs(fm[3:] - fe[3:]).max() < 1e-12
    # integration_frame equal to the state's frame is a no-op (instance.rs:119: `integration_frame != self.state.orbit().frame`)
    prop_same = nb.Propagator.default(dyn)
    prop_same.opts.integration_frame = nb.EARTH_J2000
    assert prop_same.lower(nb.EARTH_J2000, alm)[1].state_center == 0
    # a frame the almanac cannot relate to the integration frame is an error (DynamicsAlmanacError in the reference)
    with pytest.raises(nb.PropagationError, match="no ephemeris"):
        prop_m.lower(nb.Frame("Mars J2000", 499, 42828.37, 3396.19), alm)
