"""Config / wire formats: the reference's serde layouts build the same host objects as the constructors."""
from pathlib import Path

import pytest

import nyx_b200 as nb

FIX = Path(__file__).parent / "fixtures"


def test_parse_duration_forms():
    assert nb.parse_duration("1 min") == 60 * 10**9
    assert nb.parse_duration("2 h 30 min") == 9000 * 10**9
    assert nb.parse_duration("0.5 s") == 5 * 10**8
    assert nb.parse_duration("1 ms") == 10**6
    assert nb.parse_duration(42) == 42
    with pytest.raises(ValueError):
        nb.parse_duration("3 fortnights")


def test_ground_stations_from_yaml():
    st = nb.load_ground_stations(FIX / "stations.yaml")
    assert list(st) == ["Station A", "Station B"]
    a, b = st["Station A"], st["Station B"]
    ref = nb.GroundStation.dss65_madrid(5.0, nb.StochasticNoise(5e-3), nb.StochasticNoise(50e-6, 1e-7))
    assert (a.latitude_deg, a.longitude_deg, a.height_km, a.elevation_mask_deg) == (ref.latitude_deg, ref.longitude_deg, ref.height_km, 5.0)
    assert a.stochastic_noises == ref.stochastic_noises and list(a.measurement_types) == [nb.MeasurementType.Range, nb.MeasurementType.Doppler]
    assert b.elevation_mask_deg == 7.5 and list(b.measurement_types) == [nb.MeasurementType.Doppler]
    ca = a.to_c(nb.EARTH_J2000, None)
    assert ca.n_types == 2 and abs(ca.noise_var[0] - 25e-6) < 1e-18 and ca.bias[1] == 1e-7 and ca.body == -1
    cb = b.to_c(nb.EARTH_J2000, None)
    assert cb.n_types == 1 and cb.types[0] == nb.abi.MSR_DOPPLER and abs(cb.noise_var[0] - 9e-12) < 1e-24


def test_propagator_config_builds_the_closed_model_set():
    cfg = nb.PropagatorConfig.load({
        "method": "DormandPrince78",
        "options": {"init_step": "30 s", "min_step": "1 ms", "max_step": "10 min", "tolerance": 1e-10, "attempts": 20, "error_ctrl": "RSSCartesianState"},
        "dynamics": {"accel_models": {"point_masses": {"celestial_objects": [301, 10], "correction": None}},
                     "force_models": {"solar_pressure": {"phi": 1367.0, "shadow_bodies": [399, 301], "estimate": False}}}})
    alm = nb.Almanac.synthetic(nb.EARTH_J2000, 0, 3.0)
    prop = cfg.build(alm)
    assert prop.method == nb.IntegratorMethod.DormandPrince78
    assert (prop.opts.init_step, prop.opts.min_step, prop.opts.max_step) == (30 * 10**9, 10**6, 600 * 10**9)
    assert prop.opts.tolerance == 1e-10 and prop.opts.attempts == 20 and prop.opts.error_ctrl == nb.ErrorControl.RSSCartesianState
    packed = prop.dynamics.pack(nb.EARTH_J2000, alm)
    assert packed.c.point_mass_mask != 0 and packed.c.srp.contents.n_shadow == 2 and packed.c.srp.contents.estimate == 0
    # same descriptor as the constructor route
    ref = nb.SpacecraftDynamics.from_model(nb.OrbitalDynamics.point_masses([nb.MOON, nb.SUN]),
                                           nb.SolarPressure.default_no_estimation([nb.EARTH_J2000, nb.MOON_J2000], alm)).pack(nb.EARTH_J2000, alm)
    assert packed.c.point_mass_mask == ref.c.point_mass_mask and packed.c.mu_central_km3_s2 == ref.c.mu_central_km3_s2
    with pytest.raises(nb.DynamicsError, match="SolidTides"):
        nb.PropagatorConfig.load({"dynamics": {"accel_models": {"solid_tides": {"k2": 0.3}}}}).build(alm)
    assert nb.PropagatorConfig.load({}).build().opts.init_step == 60 * 10**9
