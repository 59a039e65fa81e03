"""GPU parity (through the C ABI) of `integration_frame` and of harmonic fields that belong to another body than the integration
centre / several fields at once, against the CPU oracle: every kernel family, STRICT bit-level where the family has a STRICT build.
Reference: propagators/instance.rs:117-142, 167-176, 211-220; dynamics/gravity_field.rs:149-154; dynamics/orbital.rs:44-46, 213-247."""
import numpy as np
import pytest

import nyx_b200 as nb
from nyx_b200.frames import EARTH, MOON, SUN
from tests.util import S, max_dr_dv

pytestmark = pytest.mark.gpu
DAY = 86400 * S


def _cislunar(n, seed=3):
    """Earth-centred states in a low lunar orbit: what BASELINE configs[3] says literally (cislunar, lunar harmonics + third bodies)."""
    alm = nb.Almanac.synthetic(nb.EARTH_J2000, 0, 3.0, bodies=(MOON, SUN))
    moon_eng = nb.Almanac.synthetic(nb.EARTH_J2000, 0, 3.0, bodies=(MOON,)).bodies[0]
    p0 = moon_eng.position(0)
    h = 1.0   # velocity of the Moon from central differences of the same table (only used to build plausible initial states)
    v0 = (moon_eng.position(int(h * S)) - moon_eng.position(int(-h * S))) / (2 * h)
    llo = nb.Orbit.keplerian(1737.4 + 150.0, 0.002, 85.0, 10.0, 0.0, 0.0, 0, nb.MOON_J2000).to_cartesian_pos_vel()
    x = np.concatenate([llo[:3] + p0, llo[3:] + v0])
    template = nb.Spacecraft(orbit=nb.Orbit.cartesian(*x, 0, nb.EARTH_J2000), mass=nb.Mass(1000.0, 0.0, 0.0))
    mvn = nb.MvnSpacecraft.from_cartesian_std(template, 0.5, 1e-4)
    mc = nb.MonteCarlo(template, mvn, "cislunar", seed=seed)
    return alm, nb.pack_spacecraft(ds.state for _, ds in mc.generate_states(0, n))


def _fields(deg_moon, deg_earth=None):
    luna = nb.GravityField.new(nb.GravityFieldData.from_fixture("luna_jggrx_80x80", deg_moon, deg_moon, nb.IAU_MOON_FRAME))
    models = [nb.PointMasses.new([MOON, SUN]), luna]
    if deg_earth:
        models.append(nb.GravityField.new(nb.GravityFieldData.from_fixture("jgm3_70x70", deg_earth, deg_earth, nb.IAU_EARTH_FRAME)))
    return nb.SpacecraftDynamics.new(nb.OrbitalDynamics.new(models))


@pytest.mark.parametrize("mode,kernel,deg_moon,deg_earth", [
    (nb.MODE_STRICT, nb.KERNEL_THREAD, 12, None), (nb.MODE_STRICT, nb.KERNEL_COOP, 20, 4), (nb.MODE_FAST, nb.KERNEL_THREAD, 12, 4),
    (nb.MODE_FAST, nb.KERNEL_COOP, 20, 4), (nb.MODE_FAST, nb.KERNEL_COOP, 70, None), (nb.MODE_FAST, nb.KERNEL_TRANSPOSED, 20, 4)])
def test_earth_centred_cislunar_with_lunar_harmonics(oracle, mode, kernel, deg_moon, deg_earth):
    n = 40
    alm, (st, cs, ep) = _cislunar(n)
    dyn = _fields(deg_moon, deg_earth)
    prop = nb.Propagator.default(dyn, mode=mode)
    eng = prop.engine(nb.EARTH_J2000, alm)
    eng.set_kernel(kernel)
    end = 8 * 3600 * S   # four lunar orbits
    out, oep, det, status = eng.propagate_batch(st, cs, ep, end)
    assert eng.last_kernel() == kernel
    ref, rep, rdet, rstatus = oracle.propagate_batch(dyn.pack(nb.EARTH_J2000, alm).c, prop.opts.to_c(prop.method), st, cs, ep, end)
    assert (status == 0).all() and np.array_equal(status, rstatus) and np.array_equal(oep, rep)
    dr, dv = max_dr_dv(out, ref)
    if mode == nb.MODE_STRICT:
        same = (out == ref).all(axis=0)
        assert same.mean() >= 0.95 and dr < 1e-8, (same.mean(), dr)   # |r| ~ 4e5 km: one ulp is 6e-11 km
        assert np.array_equal(det["n_steps"], rdet["n_steps"])
    else:
        assert dr < 1e-6 and dv < 1e-9, (dr, dv)
        assert np.abs(det["n_steps"] - rdet["n_steps"]).max() <= 1
    # the lunar field matters on this orbit: without it the run ends somewhere else
    bare = nb.Propagator.default(nb.SpacecraftDynamics.new(nb.OrbitalDynamics.point_masses([MOON, SUN])), mode=mode)
    assert max_dr_dv(bare.engine(nb.EARTH_J2000, alm).propagate_batch(st, cs, ep, end)[0], ref)[0] > 1.0


@pytest.mark.parametrize("mode,kernel", [(nb.MODE_STRICT, nb.KERNEL_THREAD), (nb.MODE_FAST, nb.KERNEL_THREAD), (nb.MODE_FAST, nb.KERNEL_COOP),
                                         (nb.MODE_FAST, nb.KERNEL_TRANSPOSED)])
def test_integration_frame_translations_match_the_oracle(oracle, mode, kernel):
    """States handed over in the Moon frame, integrated in the Earth frame (force_models.rs:181-212): GPU == oracle, and the
    recorded trajectory is in the integration frame."""
    n = 48
    alm, (st_e, cs, ep) = _cislunar(n, seed=9)
    ep = ep + (np.arange(n, dtype=np.int64) % 4) * 1200 * S          # start epochs differ: the translation is per trajectory
    has_field = kernel != nb.KERNEL_THREAD
    dyn = _fields(12) if has_field else nb.SpacecraftDynamics.new(nb.OrbitalDynamics.point_masses([MOON, SUN]))
    prop = nb.Propagator.default(dyn, mode=mode)
    prop.opts.integration_frame = nb.EARTH_J2000
    moon = alm.bodies[alm.body_index(MOON)]
    # Moon-frame states: the low lunar orbits of _cislunar relative to the Moon (position and velocity of the Moon at the epoch the
    # Earth-frame states were built for, velocity by central differences of the same table) — valid LLOs at any start epoch
    st_m = st_e.copy()
    st_m[:3] -= moon.position(0)[:, None]
    st_m[3:6] -= ((moon.position(S) - moon.position(-S)) / 2.0)[:, None]
    eng = prop.engine(nb.MOON_J2000, alm)
    if has_field:
        eng.set_kernel(kernel)
    end = 6 * 3600 * S
    out, oep, det, status, (t_ep, t_st, t_cnt) = eng.propagate_batch(st_m, cs, ep, end, traj_capacity=600)
    packed, opts_c = prop.lower(nb.MOON_J2000, alm)
    assert opts_c.state_center == alm.body_index(MOON) + 1
    ref, rep, rdet, rstatus, (r_ep, r_st, r_cnt) = oracle.propagate_batch(packed.c, opts_c, st_m, cs, ep, end, traj_capacity=600)
    assert (status == 0).all() and np.array_equal(status, rstatus) and np.array_equal(oep, rep)
    dr, dv = max_dr_dv(out, ref)
    if mode == nb.MODE_STRICT:
        assert ((out == ref).all(axis=0)).mean() >= 0.95 and dr < 1e-8
        assert np.array_equal(t_cnt, r_cnt) and np.array_equal(t_ep[:3], r_ep[:3])
    else:
        assert dr < 1e-6 and dv < 1e-9, (dr, dv)
    # results are Moon-relative (|r| ~ 1 900 km), the recording is in the integration frame (|r| ~ 4e5 km)
    assert np.linalg.norm(out[:3], axis=0).max() < 1.0e4 and np.linalg.norm(t_st[:3, 0, :], axis=0).min() > 3e5
    # an epoch outside the ephemeris is a per-trajectory almanac error, not an abort
    ep_bad = ep.copy(); ep_bad[2] = -400 * DAY
    s_bad = eng.propagate_batch(st_m, cs, ep_bad, end)[3]
    assert (s_bad[2] & 0xFF) == 4 and (np.delete(s_bad, 2) == 0).all()


def test_stm_and_filter_reject_what_they_do_not_model():
    alm, (st, cs, ep) = _cislunar(4)
    prop = nb.Propagator.default_dp78(_fields(8))
    with pytest.raises(nb.PropagationError, match="one harmonic field"):
        prop.engine(nb.EARTH_J2000, alm).propagate_batch_stm(st, cs, ep, 600 * S)
