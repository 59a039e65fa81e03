"""Independent arbiters for arithmetic the reference cannot pin here (anise / data files absent): the oracle's spherical-harmonic
acceleration and its dual-number gradient against the textbook potential evaluated at 40 digits (mpmath, closed-form Legendre
functions, tests/arbiters.py), the host Brent search and Hermite interpolation against scipy, the eclipse fraction against a
brute-force area quadrature.  The 200-point / 70x70 study is scripts/arbiter_harmonics.py -> profiles/r02_arbiter_harmonics.json."""
import ctypes as C
import multiprocessing as mp

import numpy as np
import pytest

import nyx_b200 as nb
from nyx_b200 import abi


def _field(fixture, deg):
    moon = fixture.startswith("luna")
    gd = nb.GravityFieldData.from_fixture(fixture, deg, deg, nb.IAU_MOON_FRAME if moon else nb.IAU_EARTH_FRAME)
    dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(nb.GravityField.new(gd)))
    packed = dyn.pack(nb.MOON_J2000 if moon else nb.EARTH_J2000, None)
    packed.c.gravity[0].rot.kind = 0
    return gd, packed


def _harm_case(job):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from scripts.arbiter_harmonics import oracle_field_accel
    from tests.arbiters import mp_harmonic_accel

    fixture, deg, seed = job
    gd, packed = _field(fixture, deg)
    gf = packed.c.gravity[0]
    rng = np.random.default_rng(seed)
    d = rng.normal(size=3)
    rb = d / np.linalg.norm(d) * gf.r_eq_km * rng.uniform(1.02, 1.5)
    got = oracle_field_accel(packed.c, rb)
    want = np.array(mp_harmonic_accel(gd.c_nm, gd.s_nm, deg, deg, gf.mu_km3_s2, gf.r_eq_km, rb))
    return float(np.abs(got - want).max() / np.abs(want).max())


def test_harmonic_acceleration_against_the_textbook_potential_at_40_digits(oracle):
    """gravity_field.rs:148-268 as restated by the oracle vs the spherical-coordinate gradient of the normalised potential with
    closed-form associated Legendre functions: <= 1e-13 relative (measured 8e-16 for 21x21, profiles/r02_arbiter_harmonics.json)."""
    jobs = [("jgm3_70x70", 21, 100 + s) for s in range(4)] + [("luna_jggrx_80x80", 12, 200 + s) for s in range(4)]
    with mp.get_context("fork").Pool(min(8, mp.cpu_count())) as pool:
        errs = pool.map(_harm_case, jobs, chunksize=1)
    assert max(errs) < 1e-13, errs


def _grad_case(seed):
    from tests.arbiters import mp_harmonic_accel
    import mpmath

    gd, packed = _field("jgm3_70x70", 8)
    gf = packed.c.gravity[0]
    rng = np.random.default_rng(seed)
    d = rng.normal(size=3)
    rb = d / np.linalg.norm(d) * gf.r_eq_km * rng.uniform(1.05, 1.4)
    # central differences of the 40-digit acceleration: step 1e-6 km on |r| ~ 7e3 km, truncation error O(h^2 a''') ~ 1e-24
    h = 1e-6
    G = np.zeros((3, 3))
    for k in range(3):
        e = np.zeros(3); e[k] = h
        ap = np.array([mpmath.mpf(v) for v in mp_harmonic_accel(gd.c_nm, gd.s_nm, 8, 8, gf.mu_km3_s2, gf.r_eq_km, rb + e, dps=40)], dtype=object)
        am = np.array([mpmath.mpf(v) for v in mp_harmonic_accel(gd.c_nm, gd.s_nm, 8, 8, gf.mu_km3_s2, gf.r_eq_km, rb - e, dps=40)], dtype=object)
        G[:, k] = [float((ap[i] - am[i]) / (2 * h)) for i in range(3)]
    return rb, G


def test_dual_number_gradient_of_the_harmonics_against_finite_differences_of_the_arbiter(oracle):
    """gravity_field.rs:273-431 (`dual_eom`, hyperdual operator rules restated from their published formulas) against differences
    of the independent 40-digit acceleration; the two-body part is added analytically.  The arbiter returns f64-rounded values, so
    the difference quotient carries 1e-16 |a| / h ~ 1e-15 km/s^2 per km of noise against gradients of ~1e-9: 1e-6 relative."""
    with mp.get_context("fork").Pool(2) as pool:
        cases = pool.map(_grad_case, [11, 12])
    gd, packed = _field("jgm3_70x70", 8)
    mu = packed.c.mu_central_km3_s2
    for rb, G_h in cases:
        y = np.concatenate([rb, [1.0, 2.0, 3.0], [1.8, 2.2, 0.0]])
        _, A = oracle.dual_eom(packed.c, 0, y, np.array([100.0, 0.0, 1.0, 1.0]))
        r = np.linalg.norm(rb)
        G_tb = -mu / r ** 3 * (np.eye(3) - 3.0 * np.outer(rb, rb) / r ** 2)
        got_h = A[3:6, 0:3] - G_tb
        assert np.abs(got_h - G_h).max() < 2e-6 * np.abs(G_h).max(), (got_h, G_h)
        assert np.abs(A[3:6, 0:3] - (G_tb + G_h)).max() < 1e-8 * np.abs(G_tb).max()   # the quotient's own noise is ~1e-15


def test_brent_and_hermite_against_scipy():
    from scipy.interpolate import KroghInterpolator
    from scipy.optimize import brentq

    from nyx_b200.event import brent
    from nyx_b200.trajectory import hermite_eval

    f = lambda t: np.cos(t) - 0.3 * t + 0.1 * np.sin(5 * t)
    want = brentq(f, 0.2, 2.0, xtol=1e-14, rtol=1e-15)
    got = brent(f, 0.2, 2.0, 1e-13)
    got = got[0] if isinstance(got, tuple) else got
    assert abs(got - want) < 1e-12
    # Hermite interpolation of a Keplerian-like arc through 13 nodes with derivatives == Krogh's divided-difference interpolator
    xs = np.linspace(0.0, 780.0, 13)
    w = 1.1e-3
    ys, yd = 7000.0 * np.cos(w * xs), -7000.0 * w * np.sin(w * xs)
    k = KroghInterpolator(np.repeat(xs, 2), np.column_stack([ys, yd]).ravel())
    for x in (1.0, 333.3, 401.7, 779.0):
        y, dy = hermite_eval(xs, ys, yd, x)
        assert abs(y - float(k(x))) < 1e-9 and abs(dy - float(k.derivative(x))) < 1e-11
        assert abs(y - 7000.0 * np.cos(w * x)) < 1e-9   # and both equal the function: 26 conditions on a smooth arc


@pytest.mark.parametrize("r_ls,r_body,d", [(4.6e-3, 9.2e-3, 8.0e-3), (4.6e-3, 2.0e-3, 1.0e-3), (4.6e-3, 4.7e-3, 6.0e-3), (4.6e-3, 9.0e-3, 2.0e-2),
                                           (4.6e-3, 9.0e-3, 3.0e-3)])
def test_occultation_against_area_quadrature(oracle, r_ls, r_body, d):
    """`occultation` (restated from anise's published apparent-disk overlap) vs a brute-force quadrature of the covered part of the
    solar disk.  Geometry built so that the apparent radii / separation are (r_ls, r_body, d) radians."""
    from tests.arbiters import sun_visible_fraction

    AU, R_sun = 1.4959787e8, 6.96e5
    dist_sun = R_sun / np.sin(r_ls)
    body_radius = 1737.4
    dist_body = body_radius / np.sin(r_body)
    # observer at the origin; light source along +x; eclipsing body at angle d from it
    r_ls_vec = np.array([dist_sun, 0.0, 0.0])                          # observer -> light source
    r_eb_vec = -np.array([np.cos(d), np.sin(d), 0.0]) * dist_body      # eclipsing body -> observer
    L = oracle.lib()
    occ = L.nyx_oracle_occultation(abi.as_double_p(np.ascontiguousarray(r_eb_vec)), abi.as_double_p(np.ascontiguousarray(r_ls_vec)), R_sun, body_radius)
    want = 1.0 - sun_visible_fraction(r_ls, r_body, d)
    assert abs(occ - want) < 2e-4, (occ, want)
