"""Unit checks of the oracle's building blocks (CPU only): the pieces whose semantics the parity claims rest on."""
import ctypes as C
import math

import numpy as np
import pytest

import nyx_b200 as nb
from nyx_b200 import abi


def test_duration_semantics(oracle):
    """hifitime: f64*Unit::Second truncates toward zero to integer ns; to_seconds = whole s + subsec ns * 1e-9."""
    L = oracle.lib()
    assert L.nyx_oracle_dur_from_seconds(83.3333333339) == 83333333333
    assert L.nyx_oracle_dur_from_seconds(-0.9999999999) == -999999999  # toward zero, not floor
    assert L.nyx_oracle_dur_from_seconds(float("nan")) == 0            # Rust `NaN as i128`
    for ns in (0, 1, 999_999_999, 60_000_000_000, 83_333_333_333, 2_700_000_000_000, 86_400_000_000_000 * 30):
        sec, sub = divmod(ns, 10**9)
        assert L.nyx_oracle_dur_to_seconds(ns) == float(sec) + float(sub) * 1e-9
        assert nb.duration_to_seconds(ns) == L.nyx_oracle_dur_to_seconds(ns)
    assert L.nyx_oracle_dur_to_seconds(-10 * 10**9) == -10.0
    assert L.nyx_oracle_dur_to_seconds(-500_000_000) == -0.5
    assert 0.001 * nb.Unit.Second == 1_000_000 and 2 * nb.Unit.Day == 172_800 * 10**9


@pytest.mark.parametrize("method", range(6))
def test_tableaux_are_consistent(oracle, method):
    """Row sums give the c_i (instance.rs:379-386), b and b* sum to 1; RK89 has the c = 4/3 stage."""
    L = oracle.lib()
    order, stages = C.c_int(), C.c_int()
    a, b = abi.c_double_p(), abi.c_double_p()
    assert L.nyx_oracle_tableau(method, C.byref(order), C.byref(stages), C.byref(a), C.byref(b)) == 0
    S_ = stages.value
    A = np.ctypeslib.as_array(a, shape=(S_ * (S_ - 1) // 2,))
    B = np.ctypeslib.as_array(b, shape=(2 * S_,))
    assert abs(B[:S_].sum() - 1.0) < 1e-14 and abs(B[S_:].sum() - 1.0) < 1e-14
    cs, idx = [], 0
    for i in range(S_ - 1):
        cs.append(A[idx: idx + i + 1].sum())
        idx += i + 1
    assert all(-1e-12 <= c <= 4.0 / 3.0 + 1e-12 for c in cs)
    assert (order.value, S_) == {0: (9, 16), 1: (8, 13), 2: (5, 7), 3: (4, 4), 4: (5, 6), 5: (6, 8)}[method]
    if method == abi.RK89:
        assert abs(cs[10] - 4.0 / 3.0) < 1e-12 and abs(cs[14] - 1.0) < 1e-13  # large cancelling a_ij: 3.8e-14 of round-off


def test_deterministic_sincos_accuracy(oracle):
    L = oracle.lib()
    s, c = C.c_double(), C.c_double()
    rng = np.random.default_rng(0)
    worst = 0.0
    for x in np.concatenate([rng.uniform(-7, 7, 4000), rng.uniform(-1e-3, 1e-3, 200), [0.0, math.pi / 2, math.pi, 6.283185307179586]]):
        L.nyx_oracle_sincos(float(x), C.byref(s), C.byref(c))
        for got, ref in ((s.value, math.sin(x)), (c.value, math.cos(x))):
            worst = max(worst, abs(got - ref) / max(np.spacing(abs(ref)), 1e-300) if abs(ref) > 1e-3 else abs(got - ref) / 2.2e-19)
    assert worst <= 2.0, worst  # <= 2 ulp


def test_rotation_is_orthonormal_and_matches_iau_model(oracle):
    L = oracle.lib()
    r = nb.frames.IAU_EARTH_ROTATION
    rot = abi.Rotation(1, 0, r.ra0_deg, r.ra1_deg_cy, r.dec0_deg, r.dec1_deg_cy, r.w0_deg, r.w1_deg_day)
    R = np.zeros(9)
    wdot = C.c_double()
    for t_ns in (0, 3600 * 10**9, 86400 * 10**9 * 3):
        L.nyx_oracle_rotation(C.byref(rot), t_ns, abi.as_double_p(R), C.byref(wdot))
        M = R.reshape(3, 3)
        assert np.abs(M @ M.T - np.eye(3)).max() < 1e-15
        d = t_ns / 1e9 / 86400.0
        W = math.radians((r.w0_deg + r.w1_deg_day * d) % 360.0)
        assert abs(math.atan2(M[0, 1], M[0, 0]) - (W + math.radians(90.0) - 2 * math.pi * round((W + math.radians(90.0)) / (2 * math.pi)))) < 1e-5
        assert M[2, 2] > 0.999999  # pole within ~0.6"/yr of +z near J2000
    assert abs(wdot.value - math.radians(r.w1_deg_day) / 86400.0) < 1e-18


def test_error_controls_against_definitions(oracle):
    """error_ctrl.rs:79-230 on a hand-computable case."""
    L = oracle.lib()
    err = np.array([3e-9, 4e-9, 0, 0, 0, 1e-12, 0, 0, 0])
    cur = np.array([7000.0, 0, 0, 0, 7.5, 0, 1.8, 2.2, 5.0])
    cand = cur + np.array([3.0, 4.0, 0, 0, 0, 0.01, 0, 0, 0])
    f = lambda k: L.nyx_oracle_error_estimate(k, abi.as_double_p(err), abi.as_double_p(cand), abi.as_double_p(cur))
    assert f(abi.RSS_CARTESIAN_STEP) == pytest.approx(max(5e-9 / 5.0, 1e-12), rel=1e-15)  # |dv| = 0.01 < sqrt(0.1): absolute
    assert f(abi.RSS_CARTESIAN_STATE) == pytest.approx(5e-9 / (0.5 * np.linalg.norm((cand + cur)[:3])), rel=1e-14)
    assert f(abi.LARGEST_ERROR) == pytest.approx(max(3e-9 / 3.0, 4e-9 / 4.0, 1e-12), rel=1e-15)
    assert f(abi.LARGEST_STEP) == pytest.approx(err.sum() / np.abs(cand - cur).sum(), rel=1e-14)
    assert f(abi.RSS_STEP) == pytest.approx(np.linalg.norm(err) / np.linalg.norm(cand - cur), rel=1e-14)


def test_occultation_geometry(oracle):
    """anise's apparent-disk model: full sun, umbra, penumbra monotone in between (cosmic/eclipse.rs:69-83)."""
    L = oracle.lib()
    sun = np.array([1.496e8, 0.0, 0.0])
    occ = lambda r: L.nyx_oracle_occultation(abi.as_double_p(np.asarray(r, float)), abi.as_double_p(sun - np.asarray(r, float)), 696000.0, 6378.14)
    assert occ([7000.0, 0, 0]) == 0.0            # sub-solar side
    assert occ([-7000.0, 0, 0]) == 1.0           # behind the Earth
    assert occ([0.0, 42000.0, 0.0]) == 0.0       # GEO, quadrature
    vals = [occ([-42000.0, y, 0.0]) for y in np.linspace(6100.0, 6700.0, 13)]
    assert vals[0] == 1.0 and vals[-1] == 0.0 and all(a >= b for a, b in zip(vals, vals[1:])) and 0 < vals[6] < 1


def test_ephemeris_tables_match_series_and_oracle_clenshaw(oracle):
    from nyx_b200 import ephem

    alm = nb.Almanac.synthetic(nb.EARTH_J2000, 0, 20.0)
    L = oracle.lib()
    pos = np.zeros(3)
    for b in alm.bodies:
        co = np.ascontiguousarray(b.coeffs)
        bc = abi.BodyC(b.frame.mu_km3_s2(), b.frame.radius_km, b.t0_ns, b.interval_ns, b.n_intervals, b.n_coeffs, abi.as_double_p(co))
        for t_ns in (0, 123456789012345, 15 * 86400 * 10**9):
            assert L.nyx_oracle_body_position(C.byref(bc), t_ns, abi.as_double_p(pos)) == 0
            assert np.array_equal(pos, b.position(t_ns))  # same Clenshaw recurrence, bit for bit
            truth = ephem.position(b.frame.ephemeris_id, nb.EARTH, t_ns / 1e9)
            assert np.linalg.norm(pos - truth) / np.linalg.norm(truth) < 1e-9
        assert L.nyx_oracle_body_position(C.byref(bc), b.t0_ns - 1, abi.as_double_p(pos)) != 0  # outside coverage
    moon = alm.bodies[alm.body_index(nb.MOON)].position(0)
    assert 3.5e5 < np.linalg.norm(moon) < 4.1e5
