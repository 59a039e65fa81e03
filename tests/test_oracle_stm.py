"""CPU checks of the oracle's STM path (oracle/nyx_oracle_od.c): the dual-number A-matrix of
`SpacecraftDynamics::dual_eom` (dynamics/spacecraft.rs:312-363) against central finite differences of the plain `eom`,
and the propagated STM against the properties the reference's own tests look at (tests/propagation/stm.rs)."""
import ctypes as C

import numpy as np
import pytest

import nyx_b200 as nb
from nyx_b200 import abi

from .util import S


def _eom(oracle, dyn_c, t_ns, y, cs):
    dy = np.zeros(9)
    rc = oracle.lib().nyx_oracle_eom(C.byref(dyn_c), int(t_ns), 0.0, abi.as_double_p(np.ascontiguousarray(y)), abi.as_double_p(cs), abi.as_double_p(dy))
    assert rc == 0
    return dy


def _fd_jacobian(oracle, dyn_c, t_ns, y, cs, cols=range(6), h_pos=1e-3, h_vel=1e-6):
    J = np.zeros((9, 9))
    for j in cols:
        h = h_pos if j < 3 else (h_vel if j < 6 else 0.05)
        yp, ym = y.copy(), y.copy()
        yp[j] += h
        ym[j] -= h
        J[:, j] = (_eom(oracle, dyn_c, t_ns, yp, cs) - _eom(oracle, dyn_c, t_ns, ym, cs)) / (2 * h)
    return J


Y0 = np.array([-2436.45, -2436.45, 6891.037, 5.088611, -5.088611, 0.0, 1.3, 2.2, 20.0])
CS = np.array([100.0, 0.0, 4.0, 0.0])


def _dyn(kind):
    frame = nb.EARTH_J2000
    alm = None
    if kind == "two_body":
        dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.two_body())
    elif kind == "harmonics":
        gd = nb.GravityFieldData.from_fixture("jgm3_70x70", 12, 12, nb.IAU_EARTH_FRAME)
        dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(nb.GravityField.new(gd)))
    elif kind == "point_masses":
        alm = nb.Almanac.synthetic(frame, 0, 3.0)
        dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.point_masses([nb.MOON, nb.SUN]))
    else:
        alm = nb.Almanac.synthetic(frame, 0, 3.0)
        srp = nb.SolarPressure.new([nb.EARTH_J2000], alm)
        dyn = nb.SpacecraftDynamics.from_model(nb.OrbitalDynamics.two_body(), srp)
    return frame, alm, dyn, dyn.pack(frame, alm)


@pytest.mark.parametrize("kind", ["two_body", "harmonics", "point_masses", "srp"])
def test_dual_eom_matches_eom_and_finite_differences(oracle, kind):
    frame, alm, dyn, packed = _dyn(kind)
    t_ns = 3600 * S
    dx, A = oracle.dual_eom(packed.c, t_ns, Y0, CS)
    dy = _eom(oracle, packed.c, t_ns, Y0, CS)
    # the real parts of the dual computation are the accelerations of eom (same formulas, hyperdual op order)
    assert np.abs(dx[:3] - dy[:3]).max() == 0.0
    assert np.abs(dx[3:6] - dy[3:6]).max() < 1e-15 * np.abs(dy[3:6]).max() * 50
    assert np.array_equal(A[:3, 3:6], np.eye(3)) and not A[6:].any() and not A[:3, :3].any() and not A[3:6, 3:6].any()
    J = _fd_jacobian(oracle, packed.c, t_ns, Y0, CS)
    G, Gfd = A[3:6, :3].copy(), J[3:6, :3]
    if kind == "point_masses":
        # AS CODED (orbital.rs:283-294): r_ij carries identity partials too, so the indirect term r_ij/|r_ij|^3 contributes
        # d/d(r_ij) as if it were the spacecraft position.  Remove that known spurious part before comparing with FD.
        for j in range(packed.c.n_bodies):
            if not (packed.c.point_mass_mask >> j) & 1:
                continue
            body = packed.c.bodies[j]
            pos = np.zeros(3)
            assert oracle.lib().nyx_oracle_body_position(C.byref(body), t_ns, abi.as_double_p(pos)) == 0
            r = np.linalg.norm(pos)
            G -= -body.mu_km3_s2 * (np.eye(3) / r**3 - 3.0 * np.outer(pos, pos) / r**5)
    scale = np.abs(Gfd).max()
    assert np.abs(G - Gfd).max() < 2e-7 * scale, (G, Gfd)
    # the gravity-gradient tensor of a potential field is symmetric and (outside the masses) trace-free
    if kind in ("two_body", "harmonics"):
        assert np.abs(G - G.T).max() < 1e-12 * scale
        assert abs(np.trace(G)) < 1e-12 * scale
    if kind == "srp":
        Jc = _fd_jacobian(oracle, packed.c, t_ns, Y0, CS, cols=[6])
        assert np.abs(A[3:6, 6] - Jc[3:6, 6]).max() < 1e-6 * np.abs(Jc[3:6, 6]).max()
    else:
        assert not A[3:6, 6].any()


def test_srp_without_estimation_has_no_cr_column(oracle):
    frame = nb.EARTH_J2000
    alm = nb.Almanac.synthetic(frame, 0, 3.0)
    srp = nb.SolarPressure.default_no_estimation([nb.EARTH_J2000], alm)
    packed = nb.SpacecraftDynamics.from_model(nb.OrbitalDynamics.two_body(), srp).pack(frame, alm)
    _, A = oracle.dual_eom(packed.c, 0, Y0, CS)
    assert not A[3:6, 6].any() and A[3:6, :3].any()


def _prop_stm(oracle, dyn, frame, alm, opts, method, y, cs, dur_ns, stm_in=None):
    packed = dyn.pack(frame, alm)
    st = np.ascontiguousarray(y.reshape(9, 1))
    out, ep, stm, det, status = oracle.propagate_batch_stm(packed.c, opts.to_c(method), st, cs.reshape(4, 1), np.zeros(1, dtype=np.int64), dur_ns,
                                                           stm_in=stm_in)
    assert status[0] == 0 and ep[0] == dur_ns
    return out[:, 0], stm[:, 0].reshape(9, 9).T, det[0]  # column-major tail -> matrix


@pytest.mark.parametrize("ecc", [1e-5, 0.2])
def test_stm_as_coded_single_and_multi_step(oracle, ecc):
    """Within one step the reference integrates Phi' = Phi_k A(t) (spacecraft.rs:203-214 uses ctx.stm), so after a step
    Phi_{k+1} = Phi_k (I + sum_i h b_i A_i): first-order in h per step, composed on the right."""
    frame = nb.EARTH_J2000.with_mu_km3_s2(398600.4415)
    orbit = nb.Orbit.keplerian(8000.0, ecc, 10.0, 5.0, 25.0, 0.0, 0, frame)
    y = np.concatenate([orbit.to_cartesian_pos_vel(), [1.8, 2.2, 0.0]])
    cs = np.array([100.0, 0.0, 1.0, 0.0])
    dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.two_body())
    opts = nb.IntegratorOptions.with_fixed_step_s(10.0)
    y1, phi1, _ = _prop_stm(oracle, dyn, frame, None, opts, nb.IntegratorMethod.RungeKutta4, y, cs, 10 * S)
    # one RK4 step from identity: I + h (A1 + 2A2 + 2A3 + A4)/6 with the A_i at the stage states: check against the trapezoid bound
    packed = dyn.pack(frame, None)
    _, A0 = oracle.dual_eom(packed.c, 0, y, cs)
    _, A1 = oracle.dual_eom(packed.c, 10 * S, y1, cs)
    approx = np.eye(9) + 10.0 * 0.5 * (A0 + A1)
    assert np.abs(phi1 - approx).max() < 1e-7
    assert np.array_equal(phi1[:3, 3:6], 10.0 * np.eye(3)) or np.abs(phi1[:3, 3:6] - 10.0 * np.eye(3)).max() < 1e-14
    # ten steps == ten single steps chained through stm_in (the composition the filter relies on between resets)
    y10, phi10, det = _prop_stm(oracle, dyn, frame, None, opts, nb.IntegratorMethod.RungeKutta4, y, cs, 100 * S)
    assert det["n_steps"] == 10
    yk, phik = y.copy(), np.eye(9)
    for _ in range(10):
        yk_next, phi_step, _ = _prop_stm(oracle, dyn, frame, None, opts, nb.IntegratorMethod.RungeKutta4, yk, cs, 10 * S)
        phik = phik @ phi_step  # Phi_k (I + B_k)
        yk = yk_next
    assert np.abs(yk - y10).max() == 0.0
    assert np.abs(phik - phi10).max() < 1e-12
    # reference test tests/propagation/stm.rs:33-118: |Phi x0 - x(t)| within 1 km of the finite-difference STM's error
    nominal = y10[:6]
    fd = np.zeros((6, 6))
    for i in range(6):
        yp = y.copy()
        yp[i] += 1e-4
        fd[:, i] = (_prop_stm(oracle, dyn, frame, None, opts, nb.IntegratorMethod.RungeKutta4, yp, cs, 100 * S)[0][:6] - nominal) / 1e-4
    delta = (phi10[:6, :6] @ y[:6] - nominal) - (fd @ y[:6] - nominal)
    for i in range(6):  # stm.rs:111-116: "less than 1 km of difference OR the hyperdual computation is better"
        if abs(delta[i]) > 1.0:
            assert delta[i] < 0.0, delta
    # and the as-coded STM is the true one to first order in the span (100 s): the position rows agree to ~|A| t^2
    assert np.abs(phi10[:6, :6] - fd).max() < 0.1  # (100 s)^2 |A| ~ 1e-2 per unit of the dr/dv block (~100)


def test_stm_adaptive_step_matches_plain_propagation(oracle):
    """With the Cartesian error controls the STM block does not influence the step control: the state part of the
    90-vector run equals the plain 9-vector run step for step (same accepted steps, same final state)."""
    frame = nb.EARTH_J2000
    gd = nb.GravityFieldData.from_fixture("jgm3_70x70", 6, 6, nb.IAU_EARTH_FRAME)
    dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(nb.GravityField.new(gd)))
    prop = nb.Propagator.default_dp78(dyn)
    packed = dyn.pack(frame, None)
    st = np.ascontiguousarray(Y0.reshape(9, 1))
    cs = CS.reshape(4, 1)
    ep = np.zeros(1, dtype=np.int64)
    ref, _, rdet, rst = oracle.propagate_batch(packed.c, prop.opts.to_c(prop.method), st, cs, ep, 3600 * S)
    out, _, stm, det, status = oracle.propagate_batch_stm(packed.c, prop.opts.to_c(prop.method), st, cs, ep, 3600 * S)
    assert status[0] == 0 and rst[0] == 0
    assert det["n_steps"][0] == rdet["n_steps"][0]
    assert np.abs(out[:6] - ref[:6]).max() < 1e-9  # real parts of the dual RHS differ from eom by rounding only
    assert np.isfinite(stm).all() and abs(np.linalg.det(stm[:, 0].reshape(9, 9).T) - 1.0) < 5e-2  # first-order-per-step STM: not exactly symplectic


def test_stm_rejects_drag_and_non_cartesian_controls(oracle):
    frame = nb.EARTH_J2000
    dyn = nb.SpacecraftDynamics.from_model(nb.OrbitalDynamics.two_body(), nb.Drag(nb.AtmDensity.Constant(1e-12), nb.IAU_EARTH_FRAME))
    packed = dyn.pack(frame, None)
    opts = nb.IntegratorOptions.default()
    st = np.ascontiguousarray(Y0.reshape(9, 1))
    with pytest.raises(RuntimeError):
        oracle.propagate_batch_stm(packed.c, opts.to_c(nb.IntegratorMethod.RungeKutta89), st, CS.reshape(4, 1), np.zeros(1, dtype=np.int64), 60 * S)
    dyn2 = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.two_body())
    o2 = nb.IntegratorOptions.with_adaptive_step_s(0.1, 30.0, 1e-12, nb.ErrorControl.RSSState)
    with pytest.raises(RuntimeError):
        oracle.propagate_batch_stm(dyn2.pack(frame, None).c, o2.to_c(nb.IntegratorMethod.RungeKutta89), st, CS.reshape(4, 1), np.zeros(1, dtype=np.int64), 60 * S)


def test_reference_two_body_dual_golden(oracle):
    """The reference's own `two_body_dual` (tests/mission_design/orbitaldyn.rs:671-775): f(x) and the 9x9 gradient of the
    two-body `dual_eom` at a GTO-like state against the values the reference asserts to a norm of 1e-16 (mu = the almanac's
    Earth GM, 398600.435436096 — also `examples/04_lro_od/README.md:122`), then its STM consistency check: RK89 fixed 10 s
    over 2 min, Phi(k,0) Phi(k-1,0)^-1 maps the state at k-1 onto the state at k to better than 0.1 km / 0.1 km/s."""
    frame = nb.EARTH_J2000.with_mu_km3_s2(398600.435436096)
    dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.two_body())
    packed = dyn.pack(frame, None)
    y = np.array([-9_042.862_233_600_335, 18_536.333_069_123_244, 6_999.957_069_486_411_5,
                  -3.288_789_003_770_57, -2.226_285_193_102_822, 1.646_738_380_722_676_5, 1.8, 2.2, 0.0])
    cs = np.array([100.0, 0.0, 0.0, 0.0])
    dx, grad = oracle.dual_eom(packed.c, 0, y, cs)
    expected_fx = np.array([-3.288_789_003_770_57, -2.226_285_193_102_822, 1.646_738_380_722_676_5,
                            0.000_348_875_166_711_715_13, -0.000_715_134_890_110_951_6, -0.000_270_059_537_180_366_4])
    assert np.linalg.norm(dx[:6] - expected_fx) < 1e-16
    expected = np.zeros((9, 9))
    expected[0, 3] = expected[1, 4] = expected[2, 5] = 1.0
    expected[3, 0] = -0.000_000_018_628_398_391_083_86
    expected[4, 0] = expected[3, 1] = -0.000_000_040_897_747_124_379_53
    expected[5, 0] = expected[3, 2] = -0.000_000_015_444_396_313_003_294
    expected[4, 1] = 0.000_000_045_253_271_058_430_05
    expected[5, 1] = expected[4, 2] = 0.000_000_031_658_391_636_846_51
    expected[5, 2] = -0.000_000_026_624_872_667_346_21
    assert np.linalg.norm(grad - expected) < 1e-16
    # STM consistency over the last step (orbitaldyn.rs:745-774)
    prop = nb.Propagator.rk89(dyn, nb.IntegratorOptions.with_fixed_step(10 * S))
    st, css, ep = y.reshape(9, 1), cs.reshape(4, 1), np.zeros(1, dtype=np.int64)
    opts = prop.opts.to_c(prop.method)
    fin, _, stm_k, _, s1 = oracle.propagate_batch_stm(packed.c, opts, st, css, ep, 120 * S)
    prev, _, stm_km1, _, s2 = oracle.propagate_batch_stm(packed.c, opts, st, css, ep, 110 * S)
    assert s1[0] == 0 and s2[0] == 0
    phi_k = stm_k[:, 0].reshape(9, 9).T      # [(c*9 + r)] -> (r, c)
    phi_km1 = stm_km1[:, 0].reshape(9, 9).T
    err = phi_k @ np.linalg.inv(phi_km1) @ prev[:, 0] - fin[:, 0]
    assert np.linalg.norm(err[:3]) < 1e-1 and np.linalg.norm(err[3:6]) < 1e-1
