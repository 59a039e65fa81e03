"""`MvnSpacecraft::new` with dispersions of orbital elements (mc/multivariate.rs:80-211) — the reference's own unit tests
(`multivariate_ut`, multivariate.rs:345-714) restated as statistics: its assertions on exact counts belong to the Pcg64Mcg +
ziggurat stream, which is not reproduced here, so the same quantities are checked against their expectations."""
import numpy as np
import pytest

import nyx_b200 as nb
from nyx_b200.param import StateParameter as P, evaluate

FRAME = nb.EARTH_J2000.with_mu_km3_s2(nb.GMAT_EARTH_GM)
CHI2_95 = {1: 3.841458820694124, 6: 12.591587243743977}


def _state(sma=8191.93, raan=306.614, aop=314.19, ta=99.8877):
    return nb.Spacecraft.from_orbit(nb.Orbit.keplerian(sma, 1e-6, 12.85, raan, aop, ta, 0, FRAME))


def test_mvn_generator_mahalanobis_chi_squared():
    """test_mvn_generator (multivariate.rs:370-420): random PSD 6x6 covariance, 95th percentile of the squared Mahalanobis
    distance of 1000 samples within 20 % of chi-squared(6)."""
    rng = np.random.default_rng(1)
    a = rng.random((6, 6))
    cov = np.zeros((9, 9))
    cov[:6, :6] = a @ a.T
    sc = _state()
    mvn = nb.MvnSpacecraft.from_spacecraft_cov(sc, cov, np.zeros(9))
    x = mvn.sample_vectors(rng, 1000)
    cov_inv = np.linalg.pinv(cov, rcond=1e-12)
    md = np.sort(np.einsum("ni,ij,nj->n", x, cov_inv, x))
    assert abs(md[950] - CHI2_95[6]) / CHI2_95[6] < 0.2
    # the generated states are the template plus the draw
    ds = mvn.apply(x[0])
    assert np.allclose(ds.state.to_vector() - sc.to_vector(), x[0], rtol=0, atol=1e-12)


def test_disperse_r_mag():
    """disperse_r_mag (multivariate.rs:422-476): a 1 km (1 sigma) dispersion of |r|: about 0.3 % of 1000 samples beyond 3 km."""
    sc = _state()
    gen = nb.MvnSpacecraft.new(sc, [nb.StateDispersion(P.Rmag, std_dev=1.0)])
    rng = np.random.default_rng(0)
    r0 = sc.orbit.rmag_km()
    dev = np.array([gen.apply(x).state.orbit.rmag_km() - r0 for x in gen.sample_vectors(rng, 1000)])
    assert (np.abs(dev) >= 3.0).sum() <= 12 and 0.9 < dev.std() < 1.1
    assert gen.apply(gen.sample_vectors(rng, 1)[0]).actual_dispersions[0][0] == "Rmag"


def test_disperse_full_cartesian():
    """disperse_full_cartesian (multivariate.rs:478-566): six Cartesian dispersions; per component ~31.7 % beyond 1 sigma."""
    sc = _state()
    std = [10.0, 10.0, 10.0, 0.2, 0.2, 0.2]
    gen = nb.MvnSpacecraft.new(sc, [nb.StateDispersion(p, std_dev=s) for p, s in zip((P.X, P.Y, P.Z, P.VX, P.VY, P.VZ), std)])
    assert np.allclose(gen.sqrt_s_v @ gen.sqrt_s_v.T, np.diag(np.array(std + [0, 0, 0]) ** 2), atol=1e-9)
    x = gen.sample_vectors(np.random.default_rng(0), 1000)
    beyond = (np.abs(x[:, :6]) > np.array(std)[None, :]).sum()
    assert abs(beyond / 6 - 317) < 40          # the reference's stream gives 312
    ds = gen.apply(x[0])
    assert [n for n, _ in ds.actual_dispersions] == ["X", "Y", "Z", "VX", "VY", "VZ"]
    assert np.allclose([v for _, v in ds.actual_dispersions], -x[0, :6], atol=1e-9)


def test_disperse_raan_only():
    """disperse_raan_only (multivariate.rs:568-633): a 0.2 deg RAAN dispersion leaves SMA and inclination within 5 %, and the
    realised RAAN dispersions pass the chi-squared(1) percentile test."""
    sc = _state(8100.0, 356.614, 14.19, 199.8877)
    gen = nb.MvnSpacecraft.new(sc, [nb.StateDispersion.zero_mean(P.RAAN, 0.2)])
    mu = FRAME.mu_km3_s2()
    rv0 = sc.orbit.to_cartesian_pos_vel().reshape(6, 1)
    md = []
    for x in gen.sample_vectors(np.random.default_rng(0), 1000):
        ds = gen.apply(x)
        rv = ds.state.orbit.to_cartesian_pos_vel().reshape(6, 1)
        for prm in (P.SemiMajorAxis, P.Inclination):
            orig, new = float(evaluate(prm, rv0, mu)[0]), float(evaluate(prm, rv, mu)[0])
            assert 100.0 * abs(orig - new) / orig < 5.0
        md.append((ds.actual_dispersions[0][1] / 0.2) ** 2)
    assert abs(np.sort(md)[950] - CHI2_95[1]) / CHI2_95[1] < 0.2


def test_disperse_keplerian():
    """disperse_keplerian (multivariate.rs:635-714): SMA / inclination / RAAN / AoP dispersions; the Cartesian sample mean stays
    within 1 (km, km/s norm) of the nominal and the sample covariance within 20 % of sqrt_s_v sqrt_s_v^T."""
    sc = _state(8100.0, 356.614, 14.19, 199.8877)
    gen = nb.MvnSpacecraft.new(sc, [nb.StateDispersion.zero_mean(P.SemiMajorAxis, 10.0), nb.StateDispersion.zero_mean(P.Inclination, 0.15),
                                    nb.StateDispersion.zero_mean(P.RAAN, 0.02), nb.StateDispersion.zero_mean(P.AoP, 0.02)])
    expected = (gen.sqrt_s_v @ gen.sqrt_s_v.T)[:6, :6]
    x = gen.sample_vectors(np.random.default_rng(0), 2000)[:, :6]
    assert np.linalg.norm(x.mean(axis=0)) < 1.0
    sample_cov = np.cov(x.T, ddof=1)
    assert np.linalg.norm(sample_cov - expected) / np.linalg.norm(expected) < 0.2
    # the well-conditioned elements come back with the requested spread (AoP is ill-defined at e = 1e-6)
    got = np.array([[v for _, v in gen.apply(xx).actual_dispersions] for xx in gen.sample_vectors(np.random.default_rng(1), 500)])
    assert abs(got[:, 0].std() - 10.0) < 1.0 and abs(got[:, 1].std() - 0.15) < 0.015 and abs(got[:, 2].std() - 0.02) < 0.002


def test_non_orbital_dispersions_and_monte_carlo_wiring():
    sc = _state()
    gen = nb.MvnSpacecraft.new(sc, [nb.StateDispersion(P.Cr, std_dev=0.1), nb.StateDispersion(P.PropMass, mean=2.0, std_dev=0.5),
                                    nb.StateDispersion(P.X, std_dev=1.0)])
    c = gen.sqrt_s_v @ gen.sqrt_s_v.T
    assert abs(c[6, 6] - 0.01) < 1e-12 and abs(c[8, 8] - 0.25) < 1e-12 and abs(c[0, 0] - 1.0) < 1e-9 and gen.mean[8] == 2.0
    with pytest.raises(nb.StateError):
        nb.MvnSpacecraft.new(sc, [nb.StateDispersion(P.TotalMass, std_dev=1.0)])
    zm = nb.MvnSpacecraft.zero_mean(sc, [nb.StateDispersion(P.PropMass, mean=2.0, std_dev=0.5)])
    assert zm.mean[8] == 0.0
    mc = nb.MonteCarlo(sc, gen, "disp", seed=3)
    states = mc.generate_states(0, 5)
    assert [n for n, _ in states[0][1].actual_dispersions] == ["Cr", "PropMass", "X"]
