"""On-device dispersions: the counter-based stream of nyx_b200/csrc/nyxb_mvn.cu and its CPU restatement
(oracle/nyx_oracle_mvn.c).  CPU part: Philox known-answer vectors + sample moments; GPU part: device == oracle."""
import ctypes as C

import numpy as np
import pytest

import nyx_b200 as nb

# Random123 known-answer vectors for philox4x32-10 (kat_vectors): counter, key -> output
KAT = [
    ((0x00000000,) * 4, (0x00000000,) * 2, (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers(oracle):
    L = oracle.lib()
    for ctr, key, want in KAT:
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        L.nyx_oracle_philox4x32_10(c, k, o)
        assert tuple(o) == want, [hex(v) for v in o]


def _mvn():
    orbit = nb.Orbit.keplerian(7000.0, 0.01, 51.6, 30.0, 40.0, 10.0, 0, nb.EARTH_J2000)
    tmpl = nb.Spacecraft(orbit=orbit, mass=nb.Mass(500.0, 50.0, 0.0), srp=nb.SRPData(2.0, 1.2))
    rng = np.random.default_rng(0)
    A = rng.normal(size=(9, 9)) * np.array([1, 1, 1, 1e-3, 1e-3, 1e-3, 0.05, 0.1, 0.5])[:, None]
    cov = A @ A.T
    return tmpl, nb.MvnSpacecraft.from_spacecraft_cov(tmpl, cov, mean=np.array([0.1, 0, 0, 0, 0, 0, 0, 0, 0.0])), cov


def test_oracle_stream_has_the_requested_moments(oracle):
    tmpl, mvn, cov = _mvn()
    n = 200_000
    st, disp = oracle.mvn_sample(7, 0, n, tmpl.to_vector(), mvn.mean, mvn.sqrt_s_v)
    assert np.allclose(st - tmpl.to_vector()[:, None], disp)
    d = disp - mvn.mean[:, None]
    sig = np.sqrt(np.diag(cov))
    assert (np.abs(d.mean(axis=1)) < 5 * sig / np.sqrt(n)).all()
    emp = d @ d.T / n
    assert np.abs(emp - cov).max() < 6 * (sig[:, None] * sig[None, :]).max() / np.sqrt(n) * 3
    corr_err = np.abs(emp / np.outer(sig, sig) - cov / np.outer(sig, sig)).max()
    assert corr_err < 0.02
    # a run's draw depends on (seed, run index) only: shards reproduce the whole
    a, _ = oracle.mvn_sample(7, 1000, 64, tmpl.to_vector(), mvn.mean, mvn.sqrt_s_v)
    assert np.array_equal(a, st[:, 1000:1064])
    b, _ = oracle.mvn_sample(8, 0, 64, tmpl.to_vector(), mvn.mean, mvn.sqrt_s_v)
    assert not np.array_equal(b, st[:, :64])


@pytest.mark.gpu
def test_device_dispersions_match_oracle_and_shard_invariance(oracle):
    tmpl, mvn, cov = _mvn()
    n = 4097
    st, disp = mvn.sample_on_device(7, n)
    ref, rdisp = oracle.mvn_sample(7, 0, n, tmpl.to_vector(), mvn.mean, mvn.sqrt_s_v)
    # same Philox integers; log / sincos come from different libms: a few ulp of the O(1) normals scaled by |L|
    scale = np.abs(mvn.sqrt_s_v).sum(axis=1)[:, None]
    assert (np.abs(disp - rdisp) <= 1e-14 * scale * 8).all()
    assert (np.abs(st - ref) <= 1e-14 * scale * 8 + 1e-12).all()
    a, _ = mvn.sample_on_device(7, 1000, first_index=0)
    b, _ = mvn.sample_on_device(7, 3097, first_index=1000)
    assert np.array_equal(np.concatenate([a, b], axis=1), st)


@pytest.mark.gpu
def test_monte_carlo_with_device_dispersions(oracle):
    tmpl, _, _ = _mvn()
    mvn = nb.MvnSpacecraft.from_cartesian_std(tmpl, 1.0, 1e-3)
    mc = nb.MonteCarlo(tmpl, mvn, "device dispersions", seed=3)
    prop = nb.Propagator.default(nb.SpacecraftDynamics.new(nb.OrbitalDynamics.two_body()), mode=nb.MODE_STRICT)
    res = mc.run_until_epoch(prop, None, 1800 * 10**9, 256, device_dispersions=True)
    assert (res.status == 0).all() and len(res.ok_runs()) == 256
    st = np.stack([r.dispersed_state.state.to_vector() for r in res.runs], axis=1)
    cs = np.empty((4, 256)); cs[0], cs[1], cs[2], cs[3] = 500.0, 0.0, 2.0, 0.0
    packed = prop.dynamics.pack(nb.EARTH_J2000, None)
    ref, _, _, rstatus = oracle.propagate_batch(packed.c, prop.opts.to_c(prop.method), st, cs, np.zeros(256, dtype=np.int64), 1800 * 10**9)
    assert np.array_equal(res.final_state_soa, ref)   # STRICT two-body: bit-identical given identical inputs
    d = st[:3] - tmpl.to_vector()[:3, None]
    assert 0.8 < d.std() < 1.2
