"""The reference's Monte Carlo sampling stream restated in libnyxb.so's host code (csrc/nyxb_rng.cu): Pcg64Mcg (rand_pcg) pinned on
the PCG family's official known-answer vector, the ziggurat StandardNormal (rand_distr) checked on its published table entries and
statistically.  Reference call sites: mc/montecarlo.rs:277-296, mc/multivariate.rs:298-302."""
import numpy as np

import nyx_b200 as nb
from nyx_b200 import abi


def test_pcg64mcg_known_answer_vector():
    """Mcg128Xsl64::new(42): the six outputs of the generator's reference (C) test suite, as carried by rand_pcg's own test."""
    lib = abi.load_library()
    out = np.zeros(6, dtype=np.uint64)
    assert lib.nyxb_pcg64mcg_u64(42, 0, 6, out.ctypes.data) == 0
    assert [int(v) for v in out] == [0x63B4A3A813CE700A, 0x382954200617AB24, 0xA7FD85AE3FE950CE, 0xD715286AA2887737,
                                     0x60C92FEE2E59F32C, 0x84C4E96BEFF30017]
    # independent restatement in Python integers, 128-bit seeds included
    M, MASK = 0x2360ED051FC65DA44385DF649FCCF645, (1 << 128) - 1
    seed = (0x0123456789ABCDEF << 64) | 0xFEDCBA9876543210
    s, want = (seed | 1) & MASK, []
    for _ in range(32):
        s = (s * M) & MASK
        rot, xsl = s >> 122, ((s >> 64) ^ s) & 0xFFFFFFFFFFFFFFFF
        want.append(((xsl >> rot) | (xsl << ((64 - rot) & 63))) & 0xFFFFFFFFFFFFFFFF)
    out = np.zeros(32, dtype=np.uint64)
    assert lib.nyxb_pcg64mcg_u64(seed & 0xFFFFFFFFFFFFFFFF, seed >> 64, 32, out.ctypes.data) == 0
    assert [int(v) for v in out] == want


def test_ziggurat_tables_match_the_published_entries():
    lib = abi.load_library()
    x, f = np.zeros(257), np.zeros(257)
    assert lib.nyxb_ziggurat_tables(x.ctypes.data, f.ctypes.data) == 0
    # first / last entries of rand_distr's ZIG_NORM_X as printed in its table (18 decimals)
    for got, want in ((x[0], "3.910757959537090045"), (x[1], "3.654152885361008796"), (x[2], "3.449278298560964462"), (x[255], "0.215241895913273806")):
        assert f"{got:.18f}" == want
    assert x[256] == 0.0 and f[256] == 1.0 and np.all(np.diff(x) < 0) and np.allclose(f, np.exp(-x * x / 2))
    # every layer has the same area V (the defining property of the table)
    V = 0.00492867323399
    area = x[1:256] * (f[2:257] - f[1:256])
    assert np.abs(area[:-1] - V).max() < 1e-15 and abs(area[-1] - V) < 1e-11   # the top layer absorbs the truncation of V


def test_reference_normals_moments_tail_and_skip():
    lib = abi.load_library()
    n = 40_000
    z = np.empty((n, 9))
    assert lib.nyxb_reference_normals(7, 0, 0, n, z.ctypes.data) == 0
    v = z.ravel()
    assert abs(v.mean()) < 4 / np.sqrt(v.size) and abs(v.var() - 1) < 0.01 and abs((v ** 3).mean()) < 0.02 and abs((v ** 4).mean() - 3) < 0.05
    tail = (np.abs(v) > 3.6541528853610088).mean()     # the ziggurat's tail branch: P(|z| > R) = 2.58e-4
    assert 1.5e-4 < tail < 3.8e-4
    assert np.abs(np.corrcoef(z.T) - np.eye(9)).max() < 0.02
    # `.skip(k)` drops whole runs of the SAME serial stream
    z2 = np.empty((100, 9))
    assert lib.nyxb_reference_normals(7, 0, 250, 100, z2.ctypes.data) == 0
    assert np.array_equal(z2, z[250:350])


def test_monte_carlo_with_the_reference_stream():
    orbit = nb.Orbit.keplerian(7000.0, 0.01, 30.0, 10.0, 20.0, 0.0, 0, nb.EARTH_J2000)
    template = nb.Spacecraft(orbit=orbit, mass=nb.Mass(100.0, 10.0, 0.0))
    mvn = nb.MvnSpacecraft.from_cartesian_std(template, 1.0, 1e-3)
    mc = nb.MonteCarlo(template, mvn, "ref-stream", seed=(1 << 100) + 12345)
    mc.stream = "reference"
    a = mc.generate_states(0, 50)
    b = mc.generate_states(10, 40)
    assert all(np.array_equal(x[1].state.to_vector(), y[1].state.to_vector()) for x, y in zip(a[10:], b))
    x = np.array([s.state.to_vector() - template.to_vector() for _, s in a])
    assert np.abs(x[:, :3]).max() < 6.0 and np.abs(x[:, 3:6]).max() < 6e-3 and np.all(x[:, 6:] == 0)
    assert not np.array_equal(x, np.array([s.state.to_vector() - template.to_vector() for _, s in mc.generate_states(0, 50, stream="numpy")]))
