"""Randomised STRICT-mode parity sweep: GPU (through the C ABI) vs the CPU oracle over the cross product of methods, step modes,
error controls, harmonic degree/order, lane counts, forward/backward spans and repeated calls.  Wherever libm is not involved
(no SRP/drag) the result must be BIT-IDENTICAL on (almost) every trajectory; the rare exception is a glibc-vs-corrected pow ulp."""
import numpy as np
import pytest

import nyx_b200 as nb
from tests.util import S, leo_ensemble, max_dr_dv, oracle_run

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.default_rng(seed)
    method = nb.IntegratorMethod(int(rng.integers(0, 6)))
    degree = int(rng.choice([0, 2, 5, 9, 16, 21]))
    order = int(rng.integers(0, degree + 1)) if degree and rng.random() < 0.4 else degree
    lanes = int(rng.choice([1, 8, 16, 32])) if degree >= 6 else 1
    fixed = bool(rng.random() < 0.15) or method == nb.IntegratorMethod.RungeKutta4
    if fixed:
        opts = nb.IntegratorOptions.with_fixed_step_s(float(rng.choice([5.0, 20.0, 45.5])))
    else:
        ctrl = nb.ErrorControl(int(rng.integers(0, 7)))
        opts = nb.IntegratorOptions.with_adaptive_step_s(0.01, float(rng.choice([60.0, 300.0, 2700.0])), float(rng.choice([1e-9, 1e-11, 1e-12])), ctrl)
        opts.init_step = int(rng.choice([10.0, 60.0])) * nb.Unit.Second
    third = bool(rng.random() < 0.4)
    spans = [int(s) for s in rng.choice([-2400, -600, 900, 1800, 4000, 7200], size=3)]
    return method, degree, order, lanes, opts, third, spans


@pytest.mark.parametrize("seed", range(32))
def test_strict_fuzz_bit_parity(oracle, seed):
    method, degree, order, lanes, opts, third, spans = _case(seed)
    frame = nb.EARTH_J2000
    almanac = nb.Almanac.synthetic(frame, 0, 1.0, pad_days=1.0) if third else None
    models = []
    if third:
        models.append(nb.PointMasses.new([nb.MOON, nb.SUN]))
    if degree:
        models.append(nb.GravityField.new(nb.GravityFieldData.from_fixture("jgm3_70x70", degree, order, nb.IAU_EARTH_FRAME)))
    dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.new(models))
    prop = nb.Propagator.new(dyn, method, opts, mode=nb.MODE_STRICT)
    eng = prop.engine(frame, almanac)
    eng.set_lanes(lanes)
    n = 24
    mc, (st, cs, ep) = leo_ensemble(n, seed=100 + seed)
    step_g = np.full(n, opts.init_step, dtype=np.int64)
    step_o = step_g.copy()
    cur_g, ep_g, cur_o, ep_o = st, ep, st, ep
    t = 0
    for span in spans:  # repeated calls on the same instances, forward and backward (instance.rs:112-115, 198-200)
        t += span * S
        cur_g, ep_g, det_g, sg = eng.propagate_batch(cur_g, cs, ep_g, t, step_g)
        cur_o, ep_o, det_o, so = oracle_run(oracle, prop, frame, almanac, cur_o, cs, ep_o, t, step_o)
        assert np.array_equal(sg, so) and np.array_equal(ep_g, ep_o) and (sg == 0).all(), (method, degree, order, lanes)
        same = (cur_g == cur_o).all(axis=0)
        assert same.mean() >= 0.9, (same.mean(), method, degree, order, lanes, opts)
        assert max_dr_dv(cur_g, cur_o)[0] < 1e-6
        assert np.array_equal(step_g[same], step_o[same])
        assert np.array_equal(det_g["n_steps"][same], det_o["n_steps"][same])
