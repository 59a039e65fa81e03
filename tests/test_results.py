"""Monte Carlo report accessors (mc/results.rs:88-427) over recorded trajectories and the batched resampling entry point
`nyxb_traj_resample`.

CPU tests drive `Results` with an engine stand-in whose `resample` is the kernel's own per-(query, trajectory) function
compiled for the host (tests/cpp/hermite_core_shim.cpp); GPU tests go through the C ABI."""
import numpy as np
import pytest

import nyx_b200 as nb
from nyx_b200.monte_carlo import DispersedState, Results, Run
from nyx_b200.param import EXPORT_PARAMS, StateError, StateParameter, evaluate
from nyx_b200.trajectory import Traj
from tests.util import S, hermite_shim, leo_ensemble, leo_state, resample_queries, resample_reference

P = StateParameter


def _dyn(degree=8):
    gd = nb.GravityFieldData.from_fixture("jgm3_70x70", degree, degree, nb.IAU_EARTH_FRAME)
    return nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(nb.GravityField.new(gd)))


class ShimEngine:
    """`Engine.resample` on the host shim, including the resident-recording convention."""

    def __init__(self, tmp_path):
        self.run = hermite_shim(tmp_path)
        self.resident = None
        self.uploads = 0

    def resample(self, q, recording=None, n=None):
        if recording is not None:
            self.resident = recording
            self.uploads += 1
        assert self.resident is not None and (n is None or n == self.resident[0].shape[1])
        return self.run(*self.resident, q)


# --------------------------------------------------------------------------------------------- StateParameter
def test_state_parameter_elements_round_trip():
    frame = nb.EARTH_J2000
    rng = np.random.default_rng(5)
    for _ in range(20):
        sma, ecc = rng.uniform(6800, 42000), rng.uniform(0.001, 0.7)
        inc, raan, aop, ta = rng.uniform(1, 179), rng.uniform(0, 360), rng.uniform(0, 360), rng.uniform(0, 360)
        rv = nb.Orbit.keplerian(sma, ecc, inc, raan, aop, ta, 0, frame).to_cartesian_pos_vel().reshape(6, 1)
        mu = frame.mu_km3_s2()
        got = {p: float(evaluate(p, rv, mu)[0]) for p in (P.SemiMajorAxis, P.Eccentricity, P.Inclination, P.RAAN, P.AoP, P.TrueAnomaly)}
        assert abs(got[P.SemiMajorAxis] - sma) < 1e-6 and abs(got[P.Eccentricity] - ecc) < 1e-10
        for p, want in ((P.Inclination, inc), (P.RAAN, raan), (P.AoP, aop), (P.TrueAnomaly, ta)):
            assert abs((got[p] - want + 180.0) % 360.0 - 180.0) < 1e-7, p
        assert abs(float(evaluate(P.AoL, rv, mu)[0]) - (aop + ta) % 360.0) < 1e-7
        assert abs(float(evaluate(P.TrueLongitude, rv, mu)[0]) - (aop + raan + ta) % 360.0) < 1e-7
        assert abs(float(evaluate(P.Period, rv, mu)[0]) - 2 * np.pi * np.sqrt(sma**3 / mu)) < 1e-6
        assert abs(float(evaluate(P.PeriapsisRadius, rv, mu)[0]) - sma * (1 - ecc)) < 1e-6
        assert abs(float(evaluate(P.Rmag, rv, mu)[0]) - np.linalg.norm(rv[:3, 0])) < 1e-9
        assert float(evaluate(P.VY, rv, mu)[0]) == rv[4, 0]


def test_state_parameter_constants_and_unavailable():
    sc = leo_state(nb.EARTH_J2000)
    rv = np.zeros((6, 3)) + sc.orbit.to_cartesian_pos_vel()[:, None]
    assert np.array_equal(evaluate(P.Cr, rv, 1.0, sc), np.full(3, sc.srp.coeff_reflectivity))
    assert np.array_equal(evaluate(P.Cd, rv, 1.0, sc, cd=[1.0, 2.0, 3.0]), [1.0, 2.0, 3.0])
    assert np.allclose(evaluate(P.TotalMass, rv, 1.0, sc, prop_mass_kg=5.0), sc.mass.dry_mass_kg + sc.mass.extra_mass_kg + 5.0)
    for p in (P.Isp, P.Thrust, P.GuidanceMode):
        with pytest.raises(StateError):
            evaluate(p, rv, 1.0, sc)
    with pytest.raises(StateError):
        evaluate(P.DryMass, rv, 1.0)   # needs a template
    assert str(P.DryMass) == "dry_mass (kg)" and P.DryMass.unit == "kg" and P.Eccentricity.unit == "" and str(P.Cr) == "cr"
    assert P.Isp not in EXPORT_PARAMS and P.X in EXPORT_PARAMS


def test_epochs_to_utc_iso():
    iso = nb.epochs_to_utc_iso([0, 20 * 365 * 86400 * S, -366 * 86400 * S])
    assert iso[0] == "2000-01-01T11:58:55.816000000"      # J2000 TT noon = 11:58:55.816 UTC (TAI-UTC = 32 s)
    assert iso[1] == "2019-12-27T11:58:50.816000000"      # TAI-UTC = 37 s
    assert iso[2] == "1998-12-31T11:58:56.816000000"      # TAI-UTC = 31 s


# --------------------------------------------------------------------------------------------- Results on recordings
def _results_from_oracle(oracle, tmp_path, end=3 * 3600 * S, n=6, fail=(4,)):
    frame = nb.EARTH_J2000
    mc, (st, cs, ep) = leo_ensemble(n, seed=3)
    prop = nb.Propagator.default(_dyn())
    packed = prop.dynamics.pack(frame, None)
    out, out_ep, det, status, rec = oracle.propagate_batch(packed.c, prop.opts.to_c(prop.method), st, cs, ep, end, traj_capacity=256)
    status = status.copy()
    runs = []
    tmpl = leo_state(frame)
    for i in range(n):
        sc = tmpl.with_vector(0, st[:, i])
        ds = DispersedState(sc, [(p, float(st[q, i] - tmpl.to_vector()[q])) for q, p in enumerate(("X", "Y", "Z", "VX", "VY", "VZ", "Cr", "Cd", "PropMass"))])
        if i in fail:
            status[i] = 1
            runs.append(Run(i, ds, nb.PropagationError("PropMathError")))
        else:
            runs.append(Run(i, ds, sc.with_vector(int(out_ep[i]), out[:, i])))
    return Results(runs, "test", out, det, status, rec, ShimEngine(tmp_path)), rec, tmpl, end


def _host_every(res, rec, tmpl, param, step, start=None, end=None, fill=None):
    """results.rs:90-160 literally: per run, Traj::every_between -> value."""
    t_ep, t_st, t_cnt = rec
    mu = tmpl.orbit.frame.mu_km3_s2()
    want = []
    for run in res.runs:
        if isinstance(run.result, Exception):
            if fill is not None:
                want.append(fill)
            continue
        k = int(t_cnt[run.index])
        tr = Traj(tmpl, t_ep[:k, run.index].copy(), np.ascontiguousarray(t_st[:, :k, run.index].T)).finalize()
        s = int(tr.epochs_ns[0]) if start is None else max(int(start), int(tr.epochs_ns[0]))
        e = int(tr.epochs_ns[-1]) if end is None else min(int(end), int(tr.epochs_ns[-1]))
        t = s
        while t <= e:
            rv = tr.at(t).orbit.to_cartesian_pos_vel().reshape(6, 1)
            want.append(float(evaluate(param, rv, mu, run.dispersed_state.state)[0]))
            t += step
    return want


def test_results_every_value_of_matches_per_run_traj_iteration(oracle, tmp_path):
    res, rec, tmpl, end = _results_from_oracle(oracle, tmp_path)
    step = 7 * 60 * S + 13
    for param in (P.X, P.SemiMajorAxis, P.Inclination):
        assert res.every_value_of(param, step) == _host_every(res, rec, tmpl, param, step)
    got = res.every_value_of_between(P.VZ, step, 1000 * S, end - 500 * S, value_if_run_failed=-1.0)
    assert got == _host_every(res, rec, tmpl, P.VZ, step, 1000 * S, end - 500 * S, fill=-1.0)
    assert got.count(-1.0) == 1
    # bounds beyond the trajectory are clamped to its span (traj.rs:155-158); an empty window gives no values
    assert res.every_value_of_between(P.X, step, -10 * S, end + 10 * S) == res.every_value_of(P.X, step)
    assert res.every_value_of_between(P.X, step, end + 1, end + 2) == []
    # a parameter no state provides: nothing, or the fill value once per state
    assert res.every_value_of(P.Isp, step) == []
    n_states = len(res.every_value_of(P.X, step))
    assert res.every_value_of(P.Isp, step, value_if_run_failed=0.5) .count(0.5) == n_states + 1
    with pytest.raises(ValueError):
        res.every_value_of(P.X, 0)


def test_results_resamples_in_chunks_with_one_upload(oracle, tmp_path, monkeypatch):
    import nyx_b200.monte_carlo as mcmod
    res, rec, tmpl, end = _results_from_oracle(oracle, tmp_path, fail=())
    whole = res.every_value_of(P.Y, 60 * S)
    monkeypatch.setattr(mcmod, "_RESAMPLE_CHUNK_BYTES", 48 * 6 * 7)   # 7 grid points per launch
    res.engine.uploads = 0
    assert res.every_value_of(P.Y, 60 * S) == whole
    assert res.engine.uploads == 1


def test_results_first_last_dispersions(oracle, tmp_path):
    res, rec, tmpl, end = _results_from_oracle(oracle, tmp_path)
    ok = [r for r in res.runs if not isinstance(r.result, Exception)]
    assert res.first_values_of(P.X) == [r.dispersed_state.state.orbit.x_km for r in ok]
    assert res.last_values_of(P.VX) == [r.result.orbit.vx_km_s for r in ok]
    assert res.last_values_of(P.VX, value_if_run_failed=9.0)[4] == 9.0 and len(res.last_values_of(P.VX, 9.0)) == len(res.runs)
    assert res.first_values_of(P.Cr) == [r.dispersed_state.state.srp.coeff_reflectivity for r in ok]
    assert res.dispersion_values_of(P.X) == [dict(r.dispersed_state.actual_dispersions)["X"] for r in res.runs]
    with pytest.raises(nb.MonteCarloError):
        res.dispersion_values_of(P.SemiMajorAxis)
    bare = Results(res.runs, "bare", res.final_state_soa, res.details, res.status)
    assert bare.last_values_of(P.X) == res.last_values_of(P.X)
    with pytest.raises(nb.MonteCarloError):
        bare.every_value_of(P.X, 60 * S)


def test_results_to_parquet(oracle, tmp_path):
    import pyarrow.parquet as pq
    res, rec, tmpl, end = _results_from_oracle(oracle, tmp_path)
    t_ep, t_st, t_cnt = rec
    # raw export: all recorded states of the successful runs
    path = res.to_parquet(tmp_path / "mc_raw.parquet")
    tab = pq.read_table(str(path))
    ok = [r.index for r in res.runs if not isinstance(r.result, Exception)]
    assert tab.num_rows == int(sum(t_cnt[i] for i in ok))
    assert tab.column_names[:2] == ["Epoch (UTC)", "Monte Carlo Run Index"]
    assert tab.column_names[2:] == [str(p) for p in EXPORT_PARAMS]
    assert sorted(set(tab["Monte Carlo Run Index"].to_pylist())) == ok
    first = ok[0]
    k = int(t_cnt[first])
    assert np.array_equal(np.array(tab["X (km)"].to_pylist()[:k]), t_st[0, :k, first])
    assert tab["Epoch (UTC)"][0].as_py() == nb.epochs_to_utc_iso([0])[0]
    assert tab.schema.metadata[b"Purpose"] == b"Monte Carlo Trajectory data"
    assert tab.schema.field("X (km)").metadata == {b"unit": b"km", b"Frame": tmpl.orbit.frame.name.encode()}
    # interpolated export on a grid, selected fields, extra metadata; thruster fields are dropped (no state provides them)
    path = res.to_parquet(tmp_path / "mc_grid.parquet", fields=[P.X, P.Isp, P.Eccentricity], step_ns=600 * S, metadata={"who": "test"})
    tab = pq.read_table(str(path))
    assert tab.column_names == ["Epoch (UTC)", "Monte Carlo Run Index", "X (km)", "Eccentricity"]
    assert tab["X (km)"].to_pylist() == res.every_value_of(P.X, 600 * S)
    assert tab.schema.metadata[b"who"] == b"test"
    failed = Results([Run(0, res.runs[4].dispersed_state, res.runs[4].result)], "f", None, None, None, rec, res.engine)
    with pytest.raises(nb.MonteCarloError):
        failed.to_parquet(tmp_path / "none.parquet")


# --------------------------------------------------------------------------------------------- GPU: through the C ABI
def _gpu_recording(mode, end, cap=256, n=9, seed=21):
    frame = nb.EARTH_J2000
    mc, (st, cs, ep) = leo_ensemble(n, seed=seed)
    prop = nb.Propagator.default(_dyn(), mode=mode)
    eng = prop.engine(frame, None)
    out, out_ep, det, status, rec = eng.propagate_batch(st, cs, ep, end, traj_capacity=cap)
    assert (status == 0).all()
    return eng, rec, leo_state(frame)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [nb.MODE_STRICT, nb.MODE_FAST])
@pytest.mark.parametrize("end", [3 * 3600 * S, -2 * 3600 * S])
def test_gpu_resample_is_bit_identical_to_traj_at(mode, end):
    """nyxb_traj_resample == Traj.at on every (query, trajectory): exact hits, both window edges, out-of-span queries,
    forward and backward recordings, from the uploaded and from the resident recording."""
    eng, (t_ep, t_st, t_cnt), sc = _gpu_recording(mode, end)
    hi = max(0, end)
    queries = resample_queries(t_ep, t_cnt, hi) if end > 0 else \
        np.concatenate([np.array([end - 1, end, end + 1, -1, 0, 1, int(t_ep[2, 1])], dtype=np.int64),
                        np.arange(end, 0, 450 * S, dtype=np.int64) + 987_654_321])
    want, want_status = resample_reference(sc, t_ep, t_st, t_cnt, queries)
    before = eng.launch_count()
    got, got_status = eng.resample(queries, n=t_ep.shape[1])            # still resident from the propagation
    assert eng.launch_count() == before + 1
    assert np.array_equal(got_status, want_status) and (want_status == 1).any()
    ok = want_status == 0
    assert np.array_equal(got[:, ok], want[:, ok]) and np.isnan(got[:, ~ok]).all()
    got2, got2_status = eng.resample(queries, (t_ep, t_st, t_cnt))     # uploaded
    assert np.array_equal(got2_status, got_status) and np.array_equal(got2[:, ok], got[:, ok])


@pytest.mark.gpu
def test_gpu_resample_ragged_and_arguments():
    eng, (t_ep, t_st, t_cnt), sc = _gpu_recording(nb.MODE_FAST, 3 * 3600 * S)
    t_cnt = t_cnt.copy()
    t_cnt -= 4 * np.arange(len(t_cnt))
    t_cnt[8], t_cnt[7] = 0, 9      # nothing recorded; fewer records than one window
    queries = resample_queries(t_ep, t_cnt, 3 * 3600 * S)
    want, want_status = resample_reference(sc, t_ep, t_st, t_cnt, queries)
    got, got_status = eng.resample(queries, (t_ep, t_st, t_cnt))
    ok = want_status == 0
    assert np.array_equal(got_status, want_status) and np.array_equal(got[:, ok], want[:, ok])
    assert (got_status[:, 8] == 1).all()
    # empty query list; wrong resident size; malformed recording
    out, st = eng.resample(np.empty(0, dtype=np.int64), (t_ep, t_st, t_cnt))
    assert out.shape == (6, 0, 9) and st.shape == (0, 9)
    with pytest.raises(nb.PropagationError):
        eng.resample(queries, n=5)
    with pytest.raises(ValueError):
        eng.resample(queries, (t_ep, t_st[:5], t_cnt))
    with pytest.raises(ValueError):
        eng.resample(queries)
    fresh = nb.Propagator.default(_dyn(4), mode=nb.MODE_FAST).engine(nb.EARTH_J2000, None)
    with pytest.raises(nb.PropagationError):
        fresh.resample(queries, n=9)       # this engine never recorded anything


@pytest.mark.gpu
def test_gpu_monte_carlo_reports():
    """MonteCarlo.run_until_epoch(traj_capacity) -> Results: the report accessors on the device against the reference's
    per-run Traj iteration on the host; the sink grows when a run overflows it."""
    frame = nb.EARTH_J2000
    tmpl = leo_state(frame)
    rv_std = nb.MvnSpacecraft.from_cartesian_std(tmpl, 1.0, 1e-3)
    mc = nb.MonteCarlo(tmpl, rv_std, "reports", seed=11)
    prop = nb.Propagator.default(_dyn(), mode=nb.MODE_FAST)
    end = 2 * 3600 * S
    res = mc.run_until_epoch(prop, None, end, 12, traj_capacity=16)   # too small on purpose
    t_ep, t_st, t_cnt = res.recording
    assert t_ep.shape[0] == int(res.details["n_steps"].max()) + 1 and np.array_equal(t_cnt, res.details["n_steps"] + 1)
    step = 9 * 60 * S
    for param in (P.X, P.Rmag, P.Eccentricity):
        assert res.every_value_of(param, step) == _host_every(res, res.recording, tmpl, param, step)
    assert res.every_value_of_between(P.Z, step, 600 * S, end - 1) == _host_every(res, res.recording, tmpl, P.Z, step, 600 * S, end - 1)
    assert res.last_values_of(P.X) == res.final_state_soa[0].tolist()
    plain = mc.run_until_epoch(prop, None, end, 12)
    assert np.array_equal(plain.final_state_soa, res.final_state_soa) and plain.recording is None
    dev = mc.run_until_epoch(prop, None, end, 12, device_dispersions=True, traj_capacity=128)
    assert len(dev.every_value_of(P.X, step)) == len(res.every_value_of(P.X, step))
