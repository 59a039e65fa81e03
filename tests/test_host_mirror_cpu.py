"""The Python host mirror above the C ABI (PropInstance, Propagator.many_*, MonteCarlo incl. events and reports) exercised on
the CPU: `Propagator.engine` is routed to an oracle-backed stand-in with the Engine's host-facing methods (tests/util.py), so
everything except the ctypes call itself runs exactly as it does on a GPU box."""
import numpy as np
import pytest

import nyx_b200 as nb
from nyx_b200.event import Event
from nyx_b200.param import StateParameter as P
from tests.util import S, leo_ensemble, leo_state, use_oracle_engine


def _dyn(degree=8):
    gd = nb.GravityFieldData.from_fixture("jgm3_70x70", degree, degree, nb.IAU_EARTH_FRAME)
    return nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(nb.GravityField.new(gd)))


def test_prop_instance_keeps_the_adapted_step_and_reports_details(oracle, monkeypatch):
    prop = nb.Propagator.default(_dyn())
    use_oracle_engine(monkeypatch, oracle, prop)
    sc = leo_state(nb.EARTH_J2000)
    inst = prop.with_(sc)
    a = inst.for_duration(1800 * S)
    d1 = inst.latest_details()
    b = inst.for_duration(1800 * S)
    assert a.epoch() == 1800 * S and b.epoch() == 3600 * S and d1.n_steps > 0
    assert inst._step_ns[0] > 60 * S and inst.latest_details().step < inst._step_ns[0]   # the adapted step survives the final partial one (instance.rs:196-200)
    # one call over the whole hour follows a different step sequence only through the final partial step at 1800 s
    whole = prop.with_(sc).for_duration(3600 * S)
    assert np.linalg.norm(whole.orbit.radius_km - b.orbit.radius_km) < 1e-6
    fin, tr = prop.with_(sc).for_duration_with_traj(1800 * S)
    assert tr.first().epoch() == 0 and tr.last().epoch() == 1800 * S and np.array_equal(tr.states[-1], fin.orbit.to_cartesian_pos_vel())
    bad = nb.Spacecraft(orbit=sc.orbit, mass=nb.Mass(100.0, -1.0, 0.0))
    with pytest.raises(nb.PropagationError, match="FuelExhausted"):
        prop.with_(bad).for_duration(60 * S)


def test_many_until_epoch_and_for_duration(oracle, monkeypatch):
    """py_md.rs:224-320: batched entry points, failed runs dropped, trajectories on request."""
    prop = nb.Propagator.default(_dyn())
    cache = use_oracle_engine(monkeypatch, oracle, prop)
    mc, _ = leo_ensemble(5, seed=9)
    scs = [ds.state for _, ds in mc.generate_states(0, 5)]
    scs[2] = nb.Spacecraft(orbit=scs[2].orbit, mass=nb.Mass(100.0, -1.0, 0.0))      # FuelExhausted: dropped
    end = 1500 * S
    finals = prop.many_until_epoch(scs, end)
    assert len(finals) == 4 and all(f.epoch() == end for f in finals)
    with_traj = prop.many_until_epoch(scs, end, trajectory=True, traj_capacity=4)   # too small on purpose: grows
    assert len(with_traj) == 4
    for (state, traj), f in zip(with_traj, finals):
        assert np.array_equal(state.to_vector(), f.to_vector()) and traj.last().epoch() == end and len(traj) > 4
        assert np.array_equal(traj.states[-1], state.orbit.to_cartesian_pos_vel())
    assert prop.many_until_epoch([], end) == []
    # many_for_duration: spacecraft with different start epochs, one launch per distinct end epoch
    shifted = [nb.Spacecraft(orbit=nb.Orbit.cartesian(*s.orbit.to_cartesian_pos_vel(), (i % 2) * 600 * S, s.orbit.frame), mass=s.mass) for i, s in enumerate(scs)]
    eng = next(iter(cache.values()))
    before = eng.launch_count()
    res = prop.many_for_duration(shifted, 900 * S)
    assert [r.epoch() for r in res] == [900 * S, 1500 * S, 1500 * S, 900 * S]      # run 2 dropped, input order kept
    assert eng.launch_count() - before <= 4


def test_monte_carlo_reports_and_events_on_the_host_mirror(oracle, monkeypatch, tmp_path):
    frame = nb.EARTH_J2000
    tmpl = leo_state(frame)
    mc = nb.MonteCarlo(tmpl, nb.MvnSpacecraft.from_cartesian_std(tmpl, 1.0, 1e-3), "cpu-mirror", seed=4)
    prop = nb.Propagator.default(_dyn())
    use_oracle_engine(monkeypatch, oracle, prop, tmp_path)
    end = 2 * 3600 * S
    res = mc.run_until_epoch(prop, None, end, 8, traj_capacity=16)
    assert res.recording[0].shape[0] == int(res.details["n_steps"].max()) + 1 and len(res.ok_runs()) == 8
    sma = res.every_value_of(P.SemiMajorAxis, 600 * S)
    assert len(sma) == 8 * 13 and max(sma) - min(sma) < 40.0
    assert res.last_values_of(P.X) == res.final_state_soa[0].tolist()
    assert len(res.dispersion_values_of(P.VZ)) == 8
    tab_path = res.to_parquet(tmp_path / "mc.parquet", fields=[P.X, P.Rmag], step_ns=1200 * S)
    import pyarrow.parquet as pq
    assert pq.read_table(str(tab_path)).num_rows == 8 * 7
    # event-terminated ensemble: stop condition in the propagation, Brent search on the recording, all through the mirror
    ev = Event.apsis()
    evres = mc.run_until_nth_event(prop, None, 4 * 3600 * S, ev, 2, 8, traj_capacity=32)
    assert len(evres.ok_runs()) == 8
    for run in evres.runs:
        state, traj = run.result
        assert abs(ev.eval(state)) < 1e-3 and traj.epochs_ns[-2] <= state.epoch() <= traj.epochs_ns[-1]
    found, tr = prop.with_(tmpl).until_nth_event(4 * 3600 * S, ev, trigger=2)
    assert abs(ev.eval(found)) < 1e-3
    with pytest.raises(nb.PropagationError, match="NthEventError"):
        prop.with_(tmpl).until_nth_event(300 * S, Event.radius(30000.0), trigger=1)


def test_kalman_od_process_on_the_host_mirror(oracle, monkeypatch, tmp_path):
    """`KalmanODProcess.process_arcs` -> packing -> (stand-in for nyxb_od_ekf_batch) -> `ODSolution`, then the reference's
    data formats either side of it: tracking arcs from parquet in, OD solution parquet out."""
    import pyarrow.parquet as pq
    from tests.od_util import leo_od_scenario, run_oracle_filter
    from oracle import pyoracle_od
    sc = leo_od_scenario(oracle, n=3, n_msr=24, seed=2)
    odp, prop = sc["odp"], sc["prop"]
    use_oracle_engine(monkeypatch, oracle, prop)
    # the arcs travel through the reference's parquet layout and come back as the batched arc
    paths = [sc["arc"].to_parquet(tmp_path / f"trk{i}.parquet", index=i) for i in range(3)]
    arc = nb.TrackingDataArc.stack([nb.TrackingDataArc.from_parquet(p) for p in paths])
    assert np.array_equal(arc.epoch_ns, sc["arc"].epoch_ns) and arc.tracker == list(sc["arc"].tracker)
    assert np.array_equal(arc.obs, sc["arc"].obs, equal_nan=True)
    sol = odp.process_arcs(sc["ests"], arc, record_estimates=True)
    assert (sol.status == 0).all() and sol.accepted().sum() > 3 * 15
    for i in range(3):   # the mirror's packing (column-major covariance, tracker indices, per-filter observations) is faithful
        ref = run_oracle_filter(pyoracle_od, sc, i)
        assert np.array_equal(sol.final_state_soa[:, i], ref["state"]) and np.array_equal(sol.covar[i], ref["covar"])
        assert np.array_equal(sol.msr_flags[:, i], ref["msr_flags"])
    est = sol.final_estimate(1)
    truth_end = sc["truth"][-1, :3, 1]
    assert np.linalg.norm(est.nominal_state.orbit.radius_km - truth_end) < np.linalg.norm(sc["ests"][1].nominal_state.orbit.radius_km - sc["truth"][0, :3, 1]) + 1.0
    tab = pq.read_table(str(sol.to_parquet(tmp_path / "od1.parquet", index=1)))
    processed = int(((sol.msr_flags[:, 1] & 1) != 0).sum())
    assert tab.num_rows == processed and tab["Tracker"].to_pylist()[0] == arc.tracker[int(np.nonzero(sol.msr_flags[:, 1] & 1)[0][0])]
    assert np.isfinite(np.array(tab["Residual ratio"].to_pylist(), dtype=float)).all()
    sig_x = [c for c in tab.column_names if c.startswith("Sigma X (")]
    assert len(sig_x) == 1 and all(v > 0 for v in tab[sig_x[0]].to_pylist())
    with pytest.raises(nb.ODError, match="observation sets"):
        odp.process_arcs(sc["ests"][:2], arc)
