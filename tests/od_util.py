"""Shared scenario builders for the orbit-determination tests (CPU oracle and GPU parity)."""
import numpy as np

import nyx_b200 as nb

S = 10**9


def leo_od_scenario(oracle, n=1, n_msr=40, cadence_s=60, seed=0, degree=4, noise=True, msr_size=2, variant=None,
                    snc=True, reject=3.0, pos_err_km=0.8, vel_err_km_s=8e-4, method=None, elevation_mask_deg=-90.0):
    """LEO spacecraft tracked by three Earth stations (range + Doppler).  The truth is propagated with the CPU oracle;
    each of the n filters starts from its own dispersed initial estimate and sees its own noisy observations."""
    frame = nb.EARTH_J2000
    gd = nb.GravityFieldData.from_fixture("jgm3_70x70", degree, degree, nb.IAU_EARTH_FRAME)
    dyn = nb.SpacecraftDynamics.new(nb.OrbitalDynamics.from_model(nb.GravityField.new(gd)))
    prop = nb.Propagator.new(dyn, method or nb.IntegratorMethod.DormandPrince78, nb.IntegratorOptions.default())
    orbit = nb.Orbit.keplerian(7000.0, 0.01, 51.6, 30.0, 40.0, 10.0, 0, frame)
    truth0 = nb.Spacecraft(orbit=orbit, mass=nb.Mass(500.0, 50.0, 0.0))
    rn, dn = nb.StochasticNoise(1e-2), nb.StochasticNoise(1e-5)  # 10 m, 1 cm/s
    devices = {
        "Madrid": nb.GroundStation.dss65_madrid(elevation_mask_deg, rn, dn),
        "Canberra": nb.GroundStation.dss34_canberra(elevation_mask_deg, rn, dn),
        "Goldstone": nb.GroundStation.dss13_goldstone(elevation_mask_deg, rn, dn),
    }
    names = list(devices)
    epochs = (np.arange(1, n_msr + 1) * cadence_s * S).astype(np.int64)
    schedule = [names[(k // 10) % 3] for k in range(n_msr)]
    # truth: fixed 10 s RK89 steps on the oracle, recorded; sample every cadence
    topts = nb.IntegratorOptions.with_fixed_step_s(10.0)
    packed = dyn.pack(frame, None)
    st, cs, ep = nb.pack_spacecraft([truth0])
    cap = n_msr * cadence_s // 10 + 2
    _, _, _, status, (t_ep, t_st, t_cnt) = oracle.propagate_batch(packed.c, topts.to_c(nb.IntegratorMethod.RungeKutta89), st, cs, ep,
                                                                  int(epochs[-1]), traj_capacity=cap)
    assert status[0] == 0
    idx = np.searchsorted(t_ep[: t_cnt[0], 0], epochs)
    assert np.array_equal(t_ep[idx, 0], epochs)
    truth = np.repeat(t_st[:, idx, 0].T[:, :, None], n, axis=2)  # [m][6][n]
    rng = np.random.default_rng(seed)
    arc = nb.simulate_tracking(epochs, truth, devices, schedule, frame, None, rng if noise else None)
    # dispersed initial estimates
    ests = []
    for i in range(n):
        d = np.concatenate([rng.normal(0, pos_err_km, 3), rng.normal(0, vel_err_km_s, 3)])
        v = truth0.to_vector()
        v[:6] += d
        sc = truth0.with_vector(0, v)
        ests.append(nb.KfEstimate.from_diag(sc, [1.0, 1.0, 1.0, 1e-6, 1e-6, 1e-6, 0.0, 0.0, 0.0]))
    odp = nb.KalmanODProcess(prop, variant if variant is not None else nb.KalmanVariant.ReferenceUpdate,
                             nb.SigmaRejection(reject) if reject is not None else None, devices, None, msr_size=msr_size)
    if snc:
        odp.with_process_noise(nb.ProcessNoise3D.from_diagonal([1e-12, 1e-12, 1e-12], 10 * nb.Unit.Minute, nb.LocalFrame.RIC))
    return dict(frame=frame, dyn=dyn, prop=prop, devices=devices, arc=arc, ests=ests, odp=odp, truth=truth, epochs=epochs, packed=packed)


def run_oracle_filter(oracle_od, sc, i):
    """One filter of the scenario on the numpy oracle."""
    odp = sc["odp"]
    names, st_c = odp.stations_c(sc["frame"])
    arc = sc["arc"]
    tracker = np.array([names.index(t) if t in names else -1 for t in arc.tracker], dtype=np.int32)
    est = sc["ests"][i]
    y9 = est.nominal_state.to_vector()
    m = est.nominal_state.mass
    cs = np.array([m.dry_mass_kg, m.extra_mass_kg, est.nominal_state.srp.area_m2, est.nominal_state.drag.area_m2])
    return oracle_od.process_arc(sc["packed"].c, sc["prop"].opts.to_c(sc["prop"].method), odp.config_c(), st_c, arc.epoch_ns, tracker,
                                 np.ascontiguousarray(arc.obs[:, :, i]), y9, cs, est.nominal_state.epoch(), est.covar)
