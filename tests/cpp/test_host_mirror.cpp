// C++ host-mirror test: reads like the reference's own propagator tests (tests/propagation/propagators.rs,
// tests/mission_design/orbitaldyn.rs) and goes through nyxb.hpp -> C ABI -> CUDA kernels.
#include <cstdio>
#include <cstring>

#include "nyxb.hpp"

using namespace nyxb;
static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

static bool same6(const Spacecraft& s, const double (&g)[6]) {
    const double v[6] = {s.x_km, s.y_km, s.z_km, s.vx_km_s, s.vy_km_s, s.vz_km_s};
    return std::memcmp(v, g, sizeof(v)) == 0;  // bit-exact
}

int main() {
    const double GMAT_EARTH_GM = 398600.4415;  // tests/propagation/mod.rs:1
    const Frame eme2k = EARTH_J2000().with_mu_km3_s2(GMAT_EARTH_GM);
    const Spacecraft init = Spacecraft::cartesian(-2436.45, -2436.45, 6891.037, 5.088611, -5.088611, 0.0, 0, eme2k);
    const auto dynamics = SpacecraftDynamics::new_(OrbitalDynamics::two_body());

    {   // gmat_val_leo_day_fixed, RK89 10 s (propagators.rs:360-369, assert :464): bit-exact
        auto setup = Propagator::rk89(dynamics, IntegratorOptions::with_fixed_step_s(10.0));
        auto prop = setup.with(init);
        auto fin = prop.for_duration(days(1));
        const double gold[6] = {-5971.19419167081, 3945.5066532332503, 2864.6366184022418, 0.049096957620019005, -4.185093318469214, 5.848940867753748};
        CHECK(same6(fin, gold));
        CHECK(prop.latest_details().n_steps == 8640);
    }
    {   // gmat_val_leo_day_adaptive, RK89 (propagators.rs:135-144, assert_eq :283-287): bit-exact
        auto setup = Propagator::rk89(dynamics, IntegratorOptions::with_adaptive_step_s(0.1, 30.0, 1e-12, ErrorControl::RSSCartesianState));
        auto prop = setup.with(init);
        auto fin = prop.for_duration(days(1));
        const double gold[6] = {-5971.194191670676, 3945.506653225158, 2864.6366184134445, 0.04909695762999346, -4.185093318475795, 5.848940867748944};
        CHECK(same6(fin, gold));
        // back-propagation restores the epoch (orbitaldyn.rs:139-153)
        auto back = prop.for_duration(-days(1));
        CHECK(back.epoch() == 0);
        CHECK(std::fabs(back.x_km - init.x_km) < 1e-5);
    }
    {   // val_two_body_dynamics (orbitaldyn.rs:102-137): default options, pck08 Earth GM
        auto setup = Propagator::default_(dynamics);
        auto fin = setup.with(Spacecraft::cartesian(-2436.45, -2436.45, 6891.037, 5.088611, -5.088611, 0.0, 0, EARTH_J2000())).for_duration(days(1));
        const double gold[6] = {-5971.194375461378, 3945.517831291771, 2864.6210708007134, 0.04908320163379219, -4.1850841921806206, 5.848947414864886};
        CHECK(same6(fin, gold));
    }
    {   // Monte Carlo (tests/monte_carlo/framework.rs:22-95 shape): J2 field, fast mode, per-run errors do not abort
        auto j2 = GravityFieldData::from_j2(-4.84165374886470e-4, IAU_EARTH());
        auto dyn = SpacecraftDynamics::new_(OrbitalDynamics::from_model(GravityField{j2}));
        auto prop = Propagator::default_(dyn).with_mode(NYXB_MODE_FAST);
        Spacecraft nominal = init;
        nominal.frame = EARTH_J2000();
        nominal.dry_mass_kg = 100.0;
        const double sd[9] = {1.0, 1.0, 1.0, 1e-3, 1e-3, 1e-3, 0, 0, 0};
        MonteCarlo mc(nominal, sd, "cpp-mc", 0);
        auto res = mc.run_until_epoch(prop, nullptr, 3600 * NS_PER_S, 100);
        CHECK(res.runs.size() == 100);
        for (auto& r : res.runs) CHECK(std::holds_alternative<Spacecraft>(r.result) && std::get<Spacecraft>(r.result).epoch() == 3600 * NS_PER_S);
        auto tail = mc.resume_run_until_epoch(prop, nullptr, 90, 3600 * NS_PER_S, 10);
        CHECK(std::get<Spacecraft>(tail.runs[0].result).x_km == std::get<Spacecraft>(res.runs[90].result).x_km);
        CHECK(res.total_steps > 100 * 30);
    }
    {   // multi-device fan-out behind the ABI (two engines on device 0 here): contiguous shards, same results as the single call
        const double sd2[9] = {1.0, 1.0, 1.0, 1e-3, 1e-3, 1e-3, 0, 0, 0};
        MonteCarlo mc2(init, sd2, "multi", 7);
        auto states = mc2.generate_states(0, 37);
        auto prop2 = Propagator::default_(dynamics);
        auto one = prop2.propagate_batch(states, 1800 * NS_PER_S);
        auto two = prop2.propagate_batch_multi(states, 1800 * NS_PER_S, {0, 0});
        CHECK(std::memcmp(one.state.data(), two.state.data(), one.state.size() * sizeof(double)) == 0);
        CHECK(one.epoch == two.epoch && one.status == two.status);
        prop2.kernel = NYXB_KERNEL_THREAD;   // explicit kernel family through nyxb_engine_set_kernel
        auto thr = prop2.propagate_batch(states, 1800 * NS_PER_S);
        CHECK(std::memcmp(one.state.data(), thr.state.data(), one.state.size() * sizeof(double)) == 0);   // two-body: the per-thread kernel either way
    }
    {   // FuelExhausted surfaces as a PropagationError from PropInstance (spacecraft.rs:163-168)
        Spacecraft bad = init; bad.prop_mass_kg = -1.0;
        bool threw = false;
        try { Propagator::default_(dynamics).with(bad).for_duration(60 * NS_PER_S); } catch (const PropagationError& e) { threw = e.status == NYXB_ERR_FUEL_EXHAUSTED; }
        CHECK(threw);
    }
    {   // STM (tests/propagation/stm.rs:33-118 shape): RK4 fixed 10 s, ten steps, two-body; chaining through stm_in reproduces the run
        auto setup = Propagator::new_(dynamics, IntegratorMethod::RungeKutta4, IntegratorOptions::with_fixed_step_s(10.0));
        auto r = setup.propagate_batch_stm({init}, 100 * NS_PER_S);
        CHECK(r.status[0] == 0 && r.details[0].n_steps == 10);
        CHECK(std::fabs(r.phi(1, 0, 0, 3) - 100.0) < 0.2 && std::fabs(r.phi(1, 0, 0, 0) - 1.0) < 0.02);   // dr/dv ~ t, dr/dr ~ 1
        CHECK(r.phi(1, 0, 6, 6) == 1.0 && r.phi(1, 0, 3, 6) == 0.0);
        auto h1 = setup.propagate_batch_stm({init}, 50 * NS_PER_S);
        Spacecraft mid = init; mid.x_km = h1.state[0]; mid.y_km = h1.state[1]; mid.z_km = h1.state[2]; mid.vx_km_s = h1.state[3]; mid.vy_km_s = h1.state[4];
        mid.vz_km_s = h1.state[5]; mid.epoch_ns = h1.epoch[0];
        auto h2 = setup.propagate_batch_stm({mid}, 100 * NS_PER_S, nullptr, &h1.stm);
        CHECK(std::memcmp(h2.stm.data(), r.stm.data(), 81 * sizeof(double)) == 0);
    }
    {   // trajectory recording + batched resampling (Traj::at, traj.rs:83-126): exact hits return the stored records bit for bit,
        // epochs in between are interpolated, epochs outside the recorded span are flagged per entry
        auto setup = Propagator::rk89(dynamics, IntegratorOptions::with_fixed_step_s(10.0));
        Spacecraft other = init; other.x_km += 1.0;
        auto tb = setup.propagate_batch_traj({init, other}, 300 * NS_PER_S, 64);
        CHECK(tb.t_count[0] == 31 && tb.t_count[1] == 31 && tb.epoch_at(30, 1) == 300 * NS_PER_S);
        CHECK(tb.state_at(0, 0, 1) == other.x_km && tb.state_at(0, 30, 0) == tb.state[0]);
        auto rs = tb.every({-1, 0, 120 * NS_PER_S, 125 * NS_PER_S, 300 * NS_PER_S, 300 * NS_PER_S + 1});
        CHECK(!rs.ok(0, 0) && !rs.ok(5, 1) && rs.ok(1, 0) && rs.ok(4, 1) && std::isnan(rs.at(0, 0, 0)));
        CHECK(rs.at(0, 2, 0) == tb.state_at(0, 12, 0) && rs.at(4, 2, 1) == tb.state_at(4, 12, 1) && rs.at(2, 4, 1) == tb.state[2 * 2 + 1]);
        const double mid = 0.5 * (tb.state_at(0, 12, 0) + tb.state_at(0, 13, 0));   // the chord misses the arc by ~ a h^2 / 8 <~ 0.1 km
        CHECK(std::fabs(rs.at(0, 3, 0) - mid) < 0.2 && rs.at(0, 3, 0) != mid);
        auto fine = setup.with(init).for_duration(125 * NS_PER_S);   // 12 steps of 10 s + a final 5 s step
        CHECK(std::fabs(rs.at(0, 3, 0) - fine.x_km) < 1e-7 && std::fabs(rs.at(5, 3, 0) - fine.vz_km_s) < 1e-10);
    }
    {   // until_nth_event for a batch (event.rs:88-211): stop at the second node crossing, then locate z = 0 inside the last step
        auto setup = Propagator::default_(dynamics);
        Spacecraft other = init; other.vz_km_s += 1e-3;
        std::vector<int32_t> crossings(2, 0);
        nyxb_event ev{NYXB_EVENT_Z, 2, 0.0, crossings.data()};
        auto tb = setup.propagate_batch_traj({init, other}, days(1), 512, nullptr, &ev);
        CHECK(tb.status[0] == 0 && tb.status[1] == 0 && crossings[0] == 2 && crossings[1] == 2 && tb.epoch[0] < days(1));
        auto loc = tb.locate(NYXB_EVENT_Z, 0.0);
        for (size_t i = 0; i < 2; ++i) {
            const int64_t k = tb.t_count[i];
            CHECK(loc.status[i] == NYXB_TRAJ_OK && loc.epoch[i] >= tb.epoch_at(k - 2, i) && loc.epoch[i] <= tb.epoch_at(k - 1, i));
            CHECK(std::fabs(loc.state[2 * 2 + i]) < 5e-3);                                   // z at the event, km: within precision x |vz|
            CHECK(tb.state_at(2, k - 2, i) * tb.state_at(2, k - 1, i) < 0.0);                // the last step brackets the crossing
        }
        std::vector<int32_t> none(2, 0);
        nyxb_event far{NYXB_EVENT_RMAG, 1, 50000.0, none.data()};
        auto miss = setup.propagate_batch_traj({init, other}, 600 * NS_PER_S, 64, nullptr, &far);
        CHECK((miss.status[0] & 0xff) == NYXB_ERR_EVENT_NOT_FOUND && miss.locate(NYXB_EVENT_RMAG, 50000.0).status[0] == NYXB_TRAJ_NO_DATA);
    }
    {   // device dispersions: shard-invariant, right spread
        Spacecraft nominal = init; nominal.frame = EARTH_J2000(); nominal.dry_mass_kg = 100.0;
        const double sd[9] = {1.0, 1.0, 1.0, 1e-3, 1e-3, 1e-3, 0, 0, 0};
        MonteCarlo mc(nominal, sd, "cpp-mvn", 5);
        auto all = mc.generate_states_on_device(0, 2000), tail = mc.generate_states_on_device(1500, 500);
        CHECK(all[1500].x_km == tail[0].x_km && all[1999].vz_km_s == tail[499].vz_km_s);
        double m = 0, v = 0;
        for (auto& sc : all) m += sc.x_km - nominal.x_km;
        m /= 2000;
        for (auto& sc : all) v += (sc.x_km - nominal.x_km - m) * (sc.x_km - nominal.x_km - m);
        CHECK(std::fabs(m) < 0.1 && std::fabs(std::sqrt(v / 2000) - 1.0) < 0.1);
    }
    {   // sequential filter (tests/orbit_determination/two_body.rs shape): EKF on synthetic range + Doppler from three DSN stations
        const Frame eme = EARTH_J2000();
        Spacecraft truth = Spacecraft::cartesian(-2436.45, -2436.45, 6891.037, 5.088611, -5.088611, 0.0, 0, eme);
        truth.dry_mass_kg = 500.0;
        const int m = 30;
        std::vector<GroundStation> dev{GroundStation::dss65_madrid(-90.0, {1e-2, 0}, {1e-5, 0}), GroundStation::dss34_canberra(-90.0, {1e-2, 0}, {1e-5, 0}),
                                       GroundStation::dss13_goldstone(-90.0, {1e-2, 0}, {1e-5, 0})};
        // truth states every 60 s (one trajectory, restartable PropInstance), noise-free observations computed with the same geometry
        auto tset = Propagator::rk89(dynamics, IntegratorOptions::with_fixed_step_s(10.0));
        auto tprop = tset.with(truth);
        TrackingDataArc arc; arc.n = 2; arc.obs.assign((size_t)m * 2 * 2, 0.0);
        Spacecraft last = truth;
        for (int k = 0; k < m; ++k) {
            last = tprop.for_duration(60 * NS_PER_S);
            const GroundStation& gs = dev[(k / 10) % 3];
            double p[3], up[3]; gs.body_fixed(p, up);
            // IAU Earth with ra0 = 0, dec0 = 90 deg: inertial -> fixed = R3(W) R1(0) R3(90 deg) = R3(W + 90 deg), W = 190.147 + 360.9856235 d
            // (the secular pole drift is < 1e-5 rad over this arc: ignored in this synthetic data)
            const double d = (double)last.epoch_ns * 1e-9 / 86400.0, w = std::fmod(190.147 + 90.0 + 360.9856235 * d, 360.0) * 3.14159265358979323846 / 180.0;
            const double wd = 360.9856235 * 3.14159265358979323846 / 180.0 / 86400.0;
            const double rtx[3] = {std::cos(w) * p[0] - std::sin(w) * p[1], std::sin(w) * p[0] + std::cos(w) * p[1], p[2]};
            const double vtx[3] = {-wd * rtx[1], wd * rtx[0], 0.0};
            const double dr[3] = {last.x_km - rtx[0], last.y_km - rtx[1], last.z_km - rtx[2]}, dv[3] = {last.vx_km_s - vtx[0], last.vy_km_s - vtx[1], last.vz_km_s - vtx[2]};
            const double rng = std::sqrt(dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
            arc.epoch_ns.push_back(last.epoch_ns); arc.tracker.push_back(gs.name);
            for (size_t i = 0; i < 2; ++i) {
                arc.obs[((size_t)k * 2 + 0) * 2 + i] = rng;
                arc.obs[((size_t)k * 2 + 1) * 2 + i] = (dr[0] * dv[0] + dr[1] * dv[1] + dr[2] * dv[2]) / rng;
            }
        }
        Spacecraft e0 = truth, e1 = truth;
        e0.x_km += 0.5; e0.vy_km_s += 4e-4; e1.z_km -= 0.6; e1.vx_km_s -= 3e-4;
        const double p0[9] = {1, 1, 1, 1e-6, 1e-6, 1e-6, 0, 0, 0};
        KalmanODProcess odp(Propagator::default_(dynamics), KalmanVariant::ReferenceUpdate, std::nullopt, dev);
        const double q[3] = {1e-12, 1e-12, 1e-12};
        odp.with_process_noise(ProcessNoise3D::from_diagonal(q, 600 * NS_PER_S, true));
        auto sol = odp.process_arcs({KfEstimate::from_diag(e0, p0), KfEstimate::from_diag(e1, p0)}, arc);
        for (size_t i = 0; i < 2; ++i) {
            CHECK(sol.status[i] == 0 && sol.epoch[i] == last.epoch_ns);
            Spacecraft f = sol.final_state(truth, i);
            const double err = std::sqrt((f.x_km - last.x_km) * (f.x_km - last.x_km) + (f.y_km - last.y_km) * (f.y_km - last.y_km) + (f.z_km - last.z_km) * (f.z_km - last.z_km));
            CHECK(err < 0.1);   // from 0.5 / 0.6 km
            CHECK(sol.covar[(size_t)(0 * 9 + 0) * 2 + i] < 0.5);   // from 1.0 km^2
        }
        int processed = 0;
        for (int k = 0; k < m; ++k) processed += (sol.msr_flags[(size_t)k * 2] & NYXB_MSRF_PROCESSED) ? 1 : 0;
        CHECK(processed == m);
    }
    std::printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
    return failures ? 1 : 0;
}
