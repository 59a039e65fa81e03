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
    {   // FuelExhausted surfaces as a PropagationError from PropInstance (spacecraft.rs:163-168)
        Spacecraft bad = init; bad.prop_mass_kg = -1.0;
        bool threw = false;
        try { Propagator::default_(dynamics).with(bad).for_duration(60 * NS_PER_S); } catch (const PropagationError& e) { threw = e.status == NYXB_ERR_FUEL_EXHAUSTED; }
        CHECK(threw);
    }
    std::printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
    return failures ? 1 : 0;
}
