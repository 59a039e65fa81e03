// nyxb.hpp — header-only C++17 host mirror of the reference's Rust surface for the propagation path,
// layered on the C ABI of nyxb.h (libnyxb.so).  The reference is compiled (Rust) code and its toolchain is not
// available in the build image, so this is the compiled-language host side a nyx maintainer would port 1:1 to a
// Rust shim.  Names, argument meaning and error behaviour follow the reference:
//
//   IntegratorMethod / ErrorControl / IntegratorOptions   propagators/rk_methods/mod.rs:65-79, error_ctrl.rs:30-71, options.rs:42-186
//   Propagator::{new_,rk89,dp78,default_,with}            propagators/propagator.rs:55-118
//   PropInstance::{for_duration,until_epoch,latest_details}   propagators/instance.rs:265-282, 495-498
//   SpacecraftDynamics / OrbitalDynamics / PointMasses / GravityField / SolarPressure / Drag   dynamics/*.rs
//   MonteCarlo::{generate_states,run_until_epoch,resume_run_until_epoch}   mc/montecarlo.rs:188-296
//   Propagator::propagate_batch_stm (Spacecraft::with_stm + propagate)       dynamics/spacecraft.rs:203-227, 312-363
//   GroundStation / StochasticNoise / ProcessNoise3D / SigmaRejection / KfEstimate / TrackingDataArc / KalmanODProcess
//                                                         od/ground_station, od/snc.rs, od/process/{mod,initializers,rejectcrit}.rs
//
// No arithmetic of the hot path lives here: every propagate call is one nyxb_propagate_batch on the GPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "nyxb.h"

namespace nyxb {

constexpr int64_t NS_PER_S = 1000000000LL;
// hifitime `f64 * Unit::Second`: truncation toward zero
inline int64_t seconds(double s) { return (int64_t)(s * 1e9); }
inline int64_t days(int64_t d) { return d * 86400 * NS_PER_S; }

enum class IntegratorMethod : int32_t { RungeKutta89 = NYXB_RK89, DormandPrince78 = NYXB_DP78, DormandPrince45 = NYXB_DP45,
                                        RungeKutta4 = NYXB_RK4, CashKarp45 = NYXB_CK45, Verner56 = NYXB_V56 };
enum class ErrorControl : int32_t { RSSCartesianState = NYXB_RSS_CARTESIAN_STATE, RSSCartesianStep = NYXB_RSS_CARTESIAN_STEP,
                                    RSSState = NYXB_RSS_STATE, RSSStep = NYXB_RSS_STEP, LargestError = NYXB_LARGEST_ERROR,
                                    LargestState = NYXB_LARGEST_STATE, LargestStep = NYXB_LARGEST_STEP };

// PropagationError (propagators/mod.rs:68-92) wrapping DynamicsError (dynamics/mod.rs:177-203)
struct PropagationError : std::runtime_error {
    int32_t status;
    explicit PropagationError(int32_t st) : std::runtime_error(describe(st)), status(st) {}
    static std::string describe(int32_t st) {
        switch (st & 0xff) {
        case NYXB_ERR_PROP_MATH: return "PropMathError: part of state vector is NaN";
        case NYXB_ERR_FUEL_EXHAUSTED: return "DynamicsError::FuelExhausted";
        case NYXB_ERR_MASSLESS: return "DynamicsError::MasslessSpacecraft";
        case NYXB_ERR_EPHEMERIS: return "DynamicsError::DynamicsAlmanacError (epoch outside ephemeris coverage)";
        default: return "propagation error " + std::to_string(st);
        }
    }
};

// IntegratorOptions (options.rs:42-186); durations in integer nanoseconds
struct IntegratorOptions {
    int64_t init_step = 60 * NS_PER_S, min_step = NS_PER_S / 1000, max_step = 2700 * NS_PER_S;
    double tolerance = 1e-12;
    int32_t attempts = 50;
    bool fixed_step = false;
    ErrorControl error_ctrl = ErrorControl::RSSCartesianStep;
    static IntegratorOptions default_() { return {}; }
    static IntegratorOptions with_adaptive_step(int64_t min_step, int64_t max_step, double tol, ErrorControl ctrl) {
        return {max_step, min_step, max_step, tol, 50, false, ctrl};  // options.rs:66-82
    }
    static IntegratorOptions with_adaptive_step_s(double mn, double mx, double tol, ErrorControl ctrl) {
        return with_adaptive_step(seconds(mn), seconds(mx), tol, ctrl);
    }
    static IntegratorOptions with_fixed_step(int64_t step) {
        return {step, step, step, 0.0, 0, true, ErrorControl::RSSCartesianStep};  // options.rs:100-111
    }
    static IntegratorOptions with_fixed_step_s(double s) { return with_fixed_step(seconds(s)); }
    static IntegratorOptions with_tolerance(double tol) { IntegratorOptions o; o.tolerance = tol; return o; }
    void set_max_step(int64_t m) { if (init_step > m) init_step = m; max_step = m; }
    void set_min_step(int64_t m) { if (init_step < m) init_step = m; min_step = m; }
};

// Subset of anise's Frame used on the path + the explicit orientation model of nyxb.h
struct Frame {
    int32_t ephemeris_id = 399;
    double mu_km3_s2 = 398600.435436096;
    double mean_equatorial_radius_km = 6378.14;
    nyxb_rotation rotation{};  // kind 0: inertial axes
    double polar_radius_km = 0.0;  // > 0: ellipsoid for geodetic coordinates (ground stations); 0: sphere
    Frame with_mu_km3_s2(double mu) const { Frame f = *this; f.mu_km3_s2 = mu; return f; }
};
inline Frame EARTH_J2000() { return {}; }
inline Frame IAU_EARTH() {
    Frame f;
    f.rotation = nyxb_rotation{1, 0, 0.0, -0.641, 90.0, -0.557, 190.147, 360.9856235};  // pck00008
    f.polar_radius_km = 6356.75;
    return f;
}

// Spacecraft (cosmic/spacecraft.rs:115-143) without thruster / guidance / STM members
struct Spacecraft {
    double x_km = 0, y_km = 0, z_km = 0, vx_km_s = 0, vy_km_s = 0, vz_km_s = 0;
    int64_t epoch_ns = 0;
    Frame frame;
    double dry_mass_kg = 0, prop_mass_kg = 0, extra_mass_kg = 0;
    double srp_area_m2 = 0, coeff_reflectivity = 1.8, drag_area_m2 = 0, coeff_drag = 2.2;  // anise defaults
    static Spacecraft cartesian(double x, double y, double z, double vx, double vy, double vz, int64_t epoch_ns, const Frame& f) {
        Spacecraft s; s.x_km = x; s.y_km = y; s.z_km = z; s.vx_km_s = vx; s.vy_km_s = vy; s.vz_km_s = vz; s.epoch_ns = epoch_ns; s.frame = f;
        return s;
    }
    int64_t epoch() const { return epoch_ns; }
};

// Ephemeris container standing in for anise's Almanac (see nyxb_body)
struct BodyEphemeris {
    int32_t ephemeris_id; double mu_km3_s2, radius_km; int64_t t0_ns, interval_ns; int32_t n_coeffs;
    std::vector<double> coeffs;  // [n_intervals][3][n_coeffs]
};
struct Almanac { std::vector<BodyEphemeris> bodies; };

struct GravityFieldData {  // io/gravity.rs:90-128
    int32_t degree = 0, order = 0;
    std::vector<double> c_nm, s_nm;  // row-major (degree+1)^2
    Frame frame;
    static GravityFieldData from_j2(double j2, const Frame& frame) {
        GravityFieldData g; g.degree = 2; g.order = 0; g.c_nm.assign(9, 0.0); g.s_nm.assign(9, 0.0); g.c_nm[2 * 3 + 0] = j2; g.frame = frame;
        return g;
    }
};
struct PointMasses { std::vector<int32_t> celestial_objects; };
struct GravityField { GravityFieldData grav_data; };
struct SolarPressure { double phi = 1367.0; int32_t light_source = 10; std::vector<int32_t> shadow_bodies; bool estimate = true; /* solarpressure.rs:47-48, 88-92 */ };
struct Drag { int32_t density = NYXB_DENSITY_EXPONENTIAL; double rho0 = 3.614e-13, r0 = 700000.0, ref_alt_m = 88667.0; Frame frame = IAU_EARTH(); };

struct OrbitalDynamics {
    std::optional<PointMasses> point_masses_;
    std::optional<GravityField> gravity_;
    static OrbitalDynamics two_body() { return {}; }
    static OrbitalDynamics point_masses(std::vector<int32_t> objs) { OrbitalDynamics o; o.point_masses_ = PointMasses{std::move(objs)}; return o; }
    static OrbitalDynamics from_model(GravityField g) { OrbitalDynamics o; o.gravity_ = std::move(g); return o; }
};
struct SpacecraftDynamics {
    OrbitalDynamics orbital_dyn;
    std::optional<SolarPressure> srp;
    std::optional<Drag> drag;
    static SpacecraftDynamics new_(OrbitalDynamics o) { SpacecraftDynamics d; d.orbital_dyn = std::move(o); return d; }
};

struct IntegrationDetails { int64_t step_ns = 0; double error = 0; int32_t attempts = 1; int64_t n_steps = 0, n_rejected = 0, n_rhs = 0; };

namespace detail {
struct EngineDeleter { void operator()(nyxb_engine* e) const { nyxb_engine_destroy(e); } };
using EnginePtr = std::unique_ptr<nyxb_engine, EngineDeleter>;

inline EnginePtr make_engine(const SpacecraftDynamics& dyn, const Frame& frame, const Almanac* almanac, IntegratorMethod method,
                             const IntegratorOptions& o, int32_t mode, int32_t device) {
    nyxb_integ_opts co{(int32_t)method, (int32_t)o.error_ctrl, o.init_step, o.min_step, o.max_step, o.tolerance, o.attempts, o.fixed_step ? 1 : 0};
    nyxb_dynamics d{};
    d.mu_central_km3_s2 = frame.mu_km3_s2;
    d.central_radius_km = frame.mean_equatorial_radius_km;
    std::vector<nyxb_body> bodies;
    auto body_index = [&](int32_t id) -> int32_t {
        if (almanac) for (size_t i = 0; i < almanac->bodies.size(); ++i) if (almanac->bodies[i].ephemeris_id == id) return (int32_t)i;
        throw std::runtime_error("planetary data from third body not loaded");
    };
    if (almanac) for (auto& b : almanac->bodies)
        bodies.push_back(nyxb_body{b.mu_km3_s2, b.radius_km, b.t0_ns, b.interval_ns, (int32_t)(b.coeffs.size() / (3 * b.n_coeffs)), b.n_coeffs, b.coeffs.data()});
    d.n_bodies = (int32_t)bodies.size();
    d.bodies = bodies.data();
    if (dyn.orbital_dyn.point_masses_)
        for (int32_t id : dyn.orbital_dyn.point_masses_->celestial_objects) {   // summation order of orbital.rs:217
            if (id == frame.ephemeris_id) continue;                              // orbital.rs:219-222
            const int32_t j = body_index(id);
            if (!((d.point_mass_mask >> j) & 1u)) d.point_mass_order[d.n_point_masses++] = j;
            d.point_mass_mask |= 1u << j;
        }
    nyxb_gravity_field g{};
    if (dyn.orbital_dyn.gravity_) {
        const auto& gd = dyn.orbital_dyn.gravity_->grav_data;
        // a field of another body than the integration centre is evaluated about that body (gravity_field.rs:149-154)
        const int32_t gbody = gd.frame.ephemeris_id == frame.ephemeris_id ? NYXB_CENTRAL_BODY : body_index(gd.frame.ephemeris_id);
        g = nyxb_gravity_field{gd.degree, gd.order, gd.frame.mu_km3_s2, gd.frame.mean_equatorial_radius_km, gd.c_nm.data(), gd.s_nm.data(), gd.frame.rotation, gbody, 0};
        d.gravity = &g;
        d.n_gravity = 1;
    }
    nyxb_srp s{};
    if (dyn.srp) {
        s.phi_w_m2 = dyn.srp->phi; s.sun_body = body_index(dyn.srp->light_source); s.n_shadow = (int32_t)dyn.srp->shadow_bodies.size(); s.estimate = dyn.srp->estimate ? 1 : 0;
        for (int q = 0; q < s.n_shadow && q < 4; ++q)
            s.shadow_body[q] = dyn.srp->shadow_bodies[q] == frame.ephemeris_id ? NYXB_CENTRAL_BODY : body_index(dyn.srp->shadow_bodies[q]);
        d.srp = &s;
    }
    nyxb_drag dr{};
    if (dyn.drag) {
        dr = nyxb_drag{dyn.drag->density, 0, dyn.drag->rho0, dyn.drag->r0, dyn.drag->ref_alt_m, dyn.drag->frame.mean_equatorial_radius_km, dyn.drag->frame.rotation};
        d.drag = &dr;
    }
    nyxb_engine* e = nyxb_engine_create(&d, &co, mode, device);
    if (!e) throw std::runtime_error(std::string("nyxb_engine_create: ") + nyxb_last_error());
    return EnginePtr(e);
}

struct Soa {
    std::vector<double> state, consts; std::vector<int64_t> epoch;
    explicit Soa(const std::vector<Spacecraft>& v) : state(9 * v.size()), consts(4 * v.size()), epoch(v.size()) {
        const size_t n = v.size();
        for (size_t i = 0; i < n; ++i) {
            const Spacecraft& s = v[i];
            const double y[9] = {s.x_km, s.y_km, s.z_km, s.vx_km_s, s.vy_km_s, s.vz_km_s, s.coeff_reflectivity, s.coeff_drag, s.prop_mass_kg};
            for (int e = 0; e < 9; ++e) state[e * n + i] = y[e];  // cosmic/spacecraft.rs:449-473
            consts[i] = s.dry_mass_kg; consts[n + i] = s.extra_mass_kg; consts[2 * n + i] = s.srp_area_m2; consts[3 * n + i] = s.drag_area_m2;
            epoch[i] = s.epoch_ns;
        }
    }
};
inline Spacecraft unpack(const Spacecraft& tmpl, const std::vector<double>& out, const std::vector<int64_t>& ep, size_t n, size_t i) {
    Spacecraft s = tmpl;
    s.x_km = out[i]; s.y_km = out[n + i]; s.z_km = out[2 * n + i]; s.vx_km_s = out[3 * n + i]; s.vy_km_s = out[4 * n + i]; s.vz_km_s = out[5 * n + i];
    s.coeff_reflectivity = out[6 * n + i]; s.coeff_drag = out[7 * n + i]; s.prop_mass_kg = out[8 * n + i]; s.epoch_ns = ep[i];
    return s;
}
}  // namespace detail

class PropInstance;

// Propagator<SpacecraftDynamics> (propagator.rs:34-118)
class Propagator {
  public:
    SpacecraftDynamics dynamics;
    IntegratorOptions opts;
    IntegratorMethod method = IntegratorMethod::RungeKutta89;
    int32_t mode = NYXB_MODE_STRICT, device = 0;
    int32_t kernel = NYXB_KERNEL_AUTO;   // nyxb_engine_set_kernel: AUTO = the library's dispatch (by mode, field degree, ensemble size)
    static Propagator new_(SpacecraftDynamics d, IntegratorMethod m, IntegratorOptions o) { Propagator p; p.dynamics = std::move(d); p.method = m; p.opts = o; return p; }
    static Propagator rk89(SpacecraftDynamics d, IntegratorOptions o) { return new_(std::move(d), IntegratorMethod::RungeKutta89, o); }
    static Propagator dp78(SpacecraftDynamics d, IntegratorOptions o) { return new_(std::move(d), IntegratorMethod::DormandPrince78, o); }
    static Propagator default_(SpacecraftDynamics d) { return rk89(std::move(d), IntegratorOptions::default_()); }
    Propagator& with_mode(int32_t m) { mode = m; return *this; }
    inline PropInstance with(const Spacecraft& state, const Almanac* almanac = nullptr) const;

    struct BatchResult { std::vector<double> state; std::vector<int64_t> epoch; std::vector<nyxb_details> details; std::vector<int32_t> status; };
    // the rayon fan-out of mc/montecarlo.rs:233-253 as ONE batched call
    BatchResult propagate_batch(const std::vector<Spacecraft>& v, int64_t end_epoch_ns, const Almanac* almanac = nullptr,
                                std::vector<int64_t>* step_io = nullptr) const {
        BatchResult r;
        const size_t n = v.size();
        r.state.resize(9 * n); r.epoch.resize(n); r.details.resize(n); r.status.resize(n);
        if (n == 0) return r;
        auto eng = detail::make_engine(dynamics, v[0].frame, almanac, method, opts, mode, device);
        if (kernel != NYXB_KERNEL_AUTO && nyxb_engine_set_kernel(eng.get(), kernel) != NYXB_RC_OK)
            throw std::runtime_error(std::string("nyxb_engine_set_kernel: ") + nyxb_last_error());
        detail::Soa soa(v);
        int32_t rc = nyxb_propagate_batch(eng.get(), n, soa.state.data(), soa.consts.data(), soa.epoch.data(), end_epoch_ns,
                                          step_io ? step_io->data() : nullptr, r.state.data(), r.epoch.data(), r.details.data(), r.status.data());
        if (rc != NYXB_RC_OK) throw std::runtime_error(std::string("nyxb_propagate_batch: ") + nyxb_last_error());
        return r;
    }
    // the same fan-out over several GPUs of this process: one engine per device, contiguous run-index shards, results copied
    // straight into the caller's arrays (nyxb_propagate_batch_multi; mc/montecarlo.rs:188-203 on G GPUs is ONE call of it)
    BatchResult propagate_batch_multi(const std::vector<Spacecraft>& v, int64_t end_epoch_ns, const std::vector<int32_t>& devices,
                                      const Almanac* almanac = nullptr, std::vector<int64_t>* step_io = nullptr) const {
        BatchResult r;
        const size_t n = v.size();
        r.state.resize(9 * n); r.epoch.resize(n); r.details.resize(n); r.status.resize(n);
        if (n == 0) return r;
        if (devices.empty()) throw std::runtime_error("propagate_batch_multi: no devices");
        std::vector<detail::EnginePtr> engs;
        std::vector<nyxb_engine*> raw;
        for (int32_t d : devices) {
            engs.push_back(detail::make_engine(dynamics, v[0].frame, almanac, method, opts, mode, d));
            if (kernel != NYXB_KERNEL_AUTO && nyxb_engine_set_kernel(engs.back().get(), kernel) != NYXB_RC_OK)
                throw std::runtime_error(std::string("nyxb_engine_set_kernel: ") + nyxb_last_error());
            raw.push_back(engs.back().get());
        }
        detail::Soa soa(v);
        int32_t rc = nyxb_propagate_batch_multi(raw.data(), (int32_t)raw.size(), n, soa.state.data(), soa.consts.data(), soa.epoch.data(),
                                                end_epoch_ns, step_io ? step_io->data() : nullptr, r.state.data(), r.epoch.data(),
                                                r.details.data(), r.status.data());
        if (rc != NYXB_RC_OK) throw std::runtime_error(std::string("nyxb_propagate_batch_multi: ") + nyxb_last_error());
        return r;
    }
    // Spacecraft::with_stm() + until_epoch for a batch: final states and the 9x9 STMs (column-major per trajectory, [81][n])
    struct StmResult { std::vector<double> state, stm; std::vector<int64_t> epoch; std::vector<nyxb_details> details; std::vector<int32_t> status;
                       double phi(size_t n, size_t i, int r, int c) const { return stm[(size_t)(c * 9 + r) * n + i]; } };
    StmResult propagate_batch_stm(const std::vector<Spacecraft>& v, int64_t end_epoch_ns, const Almanac* almanac = nullptr,
                                  const std::vector<double>* stm_in = nullptr) const {
        StmResult r;
        const size_t n = v.size();
        r.state.resize(9 * n); r.stm.resize(81 * n); r.epoch.resize(n); r.details.resize(n); r.status.resize(n);
        if (n == 0) return r;
        auto eng = detail::make_engine(dynamics, v[0].frame, almanac, method, opts, mode, device);
        detail::Soa soa(v);
        int32_t rc = nyxb_propagate_batch_stm(eng.get(), n, soa.state.data(), soa.consts.data(), soa.epoch.data(), end_epoch_ns, nullptr,
                                              stm_in ? stm_in->data() : nullptr, r.state.data(), r.epoch.data(), r.stm.data(),
                                              r.details.data(), r.status.data());
        if (rc != NYXB_RC_OK) throw std::runtime_error(std::string("nyxb_propagate_batch_stm: ") + nyxb_last_error());
        return r;
    }
    // until_epoch_with_traj for a batch (instance.rs:297-340): final states + every accepted step of every trajectory in the
    // step-major SoA sink; `every(queries)` = Traj::at (md/trajectory/traj.rs:83-126) for all trajectories and query epochs in
    // one launch on the recording still resident on the device (nyxb_traj_resample).
    struct Resampled {
        size_t m = 0, n = 0; std::vector<double> state; std::vector<int32_t> status;   // [6][m][n], [m][n]
        double at(int c, size_t j, size_t i) const { return state[((size_t)c * m + j) * n + i]; }
        bool ok(size_t j, size_t i) const { return status[j * n + i] == NYXB_TRAJ_OK; }
    };
    struct TrajBatch : BatchResult {
        int64_t capacity = 0; size_t n = 0;
        std::vector<int64_t> t_epoch, t_count; std::vector<double> t_state;   // [cap][n], [n], [6][cap][n]
        std::shared_ptr<nyxb_engine> engine;
        int64_t epoch_at(size_t s, size_t i) const { return t_epoch[s * n + i]; }
        double state_at(int c, size_t s, size_t i) const { return t_state[((size_t)c * capacity + s) * n + i]; }
        Resampled every(const std::vector<int64_t>& query_epoch_ns) const {
            Resampled r; r.m = query_epoch_ns.size(); r.n = n;
            r.state.resize(6 * r.m * n); r.status.resize(r.m * n);
            int32_t rc = nyxb_traj_resample(engine.get(), n, nullptr, r.m, query_epoch_ns.data(), r.state.data(), r.status.data());
            if (rc != NYXB_RC_OK) throw std::runtime_error(std::string("nyxb_traj_resample: ") + nyxb_last_error());
            return r;
        }
        // the Brent search of until_nth_event (event.rs:186-211) inside every trajectory's last recorded step, one launch
        struct Located { std::vector<int64_t> epoch; std::vector<double> state; std::vector<int32_t> status; };   // [n], [6][n], [n]
        Located locate(int32_t event_kind, double value, int64_t epoch_precision_ns = 1000000) const {
            Located r; r.epoch.resize(n); r.state.resize(6 * n); r.status.resize(n);
            int32_t rc = nyxb_event_locate(engine.get(), n, nullptr, event_kind, value, epoch_precision_ns, status.data(), r.epoch.data(),
                                           r.state.data(), r.status.data());
            if (rc != NYXB_RC_OK) throw std::runtime_error(std::string("nyxb_event_locate: ") + nyxb_last_error());
            return r;
        }
    };
    // `event` != nullptr: stop every run at the end of the step in which the event scalar crossed zero for the trigger-th time
    TrajBatch propagate_batch_traj(const std::vector<Spacecraft>& v, int64_t end_epoch_ns, int64_t capacity, const Almanac* almanac = nullptr,
                                   nyxb_event* event = nullptr) const {
        TrajBatch r;
        const size_t n = v.size();
        r.n = n; r.capacity = capacity;
        r.state.resize(9 * n); r.epoch.resize(n); r.details.resize(n); r.status.resize(n);
        r.t_epoch.assign((size_t)capacity * n, 0); r.t_state.assign((size_t)6 * capacity * n, 0.0); r.t_count.assign(n, 0);
        if (n == 0) return r;
        r.engine = std::shared_ptr<nyxb_engine>(detail::make_engine(dynamics, v[0].frame, almanac, method, opts, mode, device).release(), nyxb_engine_destroy);
        detail::Soa soa(v);
        nyxb_traj_sink sink{capacity, r.t_epoch.data(), r.t_state.data(), r.t_count.data()};
        int32_t rc = nyxb_propagate_batch_event(r.engine.get(), n, soa.state.data(), soa.consts.data(), soa.epoch.data(), end_epoch_ns, nullptr,
                                                r.state.data(), r.epoch.data(), r.details.data(), r.status.data(), &sink, event);
        if (rc != NYXB_RC_OK) throw std::runtime_error(std::string("nyxb_propagate_batch_event: ") + nyxb_last_error());
        return r;
    }
    // nyx-py Propagator.many_until_epoch (py_md.rs:224-271): failed runs are dropped
    std::vector<Spacecraft> many_until_epoch(const std::vector<Spacecraft>& v, int64_t end_epoch_ns, const Almanac* almanac = nullptr) const {
        auto r = propagate_batch(v, end_epoch_ns, almanac);
        std::vector<Spacecraft> out;
        for (size_t i = 0; i < v.size(); ++i) if ((r.status[i] & 0xff) == 0) out.push_back(detail::unpack(v[i], r.state, r.epoch, v.size(), i));
        return out;
    }
};

// PropInstance (instance.rs:41-499): keeps the adapted step between calls
class PropInstance {
  public:
    Spacecraft state;
    IntegrationDetails details;
    PropInstance(const Propagator& p, const Spacecraft& s, const Almanac* a) : state(s), prop_(p), almanac_(a), step_{p.opts.init_step} {
        details.step_ns = p.opts.init_step;
    }
    Spacecraft for_duration(int64_t duration_ns) { return until_epoch(state.epoch_ns + duration_ns); }
    Spacecraft until_epoch(int64_t end_ns) {
        auto r = prop_.propagate_batch({state}, end_ns, almanac_, &step_);
        if (r.status[0] & 0xff) throw PropagationError(r.status[0]);
        if (r.details[0].n_steps > 0) details = {r.details[0].step_ns, r.details[0].error, r.details[0].attempts, r.details[0].n_steps, r.details[0].n_rejected, r.details[0].n_rhs};
        state = detail::unpack(state, r.state, r.epoch, 1, 0);
        return state;
    }
    IntegrationDetails latest_details() const { return details; }
  private:
    Propagator prop_;
    const Almanac* almanac_;
    std::vector<int64_t> step_;
};
inline PropInstance Propagator::with(const Spacecraft& state, const Almanac* almanac) const { return PropInstance(*this, state, almanac); }

// MonteCarlo (mc/montecarlo.rs:48-327) with a diagonal Cartesian dispersion (MvnSpacecraft::from_spacecraft_cov with a diagonal covariance)
struct Run { size_t index; Spacecraft dispersed_state; std::variant<Spacecraft, PropagationError> result; };
struct Results { std::vector<Run> runs; std::string scenario; int64_t total_steps = 0; };
class MonteCarlo {
  public:
    Spacecraft nominal_state; double std_dev[9]; std::string scenario; uint64_t seed;
    MonteCarlo(Spacecraft nominal, const double (&sd)[9], std::string name, uint64_t seed_) : nominal_state(nominal), scenario(std::move(name)), seed(seed_) {
        for (int i = 0; i < 9; ++i) std_dev[i] = sd[i];
    }
    // mc/montecarlo.rs:277-296: one serial stream; `skip` discards the first draws
    std::vector<Spacecraft> generate_states(size_t skip, size_t num_runs) const {
        std::mt19937_64 rng(seed);
        std::normal_distribution<double> nrm(0.0, 1.0);
        std::vector<Spacecraft> out;
        for (size_t i = 0; i < skip + num_runs; ++i) {
            double z[9];
            for (double& v : z) v = nrm(rng);
            if (i < skip) continue;
            Spacecraft s = nominal_state;
            s.x_km += std_dev[0] * z[0]; s.y_km += std_dev[1] * z[1]; s.z_km += std_dev[2] * z[2];
            s.vx_km_s += std_dev[3] * z[3]; s.vy_km_s += std_dev[4] * z[4]; s.vz_km_s += std_dev[5] * z[5];
            s.coeff_reflectivity += std_dev[6] * z[6]; s.coeff_drag += std_dev[7] * z[7]; s.prop_mass_kg += std_dev[8] * z[8];
            out.push_back(s);
        }
        return out;
    }
    // the same dispersions drawn on the GPU (nyxb_mvn_sample: counter-based stream keyed by (seed, run index), shard-invariant)
    std::vector<Spacecraft> generate_states_on_device(size_t skip, size_t num_runs, int32_t device = 0) const {
        const Spacecraft& t = nominal_state;
        const double tmpl[9] = {t.x_km, t.y_km, t.z_km, t.vx_km_s, t.vy_km_s, t.vz_km_s, t.coeff_reflectivity, t.coeff_drag, t.prop_mass_kg};
        double L[81] = {0};
        for (int i = 0; i < 9; ++i) L[i * 9 + i] = std_dev[i];
        std::vector<double> st(9 * num_runs);
        if (nyxb_mvn_sample(device, seed, skip, num_runs, tmpl, nullptr, L, st.data(), nullptr) != NYXB_RC_OK)
            throw std::runtime_error(std::string("nyxb_mvn_sample: ") + nyxb_last_error());
        std::vector<int64_t> ep(num_runs, t.epoch_ns);
        std::vector<Spacecraft> out;
        for (size_t i = 0; i < num_runs; ++i) out.push_back(detail::unpack(t, st, ep, num_runs, i));
        return out;
    }
    Results run_until_epoch(const Propagator& prop, const Almanac* almanac, int64_t end_epoch_ns, size_t num_runs) const {
        return resume_run_until_epoch(prop, almanac, 0, end_epoch_ns, num_runs);
    }
    Results resume_run_until_epoch(const Propagator& prop, const Almanac* almanac, size_t skip, int64_t end_epoch_ns, size_t num_runs) const {
        auto init = generate_states(skip, num_runs);
        auto r = prop.propagate_batch(init, end_epoch_ns, almanac);
        Results res; res.scenario = scenario;
        for (size_t i = 0; i < num_runs; ++i) {
            res.total_steps += r.details[i].n_steps;
            if (r.status[i] & 0xff) res.runs.push_back(Run{i, init[i], PropagationError(r.status[i])});  // per-run error, never aborts (mc/results.rs:48-59)
            else res.runs.push_back(Run{i, init[i], detail::unpack(init[i], r.state, r.epoch, num_runs, i)});
        }
        return res;
    }
};


// ---------------------------------------------------------------------------------------------------------------------
// Orbit determination: n sequential Kalman filters over one tracking schedule in ONE launch.
// ---------------------------------------------------------------------------------------------------------------------
enum class MeasurementType : int32_t { Range = NYXB_MSR_RANGE, Doppler = NYXB_MSR_DOPPLER };          // od/msr/types.rs:31-45
enum class KalmanVariant : int32_t { ReferenceUpdate = NYXB_KF_REFERENCE_UPDATE, DeviationTracking = NYXB_KF_DEVIATION_TRACKING };
struct StochasticNoise { double sigma = 0.0, bias_constant = 0.0; double covariance() const { return sigma * sigma; } };
struct SigmaRejection { double num_sigmas = 3.0; };                                                        // process/rejectcrit.rs:35-46

// GroundStation (od/ground_station/mod.rs:47-75; builtin.rs:25-117), instantaneous Range + Doppler
struct GroundStation {
    std::string name;
    double latitude_deg = 0, longitude_deg = 0, height_km = 0, elevation_mask_deg = 0;
    Frame frame = IAU_EARTH();
    std::vector<MeasurementType> measurement_types{MeasurementType::Range, MeasurementType::Doppler};
    StochasticNoise range_noise_km{2e-3, 0.0}, doppler_noise_km_s{3e-6, 0.0};
    static GroundStation dss65_madrid(double mask, StochasticNoise r, StochasticNoise d) { return {"Madrid", 40.427222, 4.250556, 0.834939, mask, IAU_EARTH(), {MeasurementType::Range, MeasurementType::Doppler}, r, d}; }
    static GroundStation dss34_canberra(double mask, StochasticNoise r, StochasticNoise d) { return {"Canberra", -35.398333, 148.981944, 0.691750, mask, IAU_EARTH(), {MeasurementType::Range, MeasurementType::Doppler}, r, d}; }
    static GroundStation dss13_goldstone(double mask, StochasticNoise r, StochasticNoise d) { return {"Goldstone", 35.247164, 243.205, 1.07114904, mask, IAU_EARTH(), {MeasurementType::Range, MeasurementType::Doppler}, r, d}; }
    // geodetic -> body-fixed position and local zenith on the frame's ellipsoid (anise Orbit::try_latlongalt)
    void body_fixed(double pos[3], double up[3]) const {
        const double a = frame.mean_equatorial_radius_km, b = frame.polar_radius_km > 0 ? frame.polar_radius_km : a;
        const double e2 = 1.0 - (b * b) / (a * a), d2r = 3.14159265358979323846 / 180.0;
        const double sl = std::sin(latitude_deg * d2r), cl = std::cos(latitude_deg * d2r), so = std::sin(longitude_deg * d2r), co = std::cos(longitude_deg * d2r);
        const double nu = a / std::sqrt(1.0 - e2 * sl * sl);
        pos[0] = (nu + height_km) * cl * co; pos[1] = (nu + height_km) * cl * so; pos[2] = (nu * (1.0 - e2) + height_km) * sl;
        up[0] = cl * co; up[1] = cl * so; up[2] = sl;
    }
    nyxb_ground_station to_c(const Frame& integration_frame, int32_t body_index, double central_radius_km) const {
        nyxb_ground_station g{};
        body_fixed(g.pos_fixed_km, g.up_fixed);
        g.elevation_mask_deg = elevation_mask_deg; g.rot = frame.rotation;
        const bool same = frame.ephemeris_id == integration_frame.ephemeris_id;
        g.body = same ? NYXB_CENTRAL_BODY : body_index;
        g.body_radius_km = same ? -1.0 : central_radius_km;
        g.n_types = (int32_t)measurement_types.size();
        for (int q = 0; q < g.n_types && q < 2; ++q) {
            g.types[q] = (int32_t)measurement_types[q];
            const StochasticNoise& nz = measurement_types[q] == MeasurementType::Range ? range_noise_km : doppler_noise_km_s;
            g.noise_var[q] = nz.covariance(); g.bias[q] = nz.bias_constant;
        }
        return g;
    }
};

struct ProcessNoise3D {   // od/snc.rs:38-56, 118-134, 288-311
    double diag[3] = {0, 0, 0}; int64_t disable_time = 0; bool ric = false;
    static ProcessNoise3D from_diagonal(const double (&v)[3], int64_t disable, bool ric_frame = false) { ProcessNoise3D p; for (int i = 0; i < 3; ++i) p.diag[i] = v[i]; p.disable_time = disable; p.ric = ric_frame; return p; }
    static ProcessNoise3D from_velocity_km_s(const double (&v)[3], int64_t noise_duration, int64_t disable, bool ric_frame = false) {
        ProcessNoise3D p; for (int i = 0; i < 3; ++i) p.diag[i] = v[i] / ((double)noise_duration * 1e-9); p.disable_time = disable; p.ric = ric_frame; return p;
    }
};

struct KfEstimate {   // od/estimate/kfestimate.rs: nominal state + 9x9 covariance (row-major here)
    Spacecraft nominal_state; double covar[81] = {0};
    static KfEstimate from_diag(const Spacecraft& s, const double (&d)[9]) { KfEstimate e; e.nominal_state = s; for (int i = 0; i < 9; ++i) e.covar[i * 9 + i] = d[i]; return e; }
};

// one tracking schedule, n observation sets: obs[(k*2 + type)*n + i], NaN = type not in the measurement's data
struct TrackingDataArc { std::vector<int64_t> epoch_ns; std::vector<std::string> tracker; std::vector<double> obs; size_t n = 0; };

struct ODSolution {
    size_t n = 0, m = 0;
    std::vector<double> state, covar, state_dev, resid_ratio, prefit, postfit;   // [9][n], [81][n] (c*9+r), [9][n], [m][2][n] x3
    std::vector<int64_t> epoch; std::vector<int32_t> msr_flags, status; std::vector<nyxb_details> details;
    Spacecraft final_state(const Spacecraft& tmpl, size_t i) const { return detail::unpack(tmpl, state, epoch, n, i); }
};

// KalmanODProcess (od/process/{initializers.rs:60-113, mod.rs:128-497}); msr_size 2 = SpacecraftKalmanOD, 1 = SpacecraftKalmanScalarOD
class KalmanODProcess {
  public:
    Propagator prop; KalmanVariant variant; std::optional<SigmaRejection> sigma_reject; std::vector<GroundStation> devices;
    const Almanac* almanac = nullptr; std::optional<ProcessNoise3D> process_noise;
    int64_t max_step = 60 * NS_PER_S, epoch_precision = 1000; int32_t msr_size = 2;
    KalmanODProcess(Propagator p, KalmanVariant v, std::optional<SigmaRejection> rej, std::vector<GroundStation> dev, const Almanac* alm = nullptr, int32_t msr = 2)
        : prop(std::move(p)), variant(v), sigma_reject(rej), devices(std::move(dev)), almanac(alm), msr_size(msr) {}
    KalmanODProcess& with_process_noise(ProcessNoise3D snc) { process_noise = snc; return *this; }

    ODSolution process_arcs(const std::vector<KfEstimate>& initial, const TrackingDataArc& arc) const {
        const size_t n = initial.size(), m = arc.epoch_ns.size();
        if (arc.n != n || arc.obs.size() != m * 2 * n || arc.tracker.size() != m) throw std::runtime_error("arc shape does not match the filters");
        std::vector<Spacecraft> noms; for (auto& e : initial) noms.push_back(e.nominal_state);
        const Frame& frame = noms.at(0).frame;
        auto eng = detail::make_engine(prop.dynamics, frame, almanac, prop.method, prop.opts, prop.mode, prop.device);
        detail::Soa soa(noms);
        std::vector<double> cov0(81 * n);
        for (size_t i = 0; i < n; ++i) for (int r = 0; r < 9; ++r) for (int c = 0; c < 9; ++c) cov0[(size_t)(c * 9 + r) * n + i] = initial[i].covar[r * 9 + c];
        std::vector<nyxb_ground_station> st;
        for (auto& d : devices) {
            int32_t bi = NYXB_CENTRAL_BODY;
            if (d.frame.ephemeris_id != frame.ephemeris_id) {
                if (!almanac) throw std::runtime_error("an almanac with the station's body is needed");
                bi = -2;
                for (size_t j = 0; j < almanac->bodies.size(); ++j) if (almanac->bodies[j].ephemeris_id == d.frame.ephemeris_id) bi = (int32_t)j;
                if (bi == -2) throw std::runtime_error("no ephemeris loaded for the station's body");
            }
            st.push_back(d.to_c(frame, bi, frame.mean_equatorial_radius_km));
        }
        std::vector<int32_t> trk(m);
        for (size_t k = 0; k < m; ++k) { trk[k] = -1; for (size_t j = 0; j < devices.size(); ++j) if (devices[j].name == arc.tracker[k]) trk[k] = (int32_t)j; }
        nyxb_od_config cfg{};
        cfg.variant = (int32_t)variant; cfg.msr_size = msr_size; cfg.reject_num_sigmas = sigma_reject ? sigma_reject->num_sigmas : -1.0;
        cfg.max_step_ns = max_step; cfg.epoch_precision_ns = epoch_precision;
        if (process_noise) { cfg.snc_enabled = 1; cfg.snc_frame = process_noise->ric ? 1 : 0; for (int i = 0; i < 3; ++i) cfg.snc_diag[i] = process_noise->diag[i]; cfg.snc_disable_time_ns = process_noise->disable_time; }
        nyxb_tracking_arc carc{(int64_t)m, arc.epoch_ns.data(), trk.data(), arc.obs.data()};
        ODSolution s; s.n = n; s.m = m;
        s.state.resize(9 * n); s.epoch.resize(n); s.covar.resize(81 * n); s.state_dev.resize(9 * n); s.resid_ratio.resize(m * 2 * n); s.prefit.resize(m * 2 * n);
        s.postfit.resize(m * 2 * n); s.msr_flags.resize(m * n); s.details.resize(n); s.status.resize(n);
        nyxb_od_outputs out{s.state.data(), s.epoch.data(), s.covar.data(), s.state_dev.data(), s.resid_ratio.data(), s.prefit.data(), s.postfit.data(),
                            s.msr_flags.data(), nullptr, nullptr, s.details.data(), s.status.data()};
        if (nyxb_od_ekf_batch(eng.get(), &cfg, (int32_t)st.size(), st.data(), &carc, n, soa.state.data(), soa.consts.data(), soa.epoch.data(), cov0.data(), &out) != NYXB_RC_OK)
            throw std::runtime_error(std::string("nyxb_od_ekf_batch: ") + nyxb_last_error());
        return s;
    }
};

}  // namespace nyxb
