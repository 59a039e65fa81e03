/*
 * nyxb.h — C ABI of the B200 batched orbit-propagation engine (libnyxb.so).
 *
 * This is the drop-in boundary for ONE path of nyx-space/nyx (reference paths are
 * relative to /root/reference/nyx-core/src):
 *
 *   MonteCarlo::run_until_epoch            mc/montecarlo.rs:188-273
 *     -> Propagator::with(..)              propagators/propagator.rs:88-108
 *     -> PropInstance::until_epoch         propagators/instance.rs:279-282
 *     -> PropInstance::propagate           propagators/instance.rs:87-262
 *     -> PropInstance::derive              propagators/instance.rs:358-493
 *     -> SpacecraftDynamics::eom           dynamics/spacecraft.rs:191-310
 *
 * plus the "next" rows of SURVEY.md §8(f) that sit on the same path: trajectory recording, event-terminated runs,
 * STM propagation and the sequential Kalman filter of od/process (nyxb_propagate_batch_stm, nyxb_od_ekf_batch), and
 * on-device dispersions (nyxb_mvn_sample).
 *
 * A Rust shim packs `Vec<Spacecraft>` into the SoA arrays
 * below, calls nyxb_propagate_batch through `extern "C"`, and unpacks the final
 * states into `Results` / `Vec<Spacecraft>`.  Nothing here mentions torch or CUDA
 * types: plain pointers, sizes and PODs only.  All pointers are HOST pointers
 * unless the function name ends in `_dev`.
 *
 * Units follow the reference: km, km/s, kg, m^2, seconds; time is integer
 * nanoseconds (hifitime::Duration / Epoch are integer-ns types).
 */
#ifndef NYXB_H
#define NYXB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NYXB_ABI_VERSION 4 /* 4: nyxb_engine_set_kernel / nyxb_engine_last_kernel / nyxb_engine_set_tx_tuning, nyxb_tx_table_dump, nyxb_propagate_batch_multi, nyxb_reference_normals;
                              nyxb_integ_opts.state_center, nyxb_gravity_field.body, nyxb_dynamics.n_gravity / n_point_masses / point_mass_order.  Earlier: 2: nyxb_srp gained `estimate`; STM, filter and dispersion entry points.  3: nyxb_traj_resample[_dev], nyxb_event_locate[_dev] */

/* ---- IntegratorMethod — propagators/rk_methods/mod.rs:65-79 (same order) ---- */
enum nyxb_method {
    NYXB_RK89 = 0,        /* RungeKutta89 (default)   rk_methods/rk.rs:89-252   */
    NYXB_DP78 = 1,        /* DormandPrince78          rk_methods/dormand.rs:71-184 */
    NYXB_DP45 = 2,        /* DormandPrince45          rk_methods/dormand.rs:23-69 */
    NYXB_RK4 = 3,         /* RungeKutta4 (fixed)      rk_methods/rk.rs:60-81    */
    NYXB_CK45 = 4,        /* CashKarp45               rk_methods/rk.rs:21-58    */
    NYXB_V56 = 5          /* Verner56                 rk_methods/verner.rs:24-79 */
};

/* ---- ErrorControl — propagators/error_ctrl.rs:30-71 (same order) ---- */
enum nyxb_error_ctrl {
    NYXB_RSS_CARTESIAN_STATE = 0,
    NYXB_RSS_CARTESIAN_STEP = 1, /* default */
    NYXB_RSS_STATE = 2,
    NYXB_RSS_STEP = 3,
    NYXB_LARGEST_ERROR = 4,
    NYXB_LARGEST_STATE = 5,
    NYXB_LARGEST_STEP = 6
};

/* ---- IntegratorOptions — propagators/options.rs:42-60 ----
 * Durations are hifitime integer nanoseconds.  `with_fixed_step` semantics
 * (options.rs:100-111): fixed_step=1, min=max=init=step, tolerance=0, attempts=0. */
typedef struct {
    int32_t method;      /* enum nyxb_method */
    int32_t error_ctrl;  /* enum nyxb_error_ctrl */
    int64_t init_step_ns;
    int64_t min_step_ns;
    int64_t max_step_ns;
    double tolerance;
    int32_t attempts;    /* u8 in the reference */
    int32_t fixed_step;  /* bool */
    /* IntegratorOptions.integration_frame (options.rs:60; instance.rs:117-142, 167-176, 211-220).  The engine's dynamics always
     * describe the INTEGRATION frame.  0: the states handed to nyxb_propagate_batch* are expressed in it (integration_frame = None, or
     * equal to the state's frame).  k + 1: the states are expressed relative to dynamics.bodies[k] (same inertial axes): they are
     * translated into the integration frame before the loop with that body's position and velocity at each trajectory's start
     * epoch (anise `transform_to`), and translated back at the epoch each run ends at.  Recorded trajectories and event scalars are
     * in the integration frame, as in the reference (the channel sends `self.state` from inside the loop). */
    int32_t state_center;
    int32_t _pad;
} nyxb_integ_opts;

/* ---- Body-fixed frame orientation (what anise `Almanac::rotate` supplies to
 * gravity_field.rs:150-154,258-265 and drag.rs:184-189).  anise and its PCK data
 * are absent from the reference tree, so the orientation is an explicit model:
 *   kind 0: body-fixed axes == inertial axes (no rotation)
 *   kind 1: IAU pole/prime-meridian, angles in degrees,
 *           ra = ra0 + ra1*T, dec = dec0 + dec1*T, W = w0 + w1*d,
 *           T = Julian centuries, d = days past the reference epoch (t = 0 ns),
 *           DCM inertial->fixed = R3(W) R1(90deg-dec) R3(90deg+ra). */
typedef struct {
    int32_t kind;
    int32_t _pad;
    double ra0_deg, ra1_deg_cy;
    double dec0_deg, dec1_deg_cy;
    double w0_deg, w1_deg_day;
} nyxb_rotation;

/* ---- GravityField — dynamics/gravity_field.rs:36-48 + io/gravity.rs:90-96 ----
 * c_nm/s_nm: normalised coefficients, row-major [(degree+1) x (degree+1)], entry
 * (n,m) at n*(degree+1)+m.  mu and r_eq are passed explicitly because the reference
 * takes them from the *frame* (gravity_field.rs:195-207), never from the file. */
typedef struct {
    int32_t degree;
    int32_t order;
    double mu_km3_s2;
    double r_eq_km;
    const double* c_nm;
    const double* s_nm;
    nyxb_rotation rot;
    /* The body this field belongs to: NYXB_CENTRAL_BODY (-1) = the integration-frame centre, else an index into
     * nyxb_dynamics.bodies — the state is translated to that body before the harmonic sum and the acceleration vector is rotated
     * back unchanged (gravity_field.rs:149-154, 258-267: "needed for multiple harmonic fields"). */
    int32_t body;
    int32_t _pad;
} nyxb_gravity_field;

/* ---- Ephemeris of one celestial body relative to the integration-frame centre,
 * inertial axes: piecewise Chebyshev position series (what anise evaluates from an
 * SPK for orbital.rs:230-234 / solarpressure.rs:138-143 / eclipse.rs:77).
 * coeffs layout: [n_intervals][3][n_coeffs]; interval i covers
 * [t0_ns + i*interval_ns, t0_ns + (i+1)*interval_ns). */
typedef struct {
    double mu_km3_s2;
    double radius_km;      /* mean equatorial radius (shadow computations) */
    int64_t t0_ns;
    int64_t interval_ns;
    int32_t n_intervals;
    int32_t n_coeffs;
    const double* coeffs;
} nyxb_body;

#define NYXB_MAX_BODIES 8
#define NYXB_CENTRAL_BODY (-1)
#define NYXB_MAX_FIELDS 3      /* harmonic fields per dynamics (OrbitalDynamics holds a Vec of accel models, orbital.rs:44-46) */

/* ---- SolarPressure + ShadowModel — dynamics/solarpressure.rs:43-49,135-165;
 * cosmic/eclipse.rs:35-83.  Body indices refer to nyxb_dynamics.bodies;
 * NYXB_CENTRAL_BODY designates the integration-frame centre. */
typedef struct {
    double phi_w_m2;        /* solar flux at 1 AU, default 1367 (solarpressure.rs:35) */
    int32_t sun_body;       /* index of the light source in bodies[] */
    int32_t n_shadow;
    int32_t shadow_body[4];
    int32_t estimate;       /* SolarPressure.estimate (solarpressure.rs:47-48, 131-133): Cr column of the STM A-matrix */
    int32_t _pad;
} nyxb_srp;

/* ---- Drag — dynamics/drag.rs:36-42,123-130,181-284 ---- */
enum nyxb_density {
    NYXB_DENSITY_CONSTANT = 0,   /* AtmDensity::Constant(rho) */
    NYXB_DENSITY_EXPONENTIAL = 1,/* AtmDensity::Exponential{rho0,r0,ref_alt_m} */
    NYXB_DENSITY_STDATM = 2      /* AtmDensity::StdAtm{max_alt_m} */
};
typedef struct {
    int32_t density;        /* enum nyxb_density */
    int32_t _pad;
    double rho0;            /* Constant: rho ; Exponential: rho0 */
    double r0;              /* Exponential */
    double ref_alt_m;       /* Exponential: ref_alt_m ; StdAtm: max_alt_m */
    double r_eq_km;         /* frame.mean_equatorial_radius_km() */
    nyxb_rotation rot;      /* drag frame orientation (IAU_EARTH in the reference) */
} nyxb_drag;

/* ---- SpacecraftDynamics — closed set of models the GPU path accepts
 * (dynamics/sequence/config.rs:96-169 serialises exactly this set). ---- */
typedef struct {
    double mu_central_km3_s2;     /* osc.frame.mu_km3_s2()  orbital.rs:86-90 */
    double central_radius_km;     /* used when the centre is a shadow body */
    int32_t n_bodies;
    int32_t n_gravity;            /* entries of `gravity` (0 with gravity != NULL is read as 1: ABI <= 3 callers) */
    const nyxb_body* bodies;      /* ephemerides available to the models */
    uint32_t point_mass_mask;     /* bit j set: bodies[j] acts as PointMasses member (orbital.rs:213-247) */
    int32_t n_point_masses;       /* > 0: `point_mass_order` lists the members in `celestial_objects` order (orbital.rs:217), which is
                                     the summation order STRICT mode reproduces; 0: ascending body index of the mask */
    const nyxb_gravity_field* gravity;  /* NULL: none; else n_gravity fields in accel-model order.  gravity[0] is the field the
                                           cooperative kernels split over lanes / warps; the others are summed per trajectory */
    const nyxb_srp* srp;                /* NULL: none */
    const nyxb_drag* drag;              /* NULL: none */
    int32_t point_mass_order[NYXB_MAX_BODIES];
} nyxb_dynamics;

/* ---- IntegrationDetails (propagators/mod.rs:49-56) + counters for the metric ---- */
typedef struct {
    int64_t step_ns;     /* details.step of the last step */
    double error;        /* details.error of the last adaptive step */
    int32_t attempts;    /* details.attempts of the last step */
    int32_t _pad;
    int64_t n_steps;     /* accepted steps incl. the final partial step */
    int64_t n_rejected;  /* rejected attempts */
    int64_t n_rhs;       /* RHS evaluations */
} nyxb_details;

/* ---- per-trajectory status: mirrors the reference's error variants ---- */
enum nyxb_status {
    NYXB_OK = 0,
    NYXB_ERR_PROP_MATH = 1,      /* PropagationError::PropMathError (NaN)  instance.rs:432-439 */
    NYXB_ERR_FUEL_EXHAUSTED = 2, /* DynamicsError::FuelExhausted           spacecraft.rs:163-168 */
    NYXB_ERR_MASSLESS = 3,       /* DynamicsError::MasslessSpacecraft      spacecraft.rs:201-203 */
    NYXB_ERR_EPHEMERIS = 4,      /* almanac error: epoch outside ephemeris coverage */
    NYXB_ERR_EVENT_NOT_FOUND = 5,/* PropagationError::NthEventError: end epoch reached first (event.rs:177-182) */
    NYXB_WARN_MAX_ATTEMPTS = 0x100 /* OR-ed flag: instance.rs:440-445 (warn only) */
};

/* ---- execution mode ---- */
enum nyxb_mode {
    NYXB_MODE_STRICT = 0, /* reference operation order, no FMA contraction: bit-parity mode */
    NYXB_MODE_FAST = 1    /* FMA + reordered cooperative harmonics: tolerance-parity mode */
};

/* ---- library-level return codes ---- */
enum nyxb_rc {
    NYXB_RC_OK = 0,
    NYXB_RC_BAD_ARG = -1,
    NYXB_RC_NO_DEVICE = -2,
    NYXB_RC_CUDA = -3,
    NYXB_RC_UNSUPPORTED = -4
};

typedef struct nyxb_engine nyxb_engine; /* opaque: device tables for one (dynamics, opts) pair */

/* Build device-resident tables (tableau, harmonic coefficients, ephemerides) for a
 * propagator setup == `Propagator::new(dynamics, method, opts)` (propagator.rs:55-61).
 * `device` is the CUDA ordinal.  Returns NULL on failure (see nyxb_last_error). */
nyxb_engine* nyxb_engine_create(const nyxb_dynamics* dyn, const nyxb_integ_opts* opts,
                                int32_t mode, int32_t device);
void nyxb_engine_destroy(nyxb_engine* eng);

/* Propagate n independent spacecraft until `end_epoch_ns`
 * == for each i: `prop.with(state_i, almanac).until_epoch(end_epoch)`
 * (mc/montecarlo.rs:233-253; nyx-py many_until_epoch py_md.rs:224-271).
 *
 *  state_soa   [9][n]  x,y,z,vx,vy,vz,Cr,Cd,prop_mass   (cosmic/spacecraft.rs:449-473)
 *  consts_soa  [4][n]  dry_mass_kg, extra_mass_kg, srp_area_m2, drag_area_m2
 *  epoch0_ns   [n]     start epochs (ns past the reference epoch of the ephemerides)
 *  step_ns     [n] or NULL: in/out adapted step of each PropInstance (instance.rs:56);
 *                      NULL => every run starts from opts.init_step_ns
 *  out_*       same layouts; out_epoch_ns [n]; details [n]; status [n]
 * Per-trajectory failures are reported in out_status and never abort the batch
 * (mc/results.rs:48-59).  Thread-safe for distinct engines. */
int32_t nyxb_propagate_batch(nyxb_engine* eng, size_t n,
                             const double* state_soa, const double* consts_soa,
                             const int64_t* epoch0_ns, int64_t end_epoch_ns,
                             int64_t* step_ns,
                             double* out_state_soa, int64_t* out_epoch_ns,
                             nyxb_details* out_details, int32_t* out_status);

/* Same, but every array is a DEVICE pointer on the engine's device and the launch
 * is enqueued on `cuda_stream` (a cudaStream_t passed as void*, NULL = default
 * stream) without synchronising.  Used by bench.py's HBM-resident `value` leg and
 * by the multi-GPU driver (final-state all-gather is stream-ordered after it). */
int32_t nyxb_propagate_batch_dev(nyxb_engine* eng, size_t n,
                                 const double* state_soa, const double* consts_soa,
                                 const int64_t* epoch0_ns, int64_t end_epoch_ns,
                                 int64_t* step_ns,
                                 double* out_state_soa, int64_t* out_epoch_ns,
                                 nyxb_details* out_details, int32_t* out_status,
                                 void* cuda_stream);

/* Multi-GPU fan-out of one ensemble behind the boundary (mc/montecarlo.rs:233-253: runs are independent, shards are contiguous
 * run-index ranges, nothing is exchanged while integrating).  `engines[g]`: one engine per device, all created from the same
 * (dynamics, options, mode); shard g = runs [g n / G, (g+1) n / G).  HOST arrays exactly as nyxb_propagate_batch; every device
 * integrates concurrently and its results land directly in the caller's [9][n] arrays — for a host caller this is the gather of
 * final states (device-resident callers launch nyxb_propagate_batch_dev per rank and all-gather, see nyx_b200/dist.py).
 * `mc.run_until_epoch(prop, almanac, end, num_runs)` on G GPUs is ONE call of this function. */
int32_t nyxb_propagate_batch_multi(nyxb_engine* const* engines, int32_t n_engines, size_t n,
                                   const double* state_soa, const double* consts_soa,
                                   const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                                   double* out_state_soa, int64_t* out_epoch_ns,
                                   nyxb_details* out_details, int32_t* out_status);

/* ---- Trajectory recording (next row (f)-1 of SURVEY.md §8): what `for_duration_with_traj` / `until_epoch_with_traj`
 * (propagators/instance.rs:297-340) collect through the mpsc channel (instance.rs:186-193, 255-259): the start state and
 * the state after every accepted step, final partial step included.  Record s of trajectory i lives at
 *   epoch_ns[s*n + i],  state[(c*capacity + s)*n + i]  (c = x,y,z,vx,vy,vz)
 * i.e. step-major SoA: trajectories that advance together write coalesced 56-byte-per-step streams.
 * count[i] = min(n_steps + 1, capacity); records beyond `capacity` are dropped (the final state is still returned). */
typedef struct {
    int64_t capacity;
    int64_t* epoch_ns;   /* [capacity][n] */
    double* state;       /* [6][capacity][n] */
    int64_t* count;      /* [n] */
} nyxb_traj_sink;

/* ---- Event-terminated propagation (next row (f)-3): the stop condition of `PropInstance::until_nth_event`
 * (propagators/event.rs:88-211, used by MonteCarlo::run_until_nth_event mc/montecarlo.rs:93-183).  After every accepted
 * NON-final step the event scalar minus `value` is evaluated; a sign change (y_prev * y_next < 0, event.rs:141-144) counts
 * one crossing; the run stops at the end of the step that brings the count to `trigger` (the state is returned and
 * recorded).  The root search inside that last step (Brent on the Hermite-interpolated trajectory, event.rs:186-196)
 * stays on the host.  Closed set of scalars (the reference's anise `ScalarExpr` is open-ended and not in the tree). */
enum nyxb_event_kind {
    NYXB_EVENT_NONE = 0,
    NYXB_EVENT_RMAG = 1,   /* |r| km            */
    NYXB_EVENT_RDOTV = 2,  /* r . v  km^2/s  (zero at the apsides) */
    NYXB_EVENT_X = 3, NYXB_EVENT_Y = 4, NYXB_EVENT_Z = 5,   /* Cartesian components, km (Z = 0: node crossing) */
    NYXB_EVENT_VMAG = 6    /* |v| km/s          */
};
typedef struct {
    int32_t kind;       /* enum nyxb_event_kind */
    int32_t trigger;    /* 1-based number of crossings to stop at */
    double value;       /* desired value: the monitored function is scalar - value */
    int32_t* crossings; /* [n] out: crossings seen; status NYXB_ERR_EVENT_NOT_FOUND when < trigger at the end epoch */
} nyxb_event;

/* nyxb_propagate_batch_traj + stop condition (HOST arrays; event == NULL: plain until-epoch propagation). */
int32_t nyxb_propagate_batch_event(nyxb_engine* eng, size_t n,
                                   const double* state_soa, const double* consts_soa,
                                   const int64_t* epoch0_ns, int64_t end_epoch_ns,
                                   int64_t* step_ns,
                                   double* out_state_soa, int64_t* out_epoch_ns,
                                   nyxb_details* out_details, int32_t* out_status,
                                   const nyxb_traj_sink* sink, const nyxb_event* event);

/* nyxb_propagate_batch + recording into `sink` (HOST arrays; NULL sink == nyxb_propagate_batch). */
int32_t nyxb_propagate_batch_traj(nyxb_engine* eng, size_t n,
                                  const double* state_soa, const double* consts_soa,
                                  const int64_t* epoch0_ns, int64_t end_epoch_ns,
                                  int64_t* step_ns,
                                  double* out_state_soa, int64_t* out_epoch_ns,
                                  nyxb_details* out_details, int32_t* out_status,
                                  const nyxb_traj_sink* sink);

/* Device-pointer variant: the sink's arrays are DEVICE pointers (the struct itself is read on the host). */
int32_t nyxb_propagate_batch_traj_dev(nyxb_engine* eng, size_t n,
                                      const double* state_soa, const double* consts_soa,
                                      const int64_t* epoch0_ns, int64_t end_epoch_ns,
                                      int64_t* step_ns,
                                      double* out_state_soa, int64_t* out_epoch_ns,
                                      nyxb_details* out_details, int32_t* out_status,
                                      const nyxb_traj_sink* sink, void* cuda_stream);

/* ---- Batched resampling of recorded trajectories on a common epoch grid (row (f)-1): what `Traj::every` /
 * `every_between` (md/trajectory/traj.rs:148-162, traj_it.rs:32-63) yield through `Traj::at` (traj.rs:83-126: exact hit,
 * else a window of 13 records around the query — 12 at the right edge, as coded — Hermite-interpolated in (r, v),
 * md/trajectory/interpolatable.rs:53-108) and what `Results::every_value_of[_between]` (mc/results.rs:88-160) iterate over,
 * for all n trajectories and all m query epochs in ONE launch.
 *  sink            the recording as filled by nyxb_propagate_batch_traj / _event (capacity, epoch_ns, state, count);
 *                  NULL (host variant only): the recording of this engine's last host-pointer propagation, still
 *                  resident in device memory (no upload);
 *  query_epoch_ns  [m];
 *  out_state       [6][m][n]  x,y,z,vx,vy,vz of trajectory i at query j at [(c*m + j)*n + i]; NaN where no data;
 *  out_status      [m][n]     NYXB_TRAJ_OK, or NYXB_TRAJ_NO_DATA when the query lies outside the trajectory's recorded
 *                             span (TrajError::NoInterpolationData) — a per-entry status, the batch never aborts.
 * Seconds are counted from the first record of the window (the reference passes absolute ET seconds to anise's
 * hermite_eval, which is not in the tree: parity unpinned at that boundary). */
enum nyxb_traj_status { NYXB_TRAJ_OK = 0, NYXB_TRAJ_NO_DATA = 1 };
int32_t nyxb_traj_resample(nyxb_engine* eng, size_t n, const nyxb_traj_sink* sink,
                           size_t m, const int64_t* query_epoch_ns, double* out_state, int32_t* out_status);
/* Device-pointer variant: the sink's arrays, the queries and the outputs are DEVICE pointers; stream-ordered. */
int32_t nyxb_traj_resample_dev(nyxb_engine* eng, size_t n, const nyxb_traj_sink* sink,
                               size_t m, const int64_t* query_epoch_ns, double* out_state, int32_t* out_status,
                               void* cuda_stream);

/* ---- Event location on the recorded trajectories (row (f)-3): the search `until_nth_event` runs after the propagation
 * stopped (propagators/event.rs:166-211: Brent's method on `event(traj.at(epoch))` between the last state on the channel and
 * the returned state), for all n runs in ONE launch.  The bracket is the last recorded step of each trajectory; the search
 * stops when the bracket is narrower than `epoch_precision_ns`; the state at the event epoch is the Hermite-interpolated one
 * (nyxb_traj_resample's arithmetic).
 *  sink               as for nyxb_traj_resample (NULL: the resident recording of the last host-pointer propagation);
 *  kind, value        the monitored scalar (enum nyxb_event_kind) and its desired value;
 *  run_status         [n] or NULL: out_status of the propagation — runs with an error code are skipped;
 *  out_event_epoch_ns [n], out_event_state [6][n] (NaN where not located),
 *  out_status         [n]: NYXB_TRAJ_OK; NYXB_TRAJ_NO_DATA (skipped run, fewer than two records);
 *                     NYXB_EVENT_NOT_BRACKETED (the scalar has the same sign at both ends of the last step). */
enum { NYXB_EVENT_NOT_BRACKETED = 2 };
int32_t nyxb_event_locate(nyxb_engine* eng, size_t n, const nyxb_traj_sink* sink, int32_t kind, double value,
                          int64_t epoch_precision_ns, const int32_t* run_status,
                          int64_t* out_event_epoch_ns, double* out_event_state, int32_t* out_status);
/* Device-pointer variant (sink arrays, run_status and outputs are DEVICE pointers; stream-ordered). */
int32_t nyxb_event_locate_dev(nyxb_engine* eng, size_t n, const nyxb_traj_sink* sink, int32_t kind, double value,
                              int64_t epoch_precision_ns, const int32_t* run_status,
                              int64_t* out_event_epoch_ns, double* out_event_state, int32_t* out_status, void* cuda_stream);

/* ---- State-transition-matrix propagation (next row (f)-2 of SURVEY.md §8): `Spacecraft::with_stm()` + propagate.
 * The integrated vector is the reference's 90-vector [x,y,z,vx,vy,vz,Cr,Cd,prop_mass, STM 9x9 column-major]
 * (cosmic/spacecraft.rs:449-473).  Stage derivative of the STM block AS CODED in the reference:
 * `ctx.stm * grad` with `ctx` = the state at the START of the step (dynamics/spacecraft.rs:203-227,
 * propagators/instance.rs:363-364), grad = the 9x9 A-matrix of `dual_eom` (spacecraft.rs:312-363;
 * orbital.rs:116-172, 249-307; gravity_field.rs:273-431; solarpressure.rs:167-233; Cr column when the SRP model
 * estimates it; Drag has no partials: `PartialsUndefined`, drag.rs:109-118, 286-295 -> NYXB_RC_UNSUPPORTED).
 * Only the Cartesian error controls (which look at r and v alone, error_ctrl.rs:89-122) and fixed steps are accepted.
 *  stm_in_soa / out_stm_soa  [81][n], entry (row r, col c) of trajectory i at [(c*9 + r)*n + i];
 *  stm_in_soa == NULL: identity (State::with_stm, cosmic/spacecraft.rs:433-440). */
int32_t nyxb_propagate_batch_stm(nyxb_engine* eng, size_t n,
                                 const double* state_soa, const double* consts_soa,
                                 const int64_t* epoch0_ns, int64_t end_epoch_ns,
                                 int64_t* step_ns, const double* stm_in_soa,
                                 double* out_state_soa, int64_t* out_epoch_ns, double* out_stm_soa,
                                 nyxb_details* out_details, int32_t* out_status);

/* ---- Sequential Kalman orbit determination over an ensemble (next row (f)-2; BASELINE configs[4]):
 * n independent `KalmanODProcess::process_arc` runs (od/process/mod.rs:128-497) — propagate the nominal state + STM
 * to each measurement, `KalmanFilter::time_update` / `measurement_update` (od/kalman/filtering.rs:59-316), state
 * replacement (EKF) and STM reset — in ONE kernel launch, one filter per trajectory. */
enum nyxb_msr_type { NYXB_MSR_RANGE = 0, NYXB_MSR_DOPPLER = 1 };  /* od/msr/types.rs:31-45 (the two-way capable ones) */

/* GroundStation (od/ground_station/mod.rs:47-75) reduced to what the filter needs.  The host converts latitude /
 * longitude / height into the body-fixed position and the local zenith (anise `Orbit::try_latlongalt`); the tracker's
 * inertial state is R^T p (+ w x r), translated by the ephemeris of `body` when the station does not sit on the
 * integration centre (trk_device.rs:150-152 `location`).  Instantaneous measurements only
 * (`integration_time: None`, trk_device.rs:154-200); light-time correction off. */
typedef struct {
    double pos_fixed_km[3];
    double up_fixed[3];          /* unit local zenith in the body-fixed frame (elevation = asin(rho_hat . up)) */
    double elevation_mask_deg;
    nyxb_rotation rot;           /* orientation of the station's body-fixed frame */
    int32_t body;                /* NYXB_CENTRAL_BODY or index into dynamics.bodies: the body the station sits on */
    int32_t n_types;             /* 1 or 2 */
    int32_t types[2];            /* enum nyxb_msr_type, in the device's IndexSet order */
    int32_t _pad;
    double noise_var[2];         /* TrackingDevice::measurement_covar per type (trk_device.rs:223-236) */
    double bias[2];              /* TrackingDevice::measurement_bias per type (trk_device.rs:238-253) */
    double body_radius_km;       /* radius of the body the spacecraft orbits when it can obstruct the line of sight
                                    (trk_device.rs:162-166); <= 0: no obstruction test */
} nyxb_ground_station;

enum nyxb_kf_variant { NYXB_KF_REFERENCE_UPDATE = 0 /* EKF */, NYXB_KF_DEVIATION_TRACKING = 1 /* CKF */ }; /* od/kalman/mod.rs */

typedef struct {
    int32_t variant;             /* enum nyxb_kf_variant */
    int32_t msr_size;            /* MsrSize::DIM: 2 = SpacecraftKalmanOD, 1 = SpacecraftKalmanScalarOD (od/mod.rs:77-91) */
    double reject_num_sigmas;    /* SigmaRejection.num_sigmas (process/rejectcrit.rs:35-46); < 0: None */
    int64_t max_step_ns;         /* KalmanODProcess.max_step, default 1 min (process/initializers.rs:66-75) */
    int64_t epoch_precision_ns;  /* default 1 us */
    /* one ProcessNoise3D (od/snc.rs:38-56), no decay / start time */
    int32_t snc_enabled;
    int32_t snc_frame;           /* 0: state frame, 1: LocalFrame::RIC (snc.rs:226-262) */
    double snc_diag[3];          /* km^2/s^4 */
    int64_t snc_disable_time_ns;
} nyxb_od_config;

/* Tracking arc shared by the ensemble (one schedule, n observation sets) and the per-measurement outputs.
 * obs[(k*2 + t)*n + i]: observation of type t (enum nyxb_msr_type) of trajectory i at measurement k; NaN = that
 * type is not in `msr.data` (both NaN: the measurement is not in trajectory i's arc at all). */
typedef struct {
    int64_t n_msr;
    const int64_t* epoch_ns;     /* [n_msr] ascending */
    const int32_t* tracker;      /* [n_msr] index into stations; < 0: unknown tracker (process/mod.rs:400-410) */
    const double* obs;           /* [n_msr][2][n] */
} nyxb_tracking_arc;

enum nyxb_msr_flag {
    NYXB_MSRF_PROCESSED = 1,     /* a measurement update ran (accepted or rejected by the sigma test) */
    NYXB_MSRF_REJECTED = 2,      /* residual ratio above num_sigmas: time update only (filtering.rs:169-184) */
    NYXB_MSRF_NOT_VISIBLE = 4,   /* device.measure() returned None: below the mask / obstructed (process/mod.rs:386-392) */
    NYXB_MSRF_ABSENT = 8         /* no data for this trajectory */
};

typedef struct {
    double* state_soa;           /* [9][n]  final nominal state (EKF: estimate) */
    int64_t* epoch_ns;           /* [n] */
    double* covar_soa;           /* [81][n] final covariance, (r,c) at [(c*9+r)*n + i] */
    double* state_dev_soa;       /* [9][n]  final state deviation (CKF); NULL to skip */
    /* per measurement k and residual window w (msr_size 2: one window holding both types; msr_size 1: one per type) */
    double* resid_ratio;         /* [n_msr][2][n] or NULL */
    double* prefit;              /* [n_msr][2][n] or NULL  (slot = position of the type in the device's list) */
    double* postfit;             /* [n_msr][2][n] or NULL */
    int32_t* msr_flags;          /* [n_msr][n]    or NULL */
    double* est_state;           /* [n_msr][9][n] or NULL: estimated state after measurement k */
    double* est_covar_diag;      /* [n_msr][9][n] or NULL */
    nyxb_details* details;       /* [n] or NULL: n_steps / n_rhs over the whole arc */
    int32_t* status;             /* [n] */
} nyxb_od_outputs;

/* state_soa/consts_soa/epoch0_ns as in nyxb_propagate_batch (initial_estimate.nominal_state);
 * covar0_soa [81][n] initial covariance (KfEstimate.covar); HOST pointers everywhere. */
int32_t nyxb_od_ekf_batch(nyxb_engine* eng, const nyxb_od_config* cfg,
                          int32_t n_stations, const nyxb_ground_station* stations,
                          const nyxb_tracking_arc* arc, size_t n,
                          const double* state_soa, const double* consts_soa, const int64_t* epoch0_ns,
                          const double* covar0_soa, const nyxb_od_outputs* out);

/* ---- On-device Monte Carlo dispersions (next row (f)-4): `MvnSpacecraft::sample` (mc/multivariate.rs:298-331) for the
 * runs [first_index, first_index + n): state_i = template + (sqrt_s_v * z_i + mean), z_i ~ N(0, I_9) drawn from a
 * counter-based stream keyed by (seed, run index) — Philox4x32-10 + Box-Muller, see nyx_b200/csrc/nyxb_mvn.cu — so that a
 * run's draw does not depend on how the ensemble is sharded (the reference's serial Pcg64Mcg + ziggurat stream,
 * mc/montecarlo.rs:277-296, is not reproduced: DESIGN.md §3).
 *  template_state[9], mean[9] (NULL = 0), sqrt_s_v[81] row-major (V * sqrt(S) of the covariance's SVD, multivariate.rs:237-247)
 *  out_state_soa [9][n], out_dispersion_soa [9][n] or NULL (the x_i themselves).
 * `_dev`: ONLY out_state_soa / out_dispersion_soa are device pointers (stream-ordered launch); template_state, mean and sqrt_s_v
 * are HOST pointers in both variants — they are read on the host into the kernel's parameter block. */
int32_t nyxb_mvn_sample(int32_t device, uint64_t seed, uint64_t first_index, size_t n,
                        const double* template_state, const double* mean, const double* sqrt_s_v,
                        double* out_state_soa, double* out_dispersion_soa);
int32_t nyxb_mvn_sample_dev(int32_t device, uint64_t seed, uint64_t first_index, size_t n,
                            const double* template_state, const double* mean, const double* sqrt_s_v,
                            double* out_state_soa, double* out_dispersion_soa, void* cuda_stream);

/* ---- The reference's own dispersion stream (row a2; host code, no device needed): `MonteCarlo::generate_states`
 * (mc/montecarlo.rs:277-296) = one serial `Pcg64Mcg::new(seed)` (seed: u128 = seed_hi * 2^64 + seed_lo) feeding rand_distr's ziggurat
 * `StandardNormal`, nine normals per run, the first `skip` runs dropped.  out_z [n][9] row-major: x_i = sqrt_s_v z_i + mean
 * (multivariate.rs:298-302).  Pcg64Mcg is pinned on the generator's official known-answer vector; the ziggurat tables are regenerated
 * from the published formulas (see nyx_b200/csrc/nyxb_rng.cu). */
int32_t nyxb_reference_normals(uint64_t seed_lo, uint64_t seed_hi, uint64_t skip, size_t n, double* out_z);
int32_t nyxb_pcg64mcg_u64(uint64_t seed_lo, uint64_t seed_hi, size_t n, uint64_t* out);   /* raw generator output (known-answer tests) */
int32_t nyxb_ziggurat_tables(double* x257, double* f257);                                  /* the layer tables (inspection) */

/* Tuning / introspection. */
/* Kernel families behind nyxb_propagate_batch*.  AUTO picks by mode, degree and ensemble size:
 *   THREAD      one thread per trajectory (no or low-degree gravity field; STRICT and FAST)
 *   COOP        8 / 16 / 32 lanes of a warp per trajectory, harmonic sum split by columns over the lanes (STRICT and FAST)
 *   TRANSPOSED  FAST only, degree 8..70: one CTA per set of 32 trajectories, lane = trajectory, warp = column position, persistent
 *               CTAs with (set, time-slice) tickets — the kernel of large harmonics-dominated ensembles (>= 1 024 trajectories, field of degree 8..70; smaller ensembles go to the lane-cooperative kernel) */
enum nyxb_kernel { NYXB_KERNEL_AUTO = 0, NYXB_KERNEL_THREAD = 1, NYXB_KERNEL_COOP = 2, NYXB_KERNEL_TRANSPOSED = 3 };
int32_t nyxb_engine_set_kernel(nyxb_engine* eng, int32_t kernel);          /* enum nyxb_kernel; NYXB_RC_UNSUPPORTED if the setup cannot use it */
int32_t nyxb_engine_last_kernel(const nyxb_engine* eng);                   /* family used by the last propagation launch */
/* TRANSPOSED kernel: step attempts per time slice (default 64) and an upper bound on the persistent CTAs (0 = SMs x occupancy).
 * Sets are only parked when there are more sets than CTAs. */
int32_t nyxb_engine_set_tx_tuning(nyxb_engine* eng, int32_t slice_attempts, int32_t max_ctas);
int32_t nyxb_engine_set_tx_positions(nyxb_engine* eng, int32_t positions);   /* walker warps per set: 0 = by degree, 8, 10, 16 */
int32_t nyxb_engine_set_lanes(nyxb_engine* eng, int32_t lanes_per_trajectory); /* 0 = auto */
int32_t nyxb_engine_get_lanes(const nyxb_engine* eng);
int64_t nyxb_engine_launch_count(const nyxb_engine* eng); /* kernels launched so far */
double nyxb_engine_last_kernel_ms(const nyxb_engine* eng); /* CUDA-event time of the last launch (host API only) */

/* Register-resident DFMA throughput probe: returns achieved FP64 TFLOP/s (FMA = 2 flop)
 * on `device`; the FP64 roof that bench.py reports against. */
double nyxb_measure_fp64_tflops(int32_t device, int32_t iters);

/* Host-only inspection of the cooperative kernel's coefficient table (no device needed): the column -> lane
 * schedule and the packed records for `lanes` in {8,16,32}.  Two-call pattern: with recs == NULL only the sizes are
 * returned.  Layouts: recs [(L+2)/2 pairs][5 pieces][lanes][2], col_start / col_m [lanes][kmax], colseed [N+2][4]
 * (see nyx_b200/csrc/nyxb_coop.h).  Used by the CPU tests to check the table algebra against a direct evaluation. */
int32_t nyxb_coop_table_dump(const nyxb_gravity_field* field, int32_t lanes, int32_t* out_L, int32_t* out_kmax,
                             double* recs, int32_t* col_start, int32_t* col_m, double* colseed);

/* Same for the transposed kernel (`positions` in {8,10,16}): recA [(n_rec+1)][4], recK [n_rec+2], colseed [N+2][4],
 * sched [positions][2 + 2*kmax] = {first record, columns, (m, entries) per column} (see nyx_b200/csrc/nyxb_tx.h). */
int32_t nyxb_tx_table_dump(const nyxb_gravity_field* field, int32_t positions, int32_t* out_n_rec, int32_t* out_kmax,
                           double* recA, double* recK, double* colseed, int32_t* sched);

int32_t nyxb_abi_version(void);
const char* nyxb_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* NYXB_H */
