// nyxb_api.cu — C-ABI implementation (include/nyxb.h): engine lifetime, table packing and
// upload, launches.  Host side of `Propagator::new` + `MonteCarlo::run_until_epoch`'s fan-out.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <map>

#include "nyxb_coop.h"
#include "nyxb_tx.h"
#include "nyxb_device.cuh"
#include "nyxb_od.cuh"
#include "nyxb_tableaux.h"

// kernels (nyxb_kernels.cu built twice, nyxb_coop.cu)
extern "C" cudaError_t nyxb_launch_thread_strict(const DevSetup*, size_t, const double*, const double*, const long long*,
                                                 long long, long long*, double*, long long*, nyxb_details*, int*, int,
                                                 const DevSink*, cudaStream_t);
extern "C" cudaError_t nyxb_launch_thread_fast(const DevSetup*, size_t, const double*, const double*, const long long*,
                                               long long, long long*, double*, long long*, nyxb_details*, int*, int,
                                               const DevSink*, cudaStream_t);
extern "C" double nyxb_fp64_probe(int device, int iters);
extern "C" cudaError_t nyxb_launch_frame_shift(const DevBody*, double, size_t, double*, const long long*, int*, cudaStream_t);
extern "C" cudaError_t nyxb_launch_od_coop(const DevSetup*, const DevOd*, const int*, size_t, const double*, const double*, const long long*,
                                           double*, long long*, nyxb_details*, int*, cudaStream_t);
extern "C" int nyxb_od_coop_kmax(void);
extern "C" cudaError_t nyxb_launch_traj_resample(long long, const long long*, const double*, const long long*, size_t, size_t,
                                                 const long long*, double*, int*, cudaStream_t);
extern "C" cudaError_t nyxb_launch_event_locate(long long, const long long*, const double*, const long long*, size_t, int, double, long long,
                                                const int*, long long*, double*, int*, cudaStream_t);
extern "C" cudaError_t nyxb_launch_mvn(unsigned long long, unsigned long long, size_t, const double*, const double*, const double*,
                                       double*, double*, cudaStream_t);

// layout of the PODs the host mirrors rely on (nyx_b200/abi.py, tests/test_abi.py)
static_assert(sizeof(nyxb_integ_opts) == 56 && sizeof(nyxb_gravity_field) == 104 && sizeof(nyxb_dynamics) == 96 && sizeof(nyxb_rotation) == 56 && sizeof(nyxb_srp) == 40 && sizeof(nyxb_details) == 48, "ABI layout");
static_assert(sizeof(nyxb_ground_station) == 176 && sizeof(nyxb_od_config) == 72 && sizeof(nyxb_tracking_arc) == 32 && sizeof(nyxb_od_outputs) == 96, "ABI layout");

static thread_local std::string g_err;
static void set_err(const std::string& s) { g_err = s; }
#define CUDA_TRY(x)                                                                                   \
    do {                                                                                              \
        cudaError_t _e = (x);                                                                         \
        if (_e != cudaSuccess) {                                                                      \
            set_err(std::string(#x) + ": " + cudaGetErrorString(_e));                                 \
            return NYXB_RC_CUDA;                                                                      \
        }                                                                                             \
    } while (0)

struct nyxb_engine {
    int device = 0;
    int mode = NYXB_MODE_STRICT;
    int lanes = 0;  // 0 = auto
    DevSetup S;
    std::vector<void*> dev_allocs;
    long long launches = 0;
    double last_ms = 0.0;
    // grow-only device staging slab of the host-pointer entry point (no cudaMalloc/cudaFree per call)
    size_t cap = 0;
    double* d_f64 = nullptr;
    long long* d_i64 = nullptr;
    nyxb_details* d_det = nullptr;
    int* d_status = nullptr;
    cudaStream_t stream = nullptr;
    // grow-only device buffers of the trajectory sink (host-pointer entry point)
    size_t sink_bytes = 0;
    unsigned char* d_sink = nullptr;
    size_t rec_n = 0;        // the recording resident in d_sink: trajectories and capacity (nyxb_traj_resample with sink == NULL)
    long long rec_cap = 0;
    std::vector<double> h_cnm, h_snm;      // host copies for building cooperative tables lazily
    std::map<int, DevCoop> coop;           // lanes -> device tables
    std::map<int, DevCoopStrict> scoop;    // lanes -> STRICT cooperative schedules
    int kernel = NYXB_KERNEL_AUTO;         // nyxb_engine_set_kernel
    int last_kernel = NYXB_KERNEL_AUTO;    // family the last launch used
    std::map<int, DevTx> tx;               // positions -> tables of the transposed kernel
    int tx_slice = 64;                     // step attempts per time slice of the persistent transposed kernel
    int tx_positions = 0;                  // walker warps per set (0: by degree)
    int tx_max_ctas = 0;                   // 0: every resident slot (SMs x occupancy); tests shrink it to force time slicing
    size_t txq_bytes = 0;                  // grow-only queue + parking workspace of the transposed kernel
    unsigned char* d_txq = nullptr;
    int sms = 0;
    size_t frame_n = 0;                    // grow-only copy of the input states translated into the integration frame
    double* d_frame = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ~nyxb_engine() {
        for (void* p : dev_allocs) cudaFree(p);
        cudaFree(d_f64); cudaFree(d_i64); cudaFree(d_det); cudaFree(d_status); cudaFree(d_sink); cudaFree(d_txq); cudaFree(d_frame);
        if (stream) cudaStreamDestroy(stream);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }
};

static bool tableau_for(int method, int& order, int& stages, const NyxbCoef*& a, const NyxbCoef*& b) {
    switch (method) {  // rk_methods/mod.rs:81-133
    case NYXB_RK89: order = 9; stages = 16; a = NYXB_RK89_A; b = NYXB_RK89_B; return true;
    case NYXB_DP78: order = 8; stages = 13; a = NYXB_DP78_A; b = NYXB_DP78_B; return true;
    case NYXB_DP45: order = 5; stages = 7; a = NYXB_DP45_A; b = NYXB_DP45_B; return true;
    case NYXB_RK4: order = 4; stages = 4; a = NYXB_RK4_A; b = NYXB_RK4_B; return true;
    case NYXB_CK45: order = 5; stages = 6; a = NYXB_CK45_A; b = NYXB_CK45_B; return true;
    case NYXB_V56: order = 6; stages = 8; a = NYXB_V56_A; b = NYXB_V56_B; return true;
    default: return false;
    }
}

static double host_dur_to_seconds(long long total_ns) {
    const long long NPC = 3155760000000000000LL, NPS = 1000000000LL;
    long long cent = total_ns / NPC;
    if (total_ns % NPC < 0) cent -= 1;
    long long nanos = total_ns - cent * NPC;
    volatile double s = (double)(nanos / NPS);
    volatile double f = (double)(nanos % NPS) * 1e-9;
    if (cent == 0) return s + f;
    volatile double c = (double)cent * 3155760000.0;
    volatile double cs = c + s;
    return cs + f;
}

static DevRotation pack_rot(const nyxb_rotation& r) {
    DevRotation d;
    d.kind = r.kind;
    d.ra0 = r.ra0_deg; d.ra1 = r.ra1_deg_cy; d.dec0 = r.dec0_deg; d.dec1 = r.dec1_deg_cy; d.w0 = r.w0_deg; d.w1 = r.w1_deg_day;
    volatile double w = r.w1_deg_day * 1.7453292519943295e-2;
    d.wdot = (r.kind == 0) ? 0.0 : w / 86400.0;
    d.ra_dot = r.ra1_deg_cy * 1.7453292519943295e-2 / (36525.0 * 86400.0);
    d.dec_dot = r.dec1_deg_cy * 1.7453292519943295e-2 / (36525.0 * 86400.0);
    return d;
}

template <typename T>
static T* upload(nyxb_engine* e, const T* host, size_t count) {
    T* d = nullptr;
    if (cudaMalloc(&d, sizeof(T) * (count ? count : 1)) != cudaSuccess) return nullptr;
    e->dev_allocs.push_back(d);
    if (count && cudaMemcpy(d, host, sizeof(T) * count, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    return d;
}

// GravityField::new (gravity_field.rs:52-92) for one field: recursion factors, diagonal, per-thread column-walk records
static bool build_grav(nyxb_engine* e, const nyxb_gravity_field& g, DevGrav& G, bool primary) {
        if (g.degree < 1 || g.degree > NYXB_MAX_DEGREE || g.order < 0 || g.order > g.degree) {
            set_err("gravity field degree/order out of range (1..96)"); return false;
        }
        const int N = g.degree, np2 = N + 2;
        G.N = N; G.M = g.order; G.mu = g.mu_km3_s2; G.r_eq = g.r_eq_km; G.inv_r_eq = 1.0 / g.r_eq_km;
        G.rot = pack_rot(g.rot);
        // GravityField::new gravity_field.rs:52-92 (same formulas; sqrt and / are correctly rounded)
        std::vector<double> adiag(N + 3), offd(N + 2);
        adiag[0] = 1.0;
        for (int n = 1; n <= np2; ++n) {
            double nf = (double)n;
            volatile double t = 1.0 + 1.0 / (2.0 * nf);
            volatile double s = std::sqrt(t);
            volatile double v = s * adiag[n - 1];
            adiag[n] = v;
        }
        for (int n = 0; n <= N + 1; ++n) offd[n] = std::sqrt(2.0 * (double)n + 3.0);
        std::vector<DevHarm> tab((size_t)(N + 2) * (N + 3) / 2);
        const double sqrt2 = std::sqrt(2.0);
        for (int n = 0; n <= N + 1; ++n) {
            for (int m = 0; m <= n; ++m) {
                double nf = (double)n, mf = (double)m;
                DevHarm h;
                volatile double cnum = (2.0 * nf + 1.0) * (nf + mf - 1.0) * (nf - mf - 1.0);
                volatile double cden = (nf - mf) * (nf + mf) * (2.0 * nf - 3.0);
                volatile double cq = cnum / cden;
                h.c = std::sqrt(cq);
                volatile double bnum = (2.0 * nf + 1.0) * (2.0 * nf - 1.0);
                volatile double bden = (nf + mf) * (nf - mf);
                volatile double bq = bnum / bden;
                h.b = std::sqrt(bq);
                volatile double v01 = (nf - mf) * (nf + mf + 1.0);
                h.vr01 = std::sqrt(v01);
                volatile double v11n = (2.0 * nf + 1.0) * (nf + mf + 2.0) * (nf + mf + 1.0);
                volatile double v11 = v11n / (2.0 * nf + 3.0);
                h.vr11 = std::sqrt(v11);
                if (m == 0) { h.vr01 = h.vr01 / sqrt2; h.vr11 = h.vr11 / sqrt2; }
                if (n <= N) { h.cbar = g.c_nm[(size_t)n * (N + 1) + m]; h.sbar = g.s_nm[(size_t)n * (N + 1) + m]; }
                else { h.cbar = 0.0; h.sbar = 0.0; }
                if (!(n >= m + 2)) { h.b = 0.0; h.c = 0.0; }  // never read; avoid NaN/inf noise
                tab[(size_t)n * (n + 1) / 2 + m] = h;
            }
        }
        if (primary) {
            e->h_cnm.assign(g.c_nm, g.c_nm + (size_t)(N + 1) * (N + 1));
            e->h_snm.assign(g.s_nm, g.s_nm + (size_t)(N + 1) * (N + 1));
        }
        G.tab = upload(e, tab.data(), tab.size());
        G.a_diag = upload(e, adiag.data(), adiag.size());
        G.offdiag = upload(e, offd.data(), offd.size());
        {   // records of the FAST per-thread column walk (grav_accel_cols, nyxb_device.cuh), in walk order
            const int ncols = std::min(N, g.order) + 1;
            auto T = [&](int n, int m) -> const DevHarm& { return tab[(size_t)n * (n + 1) / 2 + m]; };
            std::vector<double> cr;
            cr.reserve((size_t)ncols * (N + 2) * 8);
            const double inv_req = 1.0 / g.r_eq_km;
            auto push = [&](int k, int j) {
                double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                if (j <= N && k <= g.order) { v[0] = (double)k * sqrt2 * T(j, k).cbar * inv_req; v[1] = (double)k * sqrt2 * T(j, k).sbar * inv_req; }
                if (j <= N) { v[2] = sqrt2 * T(j, k - 1).vr01 * T(j, k - 1).cbar * inv_req; v[3] = sqrt2 * T(j, k - 1).vr01 * T(j, k - 1).sbar * inv_req; }
                if (j >= 2) { v[4] = sqrt2 * T(j - 1, k - 1).vr11 * T(j - 1, k - 1).cbar * inv_req; v[5] = sqrt2 * T(j - 1, k - 1).vr11 * T(j - 1, k - 1).sbar * inv_req; }
                if (j <= N) {
                    if (j == k) { v[6] = offd[k]; v[7] = 0.0; }
                    else { v[6] = T(j + 1, k).b; v[7] = T(j + 1, k).c; }
                }
                cr.insert(cr.end(), v, v + 8);
            };
            for (int k = 1; k <= ncols; k += 2) {   // walk order of grav_accel_cols: columns in pairs, rows interleaved
                const bool two = k + 1 <= ncols;
                push(k, k);
                for (int j = k + 1; j <= N + 1; ++j) { push(k, j); if (two) push(k + 1, j); }
            }
            cr.insert(cr.end(), 16, 0.0);           // two null records: targets of the last prefetches
            G.colrec = upload(e, cr.data(), cr.size());
            G.ncols = ncols;
            if (!G.colrec) { set_err("gravity table upload failed"); return false; }
        }
        if (!G.tab || !G.a_diag || !G.offdiag) { set_err("gravity table upload failed"); return false; }
    return true;
}

extern "C" nyxb_engine* nyxb_engine_create(const nyxb_dynamics* dyn, const nyxb_integ_opts* opts, int32_t mode,
                                            int32_t device) {
    if (!dyn || !opts) { set_err("null dynamics/options"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_err("no CUDA device available: nyxb has no CPU fallback");
        return nullptr;
    }
    if (device < 0 || device >= ndev) { set_err("bad device ordinal"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_err("cudaSetDevice failed"); return nullptr; }
    if (mode != NYXB_MODE_STRICT && mode != NYXB_MODE_FAST) { set_err("bad mode"); return nullptr; }
    if (dyn->n_bodies < 0 || dyn->n_bodies > NYXB_MAX_BODIES) { set_err("n_bodies out of range"); return nullptr; }

    nyxb_engine* e = new nyxb_engine();
    e->device = device;
    e->mode = mode;
    DevSetup& S = e->S;
    memset(&S, 0, sizeof(S));

    // ---- tableau (dense rows, c accumulated left to right as instance.rs:379-386 does)
    int order, stages;
    const NyxbCoef *a, *b;
    if (!tableau_for(opts->method, order, stages, a, b)) { set_err("unknown integration method"); delete e; return nullptr; }
    S.tb.stages = stages;
    S.tb.order = order;
    int idx = 0;
    for (int i = 0; i < stages - 1; ++i) {
        volatile double ci = 0.0;
        for (int j = 0; j <= i; ++j) {
            double aij = nyxb_tableau_value(a[idx++]);
            ci = ci + aij;
            S.tb.a[i * NYXB_MAX_STAGES + j] = aij;
        }
        S.tb.c[i] = ci;
    }
    for (int i = 0; i < stages; ++i) {
        S.tb.b[i] = nyxb_tableau_value(b[i]);
        volatile double d = S.tb.b[i] - nyxb_tableau_value(b[i + stages]);
        S.tb.e[i] = d;
    }
    S.error_ctrl = opts->error_ctrl;
    S.attempts = opts->attempts;
    S.fixed_step = opts->fixed_step;
    S.init_step_ns = opts->init_step_ns;
    S.min_step_ns = opts->min_step_ns;
    S.max_step_ns = opts->max_step_ns;
    S.tolerance = opts->tolerance;
    S.min_step_s = host_dur_to_seconds(opts->min_step_ns);
    S.max_step_s = host_dur_to_seconds(opts->max_step_ns);
    S.inv_order = 1.0 / (double)order;
    S.inv_order_m1 = 1.0 / (double)(order - 1);

    // ---- dynamics
    S.mu_central = dyn->mu_central_km3_s2;
    S.central_radius = dyn->central_radius_km;
    S.n_bodies = dyn->n_bodies;
    S.point_mass_mask = dyn->point_mass_mask;
    S.n_pm = 0;
    if (dyn->n_point_masses > 0) {   // `celestial_objects` order (orbital.rs:217)
        if (dyn->n_point_masses > dyn->n_bodies) { set_err("n_point_masses out of range"); delete e; return nullptr; }
        for (int q = 0; q < dyn->n_point_masses; ++q) {
            const int j = dyn->point_mass_order[q];
            if (j < 0 || j >= dyn->n_bodies) { set_err("point_mass_order: bad body index"); delete e; return nullptr; }
            S.pm_order[S.n_pm++] = (signed char)j;
        }
    } else {
        for (int j = 0; j < dyn->n_bodies; ++j)
            if ((dyn->point_mass_mask >> j) & 1u) S.pm_order[S.n_pm++] = (signed char)j;
    }
    S.state_center = -1;
    if (opts->state_center != 0) {
        if (opts->state_center < 0 || opts->state_center > dyn->n_bodies) { set_err("state_center: bad body index"); delete e; return nullptr; }
        S.state_center = opts->state_center - 1;
    }
    for (int j = 0; j < dyn->n_bodies; ++j) {
        const nyxb_body& hb = dyn->bodies[j];
        DevBody& db = S.bodies[j];
        db.mu = hb.mu_km3_s2; db.radius = hb.radius_km; db.t0_ns = hb.t0_ns; db.interval_ns = hb.interval_ns;
        db.n_intervals = hb.n_intervals; db.n_coeffs = hb.n_coeffs;
        db.inv_interval = 1.0 / (double)hb.interval_ns;
        db.coeffs = upload(e, hb.coeffs, (size_t)hb.n_intervals * 3 * hb.n_coeffs);
        if (!db.coeffs) { set_err("ephemeris upload failed"); delete e; return nullptr; }
    }
    const int n_grav = dyn->gravity ? std::max(1, dyn->n_gravity) : 0;
    if (n_grav > NYXB_MAX_FIELDS) { set_err("too many gravity fields"); delete e; return nullptr; }
    S.grav_body = NYXB_CENTRAL_BODY;
    for (int f = 0; f < n_grav; ++f) {
        const nyxb_gravity_field& g = dyn->gravity[f];
        if (g.body != NYXB_CENTRAL_BODY && (g.body < 0 || g.body >= dyn->n_bodies)) { set_err("gravity field: bad body index"); delete e; return nullptr; }
        if (f == 0) {
            if (!build_grav(e, g, S.grav, true)) { delete e; return nullptr; }
            S.has_grav = 1;
            S.grav_body = g.body;
        } else {
            if (!build_grav(e, g, S.xgrav[f - 1], false)) { delete e; return nullptr; }
            S.xgrav_body[f - 1] = g.body;
            S.n_xgrav = f;
        }
    }
    if (dyn->srp) {
        const nyxb_srp& s = *dyn->srp;
        if (s.sun_body < 0 || s.sun_body >= dyn->n_bodies || s.n_shadow < 0 || s.n_shadow > 4) {
            set_err("bad SRP descriptor"); delete e; return nullptr;
        }
        S.has_srp = 1;
        S.srp.phi = s.phi_w_m2; S.srp.sun_body = s.sun_body; S.srp.n_shadow = s.n_shadow; S.srp.estimate = s.estimate;
        for (int q = 0; q < 4; ++q) {
            S.srp.shadow_body[q] = s.shadow_body[q];
            if (q < s.n_shadow && s.shadow_body[q] != NYXB_CENTRAL_BODY && (s.shadow_body[q] < 0 || s.shadow_body[q] >= dyn->n_bodies)) {
                set_err("bad shadow body index"); delete e; return nullptr;
            }
        }
    }
    if (dyn->drag) {
        const nyxb_drag& d = *dyn->drag;
        S.has_drag = 1;
        S.drag.density = d.density; S.drag.rho0 = d.rho0; S.drag.r0 = d.r0; S.drag.ref_alt_m = d.ref_alt_m; S.drag.r_eq = d.r_eq_km;
        S.drag.rot = pack_rot(d.rot);
    }
    cudaEventCreate(&e->ev0);
    cudaEventCreate(&e->ev1);
    return e;
}

extern "C" void nyxb_engine_destroy(nyxb_engine* eng) {
    if (!eng) return;
    cudaSetDevice(eng->device);
    delete eng;
}

static bool coop_supported(const nyxb_engine* e, int lanes) {
    if (!e->S.has_grav) return false;
    return lanes == 8 || lanes == 16 || lanes == 32;
}

static int pick_lanes(const nyxb_engine* e, size_t n) {
    if (e->lanes > 0) return e->lanes;
    // auto: cooperative lanes only pay off when the harmonic sum dominates
    if (!e->S.has_grav || e->S.grav.N < 6) return 1;
    // FAST, moderate degree, very large ensembles: one thread per trajectory (column walk, grav_accel_cols) has enough warps to
    // hide its latencies and issues ~3x fewer instructions per trajectory than the cooperative kernel
    // (measured on B200, 21x21: 100 000 trajectories 1.13e8 vs 1.07e8 steps/s; 10 000 trajectories 5.0e7 vs 9.0e7)
    if (e->mode == NYXB_MODE_FAST && e->S.grav.N < 30 && n >= 65536) return 1;
    return (e->S.grav.N >= 48) ? 32 : ((e->S.grav.N >= 30) ? 16 : 8);
}

static const DevCoopStrict* get_scoop(nyxb_engine* e, int lanes) {
    auto it = e->scoop.find(lanes);
    if (it != e->scoop.end()) return &it->second;
    CoopStrictHost h;
    nyxb_coop_strict_build_host(e->S.grav.N, e->S.grav.M, lanes, h);
    DevCoopStrict d;
    d.G = h.G; d.kc = h.kc; d.kr = h.kr;
    d.cols = upload(e, h.cols.data(), h.cols.size());
    d.rows = upload(e, h.rows.data(), h.rows.size());
    if (!d.cols || !d.rows) return nullptr;
    return &(e->scoop[lanes] = d);
}

static const DevCoop* get_coop(nyxb_engine* e, int lanes) {
    auto it = e->coop.find(lanes);
    if (it != e->coop.end()) return &it->second;
    CoopHost h;
    nyxb_coop_build_host(e->S.grav.N, e->S.grav.M, e->h_cnm.data(), e->h_snm.data(), lanes, h);
    DevCoop d;
    d.G = h.G; d.L = h.L; d.kmax = h.kmax;
    d.recs = upload(e, h.recs.data(), h.recs.size());
    d.colseed = upload(e, h.colseed.data(), h.colseed.size());
    d.col_start = upload(e, h.col_start.data(), h.col_start.size());
    d.col_m = upload(e, h.col_m.data(), h.col_m.size());
    if (!d.recs || !d.col_start || !d.col_m || !d.colseed) return nullptr;
    return &(e->coop[lanes] = d);
}

static const DevTx* get_tx(nyxb_engine* e, int P) {
    auto it = e->tx.find(P);
    if (it != e->tx.end()) return &it->second;
    TxHost h;
    nyxb_tx_build_host(e->S.grav.N, e->S.grav.M, e->h_cnm.data(), e->h_snm.data(), P, h);
    std::vector<unsigned char> blob(nyxb_tx_pack_blob(&h, e->S.grav.N, nullptr));
    nyxb_tx_pack_blob(&h, e->S.grav.N, blob.data());
    DevTx d;
    d.P = h.P; d.n_rec = h.n_rec; d.kmax = h.kmax;
    d.recA = reinterpret_cast<const double*>(upload(e, blob.data(), blob.size()));
    if (!d.recA) return nullptr;
    return &(e->tx[P] = d);
}

// positions of the transposed kernel for this field: 8 warps per set up to degree 40, 16 beyond (shared-memory footprint of the table)
static int tx_positions(const nyxb_engine* e) {
    if (e->tx_positions) return e->tx_positions;
    return e->S.grav.N <= 40 ? 8 : 16;
}

static bool tx_supported(const nyxb_engine* e) {
    return e->mode == NYXB_MODE_FAST && e->S.has_grav && e->S.grav.N >= 8 && e->S.grav.N <= 70;
}

// Transposed kernel (nyxb_tx.cu): persistent CTAs, one set of 32 trajectories per CTA, (set, time-slice) tickets.
static int32_t launch_tx(nyxb_engine* e, size_t n, const double* state, const double* consts, const int64_t* epoch0, int64_t end_epoch,
                         int64_t* step_io, double* out_state, int64_t* out_epoch, nyxb_details* out_details, int32_t* out_status,
                         const DevSink& sink, cudaStream_t stream) {
    const DevTx* tx = get_tx(e, tx_positions(e));
    if (!tx) { set_err("transposed-kernel table upload failed"); return NYXB_RC_CUDA; }
    size_t smem = 0;
    // sets of 32 trajectories.  Sets of 64 (a walker lane carrying two trajectories through every record load: half the shared-
    // memory wavefronts per FP64 instruction, but only one set context fits a CTA, so nothing covers the serial stretch between two
    // attempts) were built and measured slower: 1.36e8 against 1.51e8 steps/s on C2 (profiles/r02s_tx_variants.md); not dispatched.
    const int set_len = 32;
    const int occ = nyxb_tx_occupancy(&e->S, tx, &smem);
    if (occ < 1) { set_err("transposed kernel: tables do not fit in shared memory"); return NYXB_RC_UNSUPPORTED; }
    if (!e->sms) CUDA_TRY(cudaDeviceGetAttribute(&e->sms, cudaDevAttrMultiProcessorCount, e->device));
    const size_t n_sets = (n + set_len - 1) / set_len;
    // one persistent CTA per SM, `occ` set contexts each; tx_max_ctas (tests) shrinks the grid to force time slicing
    size_t ctas = (size_t)e->sms;
    if (e->tx_max_ctas > 0 && (size_t)e->tx_max_ctas < ctas) ctas = (size_t)e->tx_max_ctas;
    // fewer sets than SMs x contexts: spread them over all SMs (one set per CTA runs its stages at the helpers' pace instead of
    // alternating with a second set, and twice as many SMs work: 5 000 trajectories 239 ms with 79 CTAs, profiles/r02y_shards.md)
    ctas = std::min(ctas, n_sets);
    const size_t slots = ctas * occ;
    const int grid = (int)ctas;
    // workspace: [ctl 4 x i32 | ring n_sets x i32 | ws_step n i64 | ws_f64 2n | ws_flags n | details n]
    const size_t ctl_bytes = (16 + 4 * n_sets + 15) & ~(size_t)15;
    const size_t need = ctl_bytes + n * (8 + 16 + 8) + n * sizeof(nyxb_details);
    if (need > e->txq_bytes) {
        cudaFree(e->d_txq); e->d_txq = nullptr; e->txq_bytes = 0;
        CUDA_TRY(cudaMalloc(&e->d_txq, need));
        e->txq_bytes = need;
    }
    CUDA_TRY(cudaMemsetAsync(e->d_txq, 0, ctl_bytes, stream));
    DevTxQueue q;
    q.ctl = reinterpret_cast<int*>(e->d_txq);
    q.ring = q.ctl + 4;
    unsigned char* ws = e->d_txq + ctl_bytes;
    q.ws_step = reinterpret_cast<long long*>(ws);
    q.ws_f64 = reinterpret_cast<double*>(ws + 8 * n);
    q.ws_flags = reinterpret_cast<int*>(ws + 24 * n);
    q.details = out_details ? out_details : reinterpret_cast<nyxb_details*>(ws + 32 * n);
    q.n_sets = (int)n_sets;
    q.slice = (n_sets > slots) ? e->tx_slice : 0;   // every set resident: no parking
    q.trace = nullptr;
#ifdef NYXB_TX_TRACE
    // diagnostic build: timeline of CTA 0, dumped to $NYXB_TX_TRACE_FILE after the launch (synchronises the stream)
    const size_t trace_bytes = (size_t)32 * NYXB_TX_TRACE_CAP * sizeof(unsigned long long);
    const char* trace_file = getenv("NYXB_TX_TRACE_FILE");
    if (trace_file) {
        CUDA_TRY(cudaMalloc(&q.trace, trace_bytes));
        CUDA_TRY(cudaMemsetAsync(q.trace, 0, trace_bytes, stream));
    }
#endif
    cudaError_t err = nyxb_launch_tx(&e->S, tx, &q, n, state, consts, (const long long*)epoch0, end_epoch, (long long*)step_io, out_state,
                                     (long long*)out_epoch, out_status, &sink, grid, stream);
    if (err != cudaSuccess) { set_err(std::string("kernel launch: ") + cudaGetErrorString(err)); return NYXB_RC_CUDA; }
#ifdef NYXB_TX_TRACE
    if (q.trace) {
        std::vector<unsigned long long> host((size_t)32 * NYXB_TX_TRACE_CAP);
        CUDA_TRY(cudaStreamSynchronize(stream));
        CUDA_TRY(cudaMemcpy(host.data(), q.trace, trace_bytes, cudaMemcpyDeviceToHost));
        cudaFree(q.trace);
        if (FILE* f = fopen(trace_file, "wb")) { fwrite(host.data(), 1, trace_bytes, f); fclose(f); }
    }
#endif
    e->launches += 1;
    e->last_kernel = NYXB_KERNEL_TRANSPOSED;
    return NYXB_RC_OK;
}

// kernel family for this call (nyxb_engine_set_kernel overrides the automatic choice)
static int pick_kernel(const nyxb_engine* e, size_t n) {
    if (e->kernel == NYXB_KERNEL_TRANSPOSED) return tx_supported(e) ? NYXB_KERNEL_TRANSPOSED : NYXB_KERNEL_COOP;
    if (e->kernel != NYXB_KERNEL_AUTO) return e->kernel;
    if (e->lanes > 0) return e->lanes == 1 ? NYXB_KERNEL_THREAD : NYXB_KERNEL_COOP;
    // auto: transposed kernel from 32 sets up (measured on B200, 21x21, 3 days: 1 250 trajectories 3.31e7 against 2.69e7 steps/s for the
    // lane-cooperative kernel at its best lane count, 2 500: 6.65e7 / 4.65e7, 5 000: 8.44e7 / 7.33e7, 10 000: 1.58e8 / 0.96e8;
    // profiles/r02y_shards.md)
    // every degree the transposed kernel serves (8..70; 16 walker positions and one set context per CTA above degree 40): GRAIL 70x70,
    // 10 000 low lunar orbits 2.23e7 against 7.6e6 steps/s for the lane-cooperative kernel at 32 lanes, 2 000: 9.4e6 / 6.7e6
    // (profiles/r02c4_kernels.md)
    if (tx_supported(e) && n >= (size_t)1024) return NYXB_KERNEL_TRANSPOSED;
    return NYXB_KERNEL_AUTO;
}

static int32_t launch_inner(nyxb_engine* e, size_t n, const double* state, const double* consts, const int64_t* epoch0,
                            int64_t end_epoch, int64_t* step_io, double* out_state, int64_t* out_epoch,
                            nyxb_details* out_details, int32_t* out_status, const DevSink& sink, cudaStream_t stream);

// propagation launch with the integration_frame translations around it (instance.rs:117-142, 167-176, 211-220)
static int32_t launch(nyxb_engine* e, size_t n, const double* state, const double* consts, const int64_t* epoch0,
                      int64_t end_epoch, int64_t* step_io, double* out_state, int64_t* out_epoch,
                      nyxb_details* out_details, int32_t* out_status, const DevSink& sink, cudaStream_t stream) {
    if (e->S.state_center < 0 || n == 0)
        return launch_inner(e, n, state, consts, epoch0, end_epoch, step_io, out_state, out_epoch, out_details, out_status, sink, stream);
    if (n > e->frame_n) {
        cudaFree(e->d_frame); e->d_frame = nullptr; e->frame_n = 0;
        CUDA_TRY(cudaMalloc(&e->d_frame, sizeof(double) * 9 * n));
        e->frame_n = n;
    }
    const DevBody* body = &e->S.bodies[e->S.state_center];
    CUDA_TRY(cudaMemcpyAsync(e->d_frame, state, sizeof(double) * 9 * n, cudaMemcpyDeviceToDevice, stream));
    CUDA_TRY(nyxb_launch_frame_shift(body, 1.0, n, e->d_frame, (const long long*)epoch0, nullptr, stream));
    const int32_t rc = launch_inner(e, n, e->d_frame, consts, epoch0, end_epoch, step_io, out_state, out_epoch, out_details, out_status, sink, stream);
    if (rc != NYXB_RC_OK) return rc;
    CUDA_TRY(nyxb_launch_frame_shift(body, -1.0, n, out_state, (const long long*)out_epoch, out_status, stream));
    e->launches += 2;
    return NYXB_RC_OK;
}

static int32_t launch_inner(nyxb_engine* e, size_t n, const double* state, const double* consts, const int64_t* epoch0,
                            int64_t end_epoch, int64_t* step_io, double* out_state, int64_t* out_epoch,
                            nyxb_details* out_details, int32_t* out_status, const DevSink& sink, cudaStream_t stream) {
    if (pick_kernel(e, n) == NYXB_KERNEL_TRANSPOSED)
        return launch_tx(e, n, state, consts, epoch0, end_epoch, step_io, out_state, out_epoch, out_details, out_status, sink, stream);
    int lanes = pick_lanes(e, n);
    if (e->kernel == NYXB_KERNEL_THREAD) lanes = 1;
    if (e->kernel == NYXB_KERNEL_COOP && lanes == 1 && e->S.has_grav) lanes = (e->S.grav.N >= 48) ? 32 : ((e->S.grav.N >= 30) ? 16 : 8);
    e->last_kernel = lanes > 1 ? NYXB_KERNEL_COOP : NYXB_KERNEL_THREAD;
    cudaError_t err;
    if (lanes > 1 && e->mode == NYXB_MODE_STRICT) {
        const DevCoopStrict* cs = get_scoop(e, lanes);
        if (!cs) { set_err("cooperative schedule upload failed"); return NYXB_RC_CUDA; }
        err = nyxb_launch_coop_strict(&e->S, cs, n, state, consts, (const long long*)epoch0, end_epoch, (long long*)step_io,
                                      out_state, (long long*)out_epoch, out_details, out_status, &sink, stream);
    } else if (lanes > 1) {
        const DevCoop* cp = get_coop(e, lanes);
        if (!cp) { set_err("cooperative table upload failed"); return NYXB_RC_CUDA; }
        // one trajectory per lane group: register blocking over two trajectories (T = 2) was measured slower at every ensemble
        // size in two rounds (profiles/r02a_k2_variants.md) and is no longer dispatched
        err = nyxb_launch_coop(&e->S, cp, 1, n, state, consts, (const long long*)epoch0, end_epoch, (long long*)step_io,
                               out_state, (long long*)out_epoch, out_details, out_status, &sink, stream);
    } else if (e->mode == NYXB_MODE_STRICT) {
        err = nyxb_launch_thread_strict(&e->S, n, state, consts, (const long long*)epoch0, end_epoch, (long long*)step_io,
                                        out_state, (long long*)out_epoch, out_details, out_status, 64, &sink, stream);
    } else {
        const int blk = 64;   // 32 / 128 threads per CTA measured no better (profiles/README.md, r01n)
        err = nyxb_launch_thread_fast(&e->S, n, state, consts, (const long long*)epoch0, end_epoch, (long long*)step_io,
                                      out_state, (long long*)out_epoch, out_details, out_status, blk, &sink, stream);
    }
    if (err != cudaSuccess) { set_err(std::string("kernel launch: ") + cudaGetErrorString(err)); return NYXB_RC_CUDA; }
    e->launches += 1;
    return NYXB_RC_OK;
}

static DevSink make_sink(const nyxb_traj_sink* sink) {
    DevSink d{};
    if (sink && sink->capacity > 0 && sink->epoch_ns && sink->state && sink->count) {
        d.cap = sink->capacity; d.epoch = (long long*)sink->epoch_ns; d.state = sink->state; d.count = (long long*)sink->count;
    }
    return d;
}

extern "C" int32_t nyxb_propagate_batch_traj_dev(nyxb_engine* eng, size_t n, const double* state_soa, const double* consts_soa,
                                                 const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                                                 double* out_state_soa, int64_t* out_epoch_ns, nyxb_details* out_details,
                                                 int32_t* out_status, const nyxb_traj_sink* sink, void* cuda_stream) {
    if (!eng || !state_soa || !consts_soa || !epoch0_ns || !out_state_soa || !out_epoch_ns || !out_status) {
        set_err("null argument");
        return NYXB_RC_BAD_ARG;
    }
    CUDA_TRY(cudaSetDevice(eng->device));
    return launch(eng, n, state_soa, consts_soa, epoch0_ns, end_epoch_ns, step_ns, out_state_soa, out_epoch_ns, out_details,
                  out_status, make_sink(sink), (cudaStream_t)cuda_stream);
}

extern "C" int32_t nyxb_propagate_batch_dev(nyxb_engine* eng, size_t n, const double* state_soa, const double* consts_soa,
                                            const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                                            double* out_state_soa, int64_t* out_epoch_ns, nyxb_details* out_details,
                                            int32_t* out_status, void* cuda_stream) {
    return nyxb_propagate_batch_traj_dev(eng, n, state_soa, consts_soa, epoch0_ns, end_epoch_ns, step_ns, out_state_soa,
                                         out_epoch_ns, out_details, out_status, nullptr, cuda_stream);
}

extern "C" int32_t nyxb_propagate_batch_event(nyxb_engine* eng, size_t n, const double* state_soa, const double* consts_soa,
                                              const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                                              double* out_state_soa, int64_t* out_epoch_ns, nyxb_details* out_details,
                                              int32_t* out_status, const nyxb_traj_sink* sink, const nyxb_event* event) {
    if (event && event->kind != NYXB_EVENT_NONE &&
        (event->kind < NYXB_EVENT_RMAG || event->kind > NYXB_EVENT_VMAG || event->trigger < 1 || !event->crossings)) {
        set_err("bad event descriptor");
        return NYXB_RC_BAD_ARG;
    }
    const bool has_ev = event && event->kind != NYXB_EVENT_NONE;
    if (!eng || !state_soa || !consts_soa || !epoch0_ns || !out_state_soa || !out_epoch_ns || !out_status) {
        set_err("null argument");
        return NYXB_RC_BAD_ARG;
    }
    if (n == 0) return NYXB_RC_OK;
    CUDA_TRY(cudaSetDevice(eng->device));
    // device slab: [state 9n | consts 4n | out_state 9n] doubles, [epoch0 n | out_epoch n | step n] i64, details, status
    int32_t rc = NYXB_RC_OK;
    cudaError_t ce;
#define TRY2(x) do { ce = (x); if (ce != cudaSuccess) { set_err(std::string(#x) + ": " + cudaGetErrorString(ce)); return NYXB_RC_CUDA; } } while (0)
    if (!eng->stream) TRY2(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
    if (n > eng->cap) {
        cudaFree(eng->d_f64); cudaFree(eng->d_i64); cudaFree(eng->d_det); cudaFree(eng->d_status);
        eng->d_f64 = nullptr; eng->d_i64 = nullptr; eng->d_det = nullptr; eng->d_status = nullptr; eng->cap = 0;
        TRY2(cudaMalloc(&eng->d_f64, sizeof(double) * 22 * n));
        TRY2(cudaMalloc(&eng->d_i64, sizeof(long long) * 3 * n));
        TRY2(cudaMalloc(&eng->d_det, sizeof(nyxb_details) * n));
        TRY2(cudaMalloc(&eng->d_status, sizeof(int) * 2 * n));  // status | event crossings
        eng->cap = n;
    }
    double* d_f64 = eng->d_f64;
    long long* d_i64 = eng->d_i64;
    nyxb_details* d_det = eng->d_det;
    int* d_status = eng->d_status;
    cudaStream_t st = eng->stream;
    TRY2(cudaMemcpyAsync(d_f64, state_soa, sizeof(double) * 9 * n, cudaMemcpyHostToDevice, st));
    TRY2(cudaMemcpyAsync(d_f64 + 9 * n, consts_soa, sizeof(double) * 4 * n, cudaMemcpyHostToDevice, st));
    TRY2(cudaMemcpyAsync(d_i64, epoch0_ns, sizeof(long long) * n, cudaMemcpyHostToDevice, st));
    if (step_ns) TRY2(cudaMemcpyAsync(d_i64 + 2 * n, step_ns, sizeof(long long) * n, cudaMemcpyHostToDevice, st));
    // trajectory sink on the device: [epoch cap*n i64 | state 6*cap*n f64 | count n i64]
    DevSink dsink{};
    if (has_ev) { dsink.ev_kind = event->kind; dsink.ev_trigger = event->trigger; dsink.ev_value = event->value; dsink.ev_crossings = d_status + n; }
    const bool rec = sink && sink->capacity > 0 && sink->epoch_ns && sink->state && sink->count;
    eng->rec_n = 0; eng->rec_cap = 0;   // a propagation that does not record invalidates the resident recording (sink == NULL calls)
    if (rec) {
        const size_t cap = (size_t)sink->capacity;
        const size_t need = (cap * n * 7 + n) * 8;
        if (need > eng->sink_bytes) {
            cudaFree(eng->d_sink); eng->d_sink = nullptr; eng->sink_bytes = 0;
            TRY2(cudaMalloc(&eng->d_sink, need));
            eng->sink_bytes = need;
        }
        dsink.cap = sink->capacity;
        dsink.epoch = (long long*)eng->d_sink;
        dsink.state = (double*)(eng->d_sink + cap * n * 8);
        dsink.count = (long long*)(eng->d_sink + cap * n * 56);
        eng->rec_n = n; eng->rec_cap = sink->capacity;
        // slots past count[i] come back as zeros, not as whatever the allocation held (the whole sink is copied to the caller)
        TRY2(cudaMemsetAsync(eng->d_sink, 0, cap * n * 56, st));
    }
    TRY2(cudaEventRecord(eng->ev0, st));
    rc = launch(eng, n, d_f64, d_f64 + 9 * n, (const int64_t*)d_i64, end_epoch_ns, step_ns ? (int64_t*)(d_i64 + 2 * n) : nullptr,
                d_f64 + 13 * n, (int64_t*)(d_i64 + n), d_det, d_status, dsink, st);
    if (rc != NYXB_RC_OK) return rc;
    TRY2(cudaEventRecord(eng->ev1, st));
    TRY2(cudaMemcpyAsync(out_state_soa, d_f64 + 13 * n, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost, st));
    TRY2(cudaMemcpyAsync(out_epoch_ns, d_i64 + n, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
    if (step_ns) TRY2(cudaMemcpyAsync(step_ns, d_i64 + 2 * n, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
    if (out_details) TRY2(cudaMemcpyAsync(out_details, d_det, sizeof(nyxb_details) * n, cudaMemcpyDeviceToHost, st));
    TRY2(cudaMemcpyAsync(out_status, d_status, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    if (has_ev) TRY2(cudaMemcpyAsync(event->crossings, d_status + n, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    if (rec) {
        const size_t cap = (size_t)sink->capacity;
        TRY2(cudaMemcpyAsync(sink->epoch_ns, dsink.epoch, cap * n * 8, cudaMemcpyDeviceToHost, st));
        TRY2(cudaMemcpyAsync(sink->state, dsink.state, cap * n * 48, cudaMemcpyDeviceToHost, st));
        TRY2(cudaMemcpyAsync(sink->count, dsink.count, n * 8, cudaMemcpyDeviceToHost, st));
    }
    TRY2(cudaStreamSynchronize(st));
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, eng->ev0, eng->ev1) == cudaSuccess) eng->last_ms = ms;
    }
    return rc;
#undef TRY2
}

extern "C" int32_t nyxb_propagate_batch_traj(nyxb_engine* eng, size_t n, const double* state_soa, const double* consts_soa,
                                             const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                                             double* out_state_soa, int64_t* out_epoch_ns, nyxb_details* out_details,
                                             int32_t* out_status, const nyxb_traj_sink* sink) {
    return nyxb_propagate_batch_event(eng, n, state_soa, consts_soa, epoch0_ns, end_epoch_ns, step_ns, out_state_soa, out_epoch_ns,
                                      out_details, out_status, sink, nullptr);
}

extern "C" int32_t nyxb_propagate_batch(nyxb_engine* eng, size_t n, const double* state_soa, const double* consts_soa,
                                        const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                                        double* out_state_soa, int64_t* out_epoch_ns, nyxb_details* out_details,
                                        int32_t* out_status) {
    return nyxb_propagate_batch_traj(eng, n, state_soa, consts_soa, epoch0_ns, end_epoch_ns, step_ns, out_state_soa, out_epoch_ns,
                                     out_details, out_status, nullptr);
}

// ---------------------------------------------------------------------------------------------------------------------
// Multi-GPU fan-out behind the boundary (mc/montecarlo.rs:233-253: contiguous run-index ranges, no exchange while
// integrating).  engines[g] must have been created from the same (dynamics, options) on DIFFERENT devices (or the same device, for
// tests); shard g = runs [g n / G, (g+1) n / G).  All uploads and launches are enqueued first, then the results of every shard are
// copied straight into the caller's [9][n] arrays — for a host caller that IS the gather of final states.
// ---------------------------------------------------------------------------------------------------------------------
static int32_t ensure_slab(nyxb_engine* eng, size_t n) {
    cudaError_t ce;
#define TRY3(x) do { ce = (x); if (ce != cudaSuccess) { set_err(std::string(#x) + ": " + cudaGetErrorString(ce)); return NYXB_RC_CUDA; } } while (0)
    if (!eng->stream) TRY3(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
    if (n > eng->cap) {
        cudaFree(eng->d_f64); cudaFree(eng->d_i64); cudaFree(eng->d_det); cudaFree(eng->d_status);
        eng->d_f64 = nullptr; eng->d_i64 = nullptr; eng->d_det = nullptr; eng->d_status = nullptr; eng->cap = 0;
        TRY3(cudaMalloc(&eng->d_f64, sizeof(double) * 22 * n));
        TRY3(cudaMalloc(&eng->d_i64, sizeof(long long) * 3 * n));
        TRY3(cudaMalloc(&eng->d_det, sizeof(nyxb_details) * n));
        TRY3(cudaMalloc(&eng->d_status, sizeof(int) * 2 * n));
        eng->cap = n;
    }
    return NYXB_RC_OK;
#undef TRY3
}

extern "C" int32_t nyxb_propagate_batch_multi(nyxb_engine* const* engines, int32_t n_engines, size_t n, const double* state_soa,
                                              const double* consts_soa, const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                                              double* out_state_soa, int64_t* out_epoch_ns, nyxb_details* out_details, int32_t* out_status) {
    if (!engines || n_engines < 1 || !state_soa || !consts_soa || !epoch0_ns || !out_state_soa || !out_epoch_ns || !out_status) {
        set_err("null argument");
        return NYXB_RC_BAD_ARG;
    }
    for (int32_t g = 0; g < n_engines; ++g)
        if (!engines[g]) { set_err("null engine"); return NYXB_RC_BAD_ARG; }
    if (n == 0) return NYXB_RC_OK;
    const size_t G = (size_t)n_engines;
    auto lo_of = [&](size_t g) { return g * n / G; };
    cudaError_t ce;
#define TRY4(x) do { ce = (x); if (ce != cudaSuccess) { set_err(std::string(#x) + ": " + cudaGetErrorString(ce)); return NYXB_RC_CUDA; } } while (0)
    // ---- phase 1: uploads + launches on every device (asynchronous: the devices integrate concurrently)
    for (size_t g = 0; g < G; ++g) {
        nyxb_engine* eng = engines[g];
        const size_t lo = lo_of(g), m = lo_of(g + 1) - lo;
        if (m == 0) continue;
        TRY4(cudaSetDevice(eng->device));
        int32_t rc = ensure_slab(eng, m);
        if (rc != NYXB_RC_OK) return rc;
        cudaStream_t st = eng->stream;
        // strided rows of the caller's [rows][n] arrays -> dense [rows][m] shards
        TRY4(cudaMemcpy2DAsync(eng->d_f64, m * 8, state_soa + lo, n * 8, m * 8, 9, cudaMemcpyHostToDevice, st));
        TRY4(cudaMemcpy2DAsync(eng->d_f64 + 9 * m, m * 8, consts_soa + lo, n * 8, m * 8, 4, cudaMemcpyHostToDevice, st));
        TRY4(cudaMemcpyAsync(eng->d_i64, epoch0_ns + lo, m * 8, cudaMemcpyHostToDevice, st));
        if (step_ns) TRY4(cudaMemcpyAsync(eng->d_i64 + 2 * m, step_ns + lo, m * 8, cudaMemcpyHostToDevice, st));
        eng->rec_n = 0; eng->rec_cap = 0;
        TRY4(cudaEventRecord(eng->ev0, st));
        rc = launch(eng, m, eng->d_f64, eng->d_f64 + 9 * m, (const int64_t*)eng->d_i64, end_epoch_ns,
                    step_ns ? (int64_t*)(eng->d_i64 + 2 * m) : nullptr, eng->d_f64 + 13 * m, (int64_t*)(eng->d_i64 + m), eng->d_det,
                    eng->d_status, DevSink{}, st);
        if (rc != NYXB_RC_OK) return rc;
        TRY4(cudaEventRecord(eng->ev1, st));
    }
    // ---- phase 2: every shard's results straight into the caller's arrays (the gather), then one synchronisation per device
    for (size_t g = 0; g < G; ++g) {
        nyxb_engine* eng = engines[g];
        const size_t lo = lo_of(g), m = lo_of(g + 1) - lo;
        if (m == 0) continue;
        TRY4(cudaSetDevice(eng->device));
        cudaStream_t st = eng->stream;
        TRY4(cudaMemcpy2DAsync(out_state_soa + lo, n * 8, eng->d_f64 + 13 * m, m * 8, m * 8, 9, cudaMemcpyDeviceToHost, st));
        TRY4(cudaMemcpyAsync(out_epoch_ns + lo, eng->d_i64 + m, m * 8, cudaMemcpyDeviceToHost, st));
        if (step_ns) TRY4(cudaMemcpyAsync(step_ns + lo, eng->d_i64 + 2 * m, m * 8, cudaMemcpyDeviceToHost, st));
        if (out_details) TRY4(cudaMemcpyAsync(out_details + lo, eng->d_det, sizeof(nyxb_details) * m, cudaMemcpyDeviceToHost, st));
        TRY4(cudaMemcpyAsync(out_status + lo, eng->d_status, sizeof(int) * m, cudaMemcpyDeviceToHost, st));
    }
    for (size_t g = 0; g < G; ++g) {
        nyxb_engine* eng = engines[g];
        if (lo_of(g + 1) == lo_of(g)) continue;
        TRY4(cudaSetDevice(eng->device));
        TRY4(cudaStreamSynchronize(eng->stream));
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, eng->ev0, eng->ev1) == cudaSuccess) eng->last_ms = ms;
    }
    return NYXB_RC_OK;
#undef TRY4
}

// ---------------------------------------------------------------------------------------------------------------------
// (f)-2: STM propagation and the batched sequential filter (kernels in nyxb_od.cu).  Host-pointer entry points with
// per-call device buffers (these calls run for seconds; allocation cost is irrelevant).
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct DevBufs {
    std::vector<void*> p;
    ~DevBufs() { for (void* q : p) cudaFree(q); }
    template <typename T> T* alloc(size_t count) {
        T* d = nullptr;
        if (cudaMalloc(&d, sizeof(T) * (count ? count : 1)) != cudaSuccess) return nullptr;
        p.push_back(d);
        return d;
    }
    template <typename T> T* put(const T* host, size_t count, cudaStream_t st) {
        T* d = alloc<T>(count);
        if (d && count && cudaMemcpyAsync(d, host, sizeof(T) * count, cudaMemcpyHostToDevice, st) != cudaSuccess) return nullptr;
        return d;
    }
};
bool stm_supported(const nyxb_engine* e) {
    if (e->S.grav_body >= 0 || e->S.n_xgrav > 0 || e->S.state_center >= 0) {
        set_err("the STM / filter kernels take one harmonic field, of the integration centre, and states in the integration frame");
        return false;
    }
    if (e->S.has_drag) { set_err("PartialsUndefined: the drag model has no partials (drag.rs:109-118, 286-295)"); return false; }
    if (!e->S.fixed_step && e->S.error_ctrl != NYXB_RSS_CARTESIAN_STATE && e->S.error_ctrl != NYXB_RSS_CARTESIAN_STEP) {
        set_err("STM propagation accepts the Cartesian error controls or a fixed step");
        return false;
    }
    return true;
}
}  // namespace

extern "C" int32_t nyxb_propagate_batch_stm(nyxb_engine* eng, size_t n, const double* state_soa, const double* consts_soa,
                                            const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                                            const double* stm_in_soa, double* out_state_soa, int64_t* out_epoch_ns,
                                            double* out_stm_soa, nyxb_details* out_details, int32_t* out_status) {
    if (!eng || !state_soa || !consts_soa || !epoch0_ns || !out_state_soa || !out_epoch_ns || !out_stm_soa || !out_status) {
        set_err("null argument");
        return NYXB_RC_BAD_ARG;
    }
    if (!stm_supported(eng)) return NYXB_RC_UNSUPPORTED;
    if (n == 0) return NYXB_RC_OK;
    CUDA_TRY(cudaSetDevice(eng->device));
    if (!eng->stream) CUDA_TRY(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
    cudaStream_t st = eng->stream;
    DevBufs B;
    double* d_state = B.put(state_soa, 9 * n, st);
    double* d_consts = B.put(consts_soa, 4 * n, st);
    long long* d_ep = B.put((const long long*)epoch0_ns, n, st);
    long long* d_step = step_ns ? B.put((const long long*)step_ns, n, st) : nullptr;
    double* d_stm_in = stm_in_soa ? B.put(stm_in_soa, 81 * n, st) : nullptr;
    double* d_out = B.alloc<double>(9 * n);
    double* d_stm = B.alloc<double>(81 * n);
    long long* d_oep = B.alloc<long long>(n);
    nyxb_details* d_det = B.alloc<nyxb_details>(n);
    int* d_status = B.alloc<int>(n);
    if (!d_state || !d_consts || !d_ep || (step_ns && !d_step) || (stm_in_soa && !d_stm_in) || !d_out || !d_stm || !d_oep || !d_det || !d_status) {
        set_err("device allocation / upload failed");
        return NYXB_RC_CUDA;
    }
    CUDA_TRY(cudaEventRecord(eng->ev0, st));
    cudaError_t err = (eng->mode == NYXB_MODE_STRICT)
        ? nyxb_launch_stm_strict(&eng->S, n, d_state, d_consts, d_ep, end_epoch_ns, d_step, d_stm_in, d_out, d_oep, d_stm, d_det, d_status, st)
        : nyxb_launch_stm_fast(&eng->S, n, d_state, d_consts, d_ep, end_epoch_ns, d_step, d_stm_in, d_out, d_oep, d_stm, d_det, d_status, st);
    if (err != cudaSuccess) { set_err(std::string("kernel launch: ") + cudaGetErrorString(err)); return NYXB_RC_CUDA; }
    eng->launches += 1;
    CUDA_TRY(cudaEventRecord(eng->ev1, st));
    CUDA_TRY(cudaMemcpyAsync(out_state_soa, d_out, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_stm_soa, d_stm, sizeof(double) * 81 * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_epoch_ns, d_oep, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
    if (step_ns) CUDA_TRY(cudaMemcpyAsync(step_ns, d_step, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
    if (out_details) CUDA_TRY(cudaMemcpyAsync(out_details, d_det, sizeof(nyxb_details) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_status, d_status, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, eng->ev0, eng->ev1) == cudaSuccess) eng->last_ms = ms;
    return NYXB_RC_OK;
}

extern "C" int32_t nyxb_od_ekf_batch(nyxb_engine* eng, const nyxb_od_config* cfg, int32_t n_stations,
                                     const nyxb_ground_station* stations, const nyxb_tracking_arc* arc, size_t n,
                                     const double* state_soa, const double* consts_soa, const int64_t* epoch0_ns,
                                     const double* covar0_soa, const nyxb_od_outputs* out) {
    if (!eng || !cfg || !arc || !state_soa || !consts_soa || !epoch0_ns || !covar0_soa || !out || !out->state_soa || !out->epoch_ns ||
        !out->covar_soa || !out->status || (n_stations > 0 && !stations) || n_stations < 0) {
        set_err("null argument");
        return NYXB_RC_BAD_ARG;
    }
    if (!stm_supported(eng)) return NYXB_RC_UNSUPPORTED;
    if (cfg->msr_size != 1 && cfg->msr_size != 2) { set_err("msr_size must be 1 or 2"); return NYXB_RC_BAD_ARG; }
    if (cfg->variant != NYXB_KF_REFERENCE_UPDATE && cfg->variant != NYXB_KF_DEVIATION_TRACKING) { set_err("bad filter variant"); return NYXB_RC_BAD_ARG; }
    if (cfg->max_step_ns <= 0) { set_err("StepSize: max_step must be positive (process/mod.rs:147-150)"); return NYXB_RC_BAD_ARG; }
    if (arc->n_msr < 2) { set_err("TooFewMeasurements: need 2 (process/mod.rs:139-145)"); return NYXB_RC_BAD_ARG; }
    if (!arc->epoch_ns || !arc->tracker || !arc->obs) { set_err("null tracking arc arrays"); return NYXB_RC_BAD_ARG; }
    for (int32_t s = 0; s < n_stations; ++s) {
        const nyxb_ground_station& g = stations[s];
        if (g.n_types < 1 || g.n_types > 2 || (g.body != NYXB_CENTRAL_BODY && (g.body < 0 || g.body >= eng->S.n_bodies))) {
            set_err("bad ground station descriptor");
            return NYXB_RC_BAD_ARG;
        }
        for (int q = 0; q < g.n_types; ++q)
            if (g.types[q] != NYXB_MSR_RANGE && g.types[q] != NYXB_MSR_DOPPLER) { set_err("unsupported measurement type"); return NYXB_RC_UNSUPPORTED; }
        if (g.n_types % cfg->msr_size != 0) { set_err("filter misconfigured: measurement types per device must be a multiple of msr_size"); return NYXB_RC_UNSUPPORTED; }
    }
    if (n == 0) return NYXB_RC_OK;
    CUDA_TRY(cudaSetDevice(eng->device));
    if (!eng->stream) CUDA_TRY(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
    cudaStream_t st = eng->stream;
    const size_t m = (size_t)arc->n_msr;
    DevBufs B;
    std::vector<DevStation> hs((size_t)n_stations);
    for (int32_t s = 0; s < n_stations; ++s) {
        const nyxb_ground_station& g = stations[s];
        DevStation& d = hs[s];
        for (int q = 0; q < 3; ++q) { d.pos[q] = g.pos_fixed_km[q]; d.up[q] = g.up_fixed[q]; }
        d.mask_deg = g.elevation_mask_deg; d.rot = pack_rot(g.rot); d.body = g.body; d.n_types = g.n_types;
        for (int q = 0; q < 2; ++q) { d.types[q] = g.types[q]; d.noise_var[q] = g.noise_var[q]; d.bias[q] = g.bias[q]; }
        d.body_radius = g.body_radius_km;
    }
    DevOd od{};
    od.variant = cfg->variant; od.msr_size = cfg->msr_size; od.reject = cfg->reject_num_sigmas;
    od.max_step_ns = cfg->max_step_ns; od.eps_ns = cfg->epoch_precision_ns;
    od.snc_enabled = cfg->snc_enabled; od.snc_frame = cfg->snc_frame;
    for (int q = 0; q < 3; ++q) od.snc_diag[q] = cfg->snc_diag[q];
    od.snc_disable_ns = cfg->snc_disable_time_ns;
    od.n_stations = n_stations;
    od.stations = B.put(hs.data(), hs.size(), st);
    od.n_msr = arc->n_msr;
    od.msr_epoch = B.put((const long long*)arc->epoch_ns, m, st);
    od.msr_tracker = B.put((const int*)arc->tracker, m, st);
    od.obs = B.put(arc->obs, m * 2 * n, st);
    od.covar0 = B.put(covar0_soa, 81 * n, st);
    double* d_state = B.put(state_soa, 9 * n, st);
    double* d_consts = B.put(consts_soa, 4 * n, st);
    long long* d_ep = B.put((const long long*)epoch0_ns, n, st);
    double* d_out = B.alloc<double>(9 * n);
    long long* d_oep = B.alloc<long long>(n);
    nyxb_details* d_det = B.alloc<nyxb_details>(n);
    int* d_status = B.alloc<int>(n);
    od.covar = B.alloc<double>(81 * n);
    od.state_dev = out->state_dev_soa ? B.alloc<double>(9 * n) : nullptr;
    od.ratio = out->resid_ratio ? B.alloc<double>(m * 2 * n) : nullptr;
    od.prefit = out->prefit ? B.alloc<double>(m * 2 * n) : nullptr;
    od.postfit = out->postfit ? B.alloc<double>(m * 2 * n) : nullptr;
    od.flags = out->msr_flags ? B.alloc<int>(m * n) : nullptr;
    od.est_state = out->est_state ? B.alloc<double>(m * 9 * n) : nullptr;
    od.est_cov = out->est_covar_diag ? B.alloc<double>(m * 9 * n) : nullptr;
    if (!od.stations || !od.msr_epoch || !od.msr_tracker || !od.obs || !od.covar0 || !d_state || !d_consts || !d_ep || !d_out || !d_oep ||
        !d_det || !d_status || !od.covar || (out->state_dev_soa && !od.state_dev) || (out->resid_ratio && !od.ratio) ||
        (out->prefit && !od.prefit) || (out->postfit && !od.postfit) || (out->msr_flags && !od.flags) ||
        (out->est_state && !od.est_state) || (out->est_covar_diag && !od.est_cov)) {
        set_err("device allocation / upload failed");
        return NYXB_RC_CUDA;
    }
    // per-measurement records default to NaN (0xFF bytes) / 0 flags where nothing is written
    if (od.ratio) CUDA_TRY(cudaMemsetAsync(od.ratio, 0xFF, sizeof(double) * m * 2 * n, st));
    if (od.prefit) CUDA_TRY(cudaMemsetAsync(od.prefit, 0xFF, sizeof(double) * m * 2 * n, st));
    if (od.postfit) CUDA_TRY(cudaMemsetAsync(od.postfit, 0xFF, sizeof(double) * m * 2 * n, st));
    if (od.flags) CUDA_TRY(cudaMemsetAsync(od.flags, 0, sizeof(int) * m * n, st));
    if (od.est_state) CUDA_TRY(cudaMemsetAsync(od.est_state, 0xFF, sizeof(double) * m * 9 * n, st));
    if (od.est_cov) CUDA_TRY(cudaMemsetAsync(od.est_cov, 0xFF, sizeof(double) * m * 9 * n, st));
    // FAST mode with a gravity field: one WARP per filter, the harmonic gradient split by columns over the lanes
    // (nyxb_od_coop.cu); columns -> lanes by longest-processing-time.  nyxb_engine_set_kernel(NYXB_KERNEL_THREAD) forces the
    // per-thread kernel.
    const int* d_cols = nullptr;
    bool coop = eng->mode == NYXB_MODE_FAST && eng->S.has_grav && eng->S.grav.N >= 8 && eng->kernel != NYXB_KERNEL_THREAD;
    if (coop) {
        const int N = eng->S.grav.N, mtop = eng->S.grav.M < N ? eng->S.grav.M : N, kmax = nyxb_od_coop_kmax();
        std::vector<int> cols(32 * (size_t)kmax, -1), cnt(32, 0);
        std::vector<long long> load(32, 0);
        for (int m = 0; m <= mtop && coop; ++m) {   // columns in decreasing length order: m = 0, 1 (same length), 2, ...
            int best = 0;
            for (int l = 1; l < 32; ++l)
                if (load[l] < load[best] || (load[l] == load[best] && cnt[l] < cnt[best])) best = l;
            if (cnt[best] >= kmax) { coop = false; break; }
            cols[(size_t)best * kmax + cnt[best]++] = m;
            load[best] += N - (m > 0 ? m : 1) + 1 + 6;   // entries + per-column overhead
        }
        if (coop) {
            d_cols = B.put(cols.data(), cols.size(), st);
            if (!d_cols) { set_err("device allocation / upload failed"); return NYXB_RC_CUDA; }
        }
    }
    CUDA_TRY(cudaEventRecord(eng->ev0, st));
    cudaError_t err = coop
        ? nyxb_launch_od_coop(&eng->S, &od, d_cols, n, d_state, d_consts, d_ep, d_out, d_oep, d_det, d_status, st)
        : (eng->mode == NYXB_MODE_STRICT)
            ? nyxb_launch_od_strict(&eng->S, &od, n, d_state, d_consts, d_ep, d_out, d_oep, d_det, d_status, st)
            : nyxb_launch_od_fast(&eng->S, &od, n, d_state, d_consts, d_ep, d_out, d_oep, d_det, d_status, st);
    if (err != cudaSuccess) { set_err(std::string("kernel launch: ") + cudaGetErrorString(err)); return NYXB_RC_CUDA; }
    eng->launches += 1;
    CUDA_TRY(cudaEventRecord(eng->ev1, st));
    CUDA_TRY(cudaMemcpyAsync(out->state_soa, d_out, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out->epoch_ns, d_oep, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out->covar_soa, od.covar, sizeof(double) * 81 * n, cudaMemcpyDeviceToHost, st));
    if (od.state_dev) CUDA_TRY(cudaMemcpyAsync(out->state_dev_soa, od.state_dev, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost, st));
    if (od.ratio) CUDA_TRY(cudaMemcpyAsync(out->resid_ratio, od.ratio, sizeof(double) * m * 2 * n, cudaMemcpyDeviceToHost, st));
    if (od.prefit) CUDA_TRY(cudaMemcpyAsync(out->prefit, od.prefit, sizeof(double) * m * 2 * n, cudaMemcpyDeviceToHost, st));
    if (od.postfit) CUDA_TRY(cudaMemcpyAsync(out->postfit, od.postfit, sizeof(double) * m * 2 * n, cudaMemcpyDeviceToHost, st));
    if (od.flags) CUDA_TRY(cudaMemcpyAsync(out->msr_flags, od.flags, sizeof(int) * m * n, cudaMemcpyDeviceToHost, st));
    if (od.est_state) CUDA_TRY(cudaMemcpyAsync(out->est_state, od.est_state, sizeof(double) * m * 9 * n, cudaMemcpyDeviceToHost, st));
    if (od.est_cov) CUDA_TRY(cudaMemcpyAsync(out->est_covar_diag, od.est_cov, sizeof(double) * m * 9 * n, cudaMemcpyDeviceToHost, st));
    if (out->details) CUDA_TRY(cudaMemcpyAsync(out->details, d_det, sizeof(nyxb_details) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out->status, d_status, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, eng->ev0, eng->ev1) == cudaSuccess) eng->last_ms = ms;
    return NYXB_RC_OK;
}

extern "C" int32_t nyxb_mvn_sample_dev(int32_t device, uint64_t seed, uint64_t first_index, size_t n, const double* template_state,
                                       const double* mean, const double* sqrt_s_v, double* out_state_soa, double* out_dispersion_soa,
                                       void* cuda_stream) {
    if (!template_state || !sqrt_s_v || !out_state_soa) { set_err("null argument"); return NYXB_RC_BAD_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_err("no CUDA device available: nyxb has no CPU fallback"); return NYXB_RC_NO_DEVICE; }
    if (device < 0 || device >= ndev) { set_err("bad device ordinal"); return NYXB_RC_BAD_ARG; }
    CUDA_TRY(cudaSetDevice(device));
    cudaError_t err = nyxb_launch_mvn(seed, first_index, n, template_state, mean, sqrt_s_v, out_state_soa, out_dispersion_soa,
                                      (cudaStream_t)cuda_stream);
    if (err != cudaSuccess) { set_err(std::string("kernel launch: ") + cudaGetErrorString(err)); return NYXB_RC_CUDA; }
    return NYXB_RC_OK;
}

extern "C" int32_t nyxb_mvn_sample(int32_t device, uint64_t seed, uint64_t first_index, size_t n, const double* template_state,
                                   const double* mean, const double* sqrt_s_v, double* out_state_soa, double* out_dispersion_soa) {
    if (!template_state || !sqrt_s_v || !out_state_soa) { set_err("null argument"); return NYXB_RC_BAD_ARG; }
    if (n == 0) return NYXB_RC_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_err("no CUDA device available: nyxb has no CPU fallback"); return NYXB_RC_NO_DEVICE; }
    if (device < 0 || device >= ndev) { set_err("bad device ordinal"); return NYXB_RC_BAD_ARG; }
    CUDA_TRY(cudaSetDevice(device));
    DevBufs B;
    double* d_out = B.alloc<double>(9 * n);
    double* d_disp = out_dispersion_soa ? B.alloc<double>(9 * n) : nullptr;
    if (!d_out || (out_dispersion_soa && !d_disp)) { set_err("device allocation failed"); return NYXB_RC_CUDA; }
    int32_t rc = nyxb_mvn_sample_dev(device, seed, first_index, n, template_state, mean, sqrt_s_v, d_out, d_disp, nullptr);
    if (rc != NYXB_RC_OK) return rc;
    CUDA_TRY(cudaMemcpy(out_state_soa, d_out, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost));
    if (d_disp) CUDA_TRY(cudaMemcpy(out_dispersion_soa, d_disp, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost));
    return NYXB_RC_OK;
}

// ---- batched Hermite resampling of recorded trajectories (nyxb_traj.cu)
extern "C" int32_t nyxb_traj_resample_dev(nyxb_engine* eng, size_t n, const nyxb_traj_sink* sink, size_t m, const int64_t* query_epoch_ns,
                                          double* out_state, int32_t* out_status, void* cuda_stream) {
    if (!eng || !sink || !query_epoch_ns || !out_state || !out_status) { set_err("null argument"); return NYXB_RC_BAD_ARG; }
    if (sink->capacity <= 0 || !sink->epoch_ns || !sink->state || !sink->count) { set_err("empty trajectory sink"); return NYXB_RC_BAD_ARG; }
    CUDA_TRY(cudaSetDevice(eng->device));
    cudaError_t err = nyxb_launch_traj_resample(sink->capacity, (const long long*)sink->epoch_ns, sink->state, (const long long*)sink->count,
                                                n, m, (const long long*)query_epoch_ns, out_state, out_status, (cudaStream_t)cuda_stream);
    if (err != cudaSuccess) { set_err(std::string("kernel launch: ") + cudaGetErrorString(err)); return NYXB_RC_CUDA; }
    if (n && m) eng->launches += 1;
    return NYXB_RC_OK;
}

// upload `sink` into the engine's slab (or check the resident recording when sink == NULL) and describe it with device pointers
static int32_t resident_sink(nyxb_engine* eng, size_t n, const nyxb_traj_sink* sink, cudaStream_t st, nyxb_traj_sink* dsink) {
    if (sink && (sink->capacity <= 0 || !sink->epoch_ns || !sink->state || !sink->count)) { set_err("empty trajectory sink"); return NYXB_RC_BAD_ARG; }
    if (!sink && (eng->rec_cap <= 0 || eng->rec_n != n || !eng->d_sink)) {
        set_err("no resident recording of this many trajectories: pass the sink of nyxb_propagate_batch_traj");
        return NYXB_RC_BAD_ARG;
    }
    if (sink) {   // [epoch cap*n i64 | state 6*cap*n f64 | count n i64]
        const size_t cap = (size_t)sink->capacity;
        const size_t need = (cap * n * 7 + n) * 8;
        if (need > eng->sink_bytes) {
            cudaFree(eng->d_sink); eng->d_sink = nullptr; eng->sink_bytes = 0; eng->rec_n = 0; eng->rec_cap = 0;
            CUDA_TRY(cudaMalloc(&eng->d_sink, need));
            eng->sink_bytes = need;
        }
        CUDA_TRY(cudaMemcpyAsync(eng->d_sink, sink->epoch_ns, cap * n * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(eng->d_sink + cap * n * 8, sink->state, cap * n * 48, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(eng->d_sink + cap * n * 56, sink->count, n * 8, cudaMemcpyHostToDevice, st));
        eng->rec_n = n; eng->rec_cap = sink->capacity;
    }
    const size_t cap = (size_t)eng->rec_cap;
    dsink->capacity = eng->rec_cap;
    dsink->epoch_ns = (int64_t*)eng->d_sink;
    dsink->state = (double*)(eng->d_sink + cap * n * 8);
    dsink->count = (int64_t*)(eng->d_sink + cap * n * 56);
    return NYXB_RC_OK;
}

extern "C" int32_t nyxb_traj_resample(nyxb_engine* eng, size_t n, const nyxb_traj_sink* sink, size_t m, const int64_t* query_epoch_ns,
                                      double* out_state, int32_t* out_status) {
    if (!eng || !query_epoch_ns || !out_state || !out_status) { set_err("null argument"); return NYXB_RC_BAD_ARG; }
    if (sink && (sink->capacity <= 0 || !sink->epoch_ns || !sink->state || !sink->count)) { set_err("empty trajectory sink"); return NYXB_RC_BAD_ARG; }
    if (!sink && (eng->rec_cap <= 0 || eng->rec_n != n || !eng->d_sink)) {
        set_err("no resident recording of this many trajectories: pass the sink of nyxb_propagate_batch_traj");
        return NYXB_RC_BAD_ARG;
    }
    if (n == 0 || m == 0) return NYXB_RC_OK;
    CUDA_TRY(cudaSetDevice(eng->device));
    if (!eng->stream) CUDA_TRY(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
    cudaStream_t st = eng->stream;
    nyxb_traj_sink dsink;
    int32_t rc = resident_sink(eng, n, sink, st, &dsink);
    if (rc != NYXB_RC_OK) return rc;
    DevBufs B;
    long long* d_q = B.put((const long long*)query_epoch_ns, m, st);
    double* d_out = B.alloc<double>(6 * m * n);
    int* d_status = B.alloc<int>(m * n);
    if (!d_q || !d_out || !d_status) { set_err("device allocation / upload failed"); return NYXB_RC_CUDA; }
    CUDA_TRY(cudaEventRecord(eng->ev0, st));
    rc = nyxb_traj_resample_dev(eng, n, &dsink, m, (const int64_t*)d_q, d_out, d_status, st);
    if (rc != NYXB_RC_OK) return rc;
    CUDA_TRY(cudaEventRecord(eng->ev1, st));
    CUDA_TRY(cudaMemcpyAsync(out_state, d_out, sizeof(double) * 6 * m * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_status, d_status, sizeof(int) * m * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, eng->ev0, eng->ev1) == cudaSuccess) eng->last_ms = ms;
    return NYXB_RC_OK;
}

// ---- event location on recorded trajectories (nyxb_traj.cu)
static bool event_kind_ok(int32_t kind) { return kind >= NYXB_EVENT_RMAG && kind <= NYXB_EVENT_VMAG; }

extern "C" int32_t nyxb_event_locate_dev(nyxb_engine* eng, size_t n, const nyxb_traj_sink* sink, int32_t kind, double value,
                                         int64_t epoch_precision_ns, const int32_t* run_status, int64_t* out_event_epoch_ns,
                                         double* out_event_state, int32_t* out_status, void* cuda_stream) {
    if (!eng || !sink || !out_event_epoch_ns || !out_event_state || !out_status) { set_err("null argument"); return NYXB_RC_BAD_ARG; }
    if (sink->capacity <= 0 || !sink->epoch_ns || !sink->state || !sink->count) { set_err("empty trajectory sink"); return NYXB_RC_BAD_ARG; }
    if (!event_kind_ok(kind) || epoch_precision_ns < 0) { set_err("bad event descriptor"); return NYXB_RC_BAD_ARG; }
    CUDA_TRY(cudaSetDevice(eng->device));
    cudaError_t err = nyxb_launch_event_locate(sink->capacity, (const long long*)sink->epoch_ns, sink->state, (const long long*)sink->count, n,
                                               kind, value, epoch_precision_ns, run_status, (long long*)out_event_epoch_ns, out_event_state,
                                               out_status, (cudaStream_t)cuda_stream);
    if (err != cudaSuccess) { set_err(std::string("kernel launch: ") + cudaGetErrorString(err)); return NYXB_RC_CUDA; }
    if (n) eng->launches += 1;
    return NYXB_RC_OK;
}

extern "C" int32_t nyxb_event_locate(nyxb_engine* eng, size_t n, const nyxb_traj_sink* sink, int32_t kind, double value,
                                     int64_t epoch_precision_ns, const int32_t* run_status, int64_t* out_event_epoch_ns,
                                     double* out_event_state, int32_t* out_status) {
    if (!eng || !out_event_epoch_ns || !out_event_state || !out_status) { set_err("null argument"); return NYXB_RC_BAD_ARG; }
    if (!event_kind_ok(kind) || epoch_precision_ns < 0) { set_err("bad event descriptor"); return NYXB_RC_BAD_ARG; }
    if (n == 0) return NYXB_RC_OK;
    CUDA_TRY(cudaSetDevice(eng->device));
    if (!eng->stream) CUDA_TRY(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
    cudaStream_t st = eng->stream;
    nyxb_traj_sink dsink;
    int32_t rc = resident_sink(eng, n, sink, st, &dsink);
    if (rc != NYXB_RC_OK) return rc;
    DevBufs B;
    int* d_run = run_status ? B.put((const int*)run_status, n, st) : nullptr;
    long long* d_ep = B.alloc<long long>(n);
    double* d_out = B.alloc<double>(6 * n);
    int* d_status = B.alloc<int>(n);
    if ((run_status && !d_run) || !d_ep || !d_out || !d_status) { set_err("device allocation / upload failed"); return NYXB_RC_CUDA; }
    CUDA_TRY(cudaEventRecord(eng->ev0, st));
    rc = nyxb_event_locate_dev(eng, n, &dsink, kind, value, epoch_precision_ns, d_run, (int64_t*)d_ep, d_out, d_status, st);
    if (rc != NYXB_RC_OK) return rc;
    CUDA_TRY(cudaEventRecord(eng->ev1, st));
    CUDA_TRY(cudaMemcpyAsync(out_event_epoch_ns, d_ep, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_event_state, d_out, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_status, d_status, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, eng->ev0, eng->ev1) == cudaSuccess) eng->last_ms = ms;
    return NYXB_RC_OK;
}

extern "C" int32_t nyxb_engine_set_lanes(nyxb_engine* eng, int32_t lanes) {
    if (!eng) return NYXB_RC_BAD_ARG;
    if (lanes != 0 && lanes != 1 && lanes != 8 && lanes != 16 && lanes != 32) { set_err("lanes must be 0,1,8,16,32"); return NYXB_RC_BAD_ARG; }
    if (lanes > 1 && !coop_supported(eng, lanes)) {
        set_err("cooperative lanes need a gravity field");
        return NYXB_RC_UNSUPPORTED;
    }
    eng->lanes = lanes;
    return NYXB_RC_OK;
}
extern "C" int32_t nyxb_engine_set_kernel(nyxb_engine* eng, int32_t kernel) {
    if (!eng) return NYXB_RC_BAD_ARG;
    if (kernel < NYXB_KERNEL_AUTO || kernel > NYXB_KERNEL_TRANSPOSED) { set_err("unknown kernel family"); return NYXB_RC_BAD_ARG; }
    if (kernel == NYXB_KERNEL_TRANSPOSED && !tx_supported(eng)) {
        set_err("the transposed kernel needs FAST mode and a gravity field of degree 8..70");
        return NYXB_RC_UNSUPPORTED;
    }
    if (kernel == NYXB_KERNEL_COOP && !eng->S.has_grav) { set_err("cooperative lanes need a gravity field"); return NYXB_RC_UNSUPPORTED; }
    eng->kernel = kernel;
    return NYXB_RC_OK;
}
extern "C" int32_t nyxb_engine_last_kernel(const nyxb_engine* eng) { return eng ? eng->last_kernel : 0; }
extern "C" int32_t nyxb_engine_set_tx_tuning(nyxb_engine* eng, int32_t slice_attempts, int32_t max_ctas) {
    if (!eng || slice_attempts < 1 || max_ctas < 0) { set_err("slice_attempts >= 1, max_ctas >= 0"); return NYXB_RC_BAD_ARG; }
    eng->tx_slice = slice_attempts;
    eng->tx_max_ctas = max_ctas;
    return NYXB_RC_OK;
}
extern "C" int32_t nyxb_engine_set_tx_positions(nyxb_engine* eng, int32_t positions) {
    if (!eng || (positions != 0 && positions != 8 && positions != 10 && positions != 16)) { set_err("positions: 0 (auto), 8, 10, 16"); return NYXB_RC_BAD_ARG; }
    eng->tx_positions = positions;
    return NYXB_RC_OK;
}
extern "C" int32_t nyxb_engine_get_lanes(const nyxb_engine* eng) { return eng ? pick_lanes(eng, 0) : 0; }
extern "C" int64_t nyxb_engine_launch_count(const nyxb_engine* eng) { return eng ? eng->launches : 0; }
extern "C" double nyxb_engine_last_kernel_ms(const nyxb_engine* eng) { return eng ? eng->last_ms : 0.0; }
extern "C" double nyxb_measure_fp64_tflops(int32_t device, int32_t iters) { return nyxb_fp64_probe(device, iters); }
extern "C" int32_t nyxb_coop_table_dump(const nyxb_gravity_field* f, int32_t lanes, int32_t* out_L, int32_t* out_kmax,
                                        double* recs, int32_t* col_start, int32_t* col_m, double* colseed) {
    if (!f || !f->c_nm || !f->s_nm || f->degree < 2 || (lanes != 8 && lanes != 16 && lanes != 32) || !out_L || !out_kmax) {
        set_err("bad argument");
        return NYXB_RC_BAD_ARG;
    }
    CoopHost h;
    nyxb_coop_build_host(f->degree, f->order, f->c_nm, f->s_nm, lanes, h);
    *out_L = h.L; *out_kmax = h.kmax;
    if (recs) std::copy(h.recs.begin(), h.recs.end(), recs);
    if (col_start) std::copy(h.col_start.begin(), h.col_start.end(), col_start);
    if (col_m) std::copy(h.col_m.begin(), h.col_m.end(), col_m);
    if (colseed) std::copy(h.colseed.begin(), h.colseed.end(), colseed);
    return NYXB_RC_OK;
}

extern "C" int32_t nyxb_tx_table_dump(const nyxb_gravity_field* f, int32_t positions, int32_t* out_n_rec, int32_t* out_kmax,
                                      double* recA, double* recK, double* colseed, int32_t* sched) {
    if (!f || !f->c_nm || !f->s_nm || f->degree < 2 || (positions != 8 && positions != 10 && positions != 16) || !out_n_rec || !out_kmax) {
        set_err("bad argument");
        return NYXB_RC_BAD_ARG;
    }
    TxHost h;
    nyxb_tx_build_host(f->degree, f->order, f->c_nm, f->s_nm, positions, h);
    *out_n_rec = h.n_rec; *out_kmax = h.kmax;
    if (recA) std::copy(h.recA.begin(), h.recA.end(), recA);
    if (recK) std::copy(h.recK.begin(), h.recK.end(), recK);
    if (colseed) std::copy(h.colseed.begin(), h.colseed.end(), colseed);
    if (sched) std::copy(h.sched.begin(), h.sched.end(), sched);
    return NYXB_RC_OK;
}

extern "C" int32_t nyxb_abi_version(void) { return NYXB_ABI_VERSION; }
extern "C" const char* nyxb_last_error(void) { return g_err.c_str(); }
