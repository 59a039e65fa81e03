/*
 * nyx_oracle_od.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as nyx_oracle.c).
 *
 * CPU restatement of the reference's state-transition-matrix path:
 *   SpacecraftDynamics::eom, `Some(stm)` branch      dynamics/spacecraft.rs:203-227
 *   SpacecraftDynamics::dual_eom                     dynamics/spacecraft.rs:312-363
 *   OrbitalDynamics::dual_eom                        dynamics/orbital.rs:116-172
 *   PointMasses::gradient                            dynamics/orbital.rs:249-307
 *   GravityField::gradient                           dynamics/gravity_field.rs:273-431
 *   SolarPressure::gradient                          dynamics/solarpressure.rs:167-233
 *   PropInstance::{propagate, single_step, derive} on the 90-vector   propagators/instance.rs:87-262, 343-493
 * plus a small stateful `PropInstance` handle so that the Kalman-filter loop (od/process/mod.rs:211-426,
 * restated in numpy in oracle/pyoracle_od.py) can drive it one `for_duration` at a time.
 *
 * The reference obtains the partials with forward-mode dual numbers (`hyperdual = 1.5.0`, not in the tree).  The
 * same is done here with a 3-partial dual type: only the partials with respect to position are ever read
 * (orbital.rs:148-156, 296-301; gravity_field.rs:420-428; solarpressure.rs:216-222).  The operator formulas are the
 * published ones of that crate restated from memory (product rule etc.): PARITY UNPINNED at the rounding level —
 * no reference test asserts an STM value (tests/propagation/stm.rs compares against finite differences with a 1 km
 * tolerance); tests/test_oracle_stm.py checks this file against central finite differences of nyx_oracle.c.
 *
 * As coded in the reference, the stage derivative of the STM block is `ctx.stm * grad` where `ctx` is the state at
 * the START of the step (spacecraft.rs:203-214 matches on `ctx.stm`, instance.rs:364 passes `&self.state`), i.e.
 * Phi_{k+1} = Phi_k + Phi_k * sum_i (h b_i) A_i — restated as is.
 */
#include "nyx_oracle_priv.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* dual numbers with 3 partials (hyperdual::OHyperdual restricted to d/dx,y,z) */
/* ------------------------------------------------------------------------- */
typedef struct { double v, d[3]; } d3;

static inline d3 d3c(double v) { d3 r = { v, {0.0, 0.0, 0.0} }; return r; }
static inline d3 d3var(double v, int i) { d3 r = { v, {0.0, 0.0, 0.0} }; r.d[i] = 1.0; return r; }
static inline d3 d3add(d3 a, d3 b) { d3 r; r.v = a.v + b.v; for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
static inline d3 d3sub(d3 a, d3 b) { d3 r; r.v = a.v - b.v; for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
/* Mul: dual_i = rhs.re * self_i + self.re * rhs_i */
static inline d3 d3mul(d3 a, d3 b) { d3 r; r.v = a.v * b.v; for (int i = 0; i < 3; ++i) r.d[i] = b.v * a.d[i] + a.v * b.d[i]; return r; }
/* Div: dual_i = (rhs.re * self_i - self.re * rhs_i) / rhs.re^2 */
static inline d3 d3div(d3 a, d3 b) {
    d3 r; double den = b.v * b.v;
    r.v = a.v / b.v;
    for (int i = 0; i < 3; ++i) r.d[i] = (b.v * a.d[i] - a.v * b.d[i]) / den;
    return r;
}
static inline d3 d3scale(d3 a, double c) { d3 r; r.v = a.v * c; for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] * c; return r; }
static inline d3 d3divs(d3 a, double c) { d3 r; r.v = a.v / c; for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] / c; return r; }
static inline double powi_f(double x, int n) { double r = 1.0; for (int i = 0; i < n; ++i) r *= x; return r; }
/* powi(n): real^n, dual_i = n * real^(n-1) * self_i */
static inline d3 d3powi(d3 a, int n) {
    d3 r; double nf = (double)n, p = powi_f(a.v, n - 1);
    r.v = powi_f(a.v, n);
    for (int i = 0; i < 3; ++i) r.d[i] = nf * p * a.d[i];
    return r;
}
static inline d3 d3sqrt(d3 a) {
    d3 r; r.v = sqrt(a.v);
    double dd = 1.0 / (2.0 * r.v);
    for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] * dd;
    return r;
}
/* hyperdual::linalg::norm: sqrt(sum_i v_i.powi(2)) */
static inline d3 d3norm(const d3 v[3]) {
    d3 s = d3c(0.0);
    for (int i = 0; i < 3; ++i) s = d3add(s, d3powi(v[i], 2));
    return d3sqrt(s);
}

static inline double norm3(const double v[3]) { return sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }

/* ------------------------------------------------------------------------- */
/* GravityField::gradient — gravity_field.rs:273-431                          */
/* scratch: (N+3)^2 d3                                                        */
/* ------------------------------------------------------------------------- */
static void grav_gradient(const nyx_oracle_grav* h, int64_t t_ns, const double r_in[3], d3* a, double acc[3], double grad[9]) {
    int N = h->n, M = h->m, dim = h->dim, np2 = N + 2;
    double R[9], wdot;
    nyx_oracle_rotation(&h->rot, t_ns, R, &wdot);
    double rb[3];
    for (int i = 0; i < 3; ++i) rb[i] = (R[3 * i] * r_in[0] + R[3 * i + 1] * r_in[1]) + R[3 * i + 2] * r_in[2];
    d3 radius[3] = { d3var(rb[0], 0), d3var(rb[1], 1), d3var(rb[2], 2) };     /* :287 */
    d3 r_ = d3norm(radius);                                                   /* :290 */
    d3 s_ = d3div(radius[0], r_), t_ = d3div(radius[1], r_), u_ = d3div(radius[2], r_);
    for (int k = 0; k < dim * dim; ++k) a[k] = d3c(0.0);                      /* :297 */
    for (int k = 0; k <= N + 1; ++k) a[k * dim + k] = d3c(h->a_diag[k]);      /* :299-301 */
    a[1 * dim + 0] = d3scale(u_, sqrt(3.0));                                  /* :304 */
    for (int n = 1; n <= N + 1; ++n) {                                        /* :305-309 */
        double nf = (double)n;
        a[(n + 1) * dim + n] = d3mul(d3mul(d3c(sqrt(2.0 * nf + 3.0)), u_), a[n * dim + n]);
    }
    for (int m = 0; m <= M + 1; ++m)                                          /* :311-317 */
        for (int n = m + 2; n <= N + 1; ++n)
            a[n * dim + m] = d3sub(d3mul(d3mul(u_, d3c(h->b_nm[n * np2 + m])), a[(n - 1) * dim + m]),
                                   d3mul(d3c(h->c_nm[n * np2 + m]), a[(n - 2) * dim + m]));
    int mm = N < M ? N : M;                                                   /* :320-329 */
    d3 r_m[mm + 2], i_m[mm + 2];
    r_m[0] = d3c(1.0); i_m[0] = d3c(0.0);
    for (int m = 1; m <= mm; ++m) {
        r_m[m] = d3sub(d3mul(s_, r_m[m - 1]), d3mul(t_, i_m[m - 1]));
        i_m[m] = d3add(d3mul(s_, i_m[m - 1]), d3mul(t_, r_m[m - 1]));
    }
    d3 eq_radius = d3c(h->r_eq);                                              /* :345-347 */
    d3 rho = d3div(eq_radius, r_);
    d3 rho_np1 = d3mul(d3div(d3c(h->mu), r_), rho);
    d3 a0 = d3c(0.0), a1 = d3c(0.0), a2 = d3c(0.0), a3 = d3c(0.0);
    d3 sqrt2 = d3c(sqrt(2.0));
    for (int n = 1; n <= N; ++n) {                                            /* :355-404 */
        d3 sum0 = d3c(0.0), sum1 = d3c(0.0), sum2 = d3c(0.0), sum3 = d3c(0.0);
        rho_np1 = d3mul(rho_np1, rho);
        int mtop = n < M ? n : M;
        for (int m = 0; m <= mtop; ++m) {
            d3 cv = d3c(h->cbar[n * (N + 1) + m]), sv = d3c(h->sbar[n * (N + 1) + m]);
            d3 d_ = d3mul(d3add(d3mul(cv, r_m[m]), d3mul(sv, i_m[m])), sqrt2);
            d3 e_ = d3c(0.0), f_ = d3c(0.0);
            if (m != 0) {
                e_ = d3mul(d3add(d3mul(cv, r_m[m - 1]), d3mul(sv, i_m[m - 1])), sqrt2);
                f_ = d3mul(d3sub(d3mul(sv, r_m[m - 1]), d3mul(cv, i_m[m - 1])), sqrt2);
            }
            d3 mf = d3c((double)m);
            sum0 = d3add(sum0, d3mul(d3mul(mf, a[n * dim + m]), e_));
            sum1 = d3add(sum1, d3mul(d3mul(mf, a[n * dim + m]), f_));
            sum2 = d3add(sum2, d3mul(d3mul(d3c(h->vr01[n * np2 + m]), a[n * dim + m + 1]), d_));
            sum3 = d3add(sum3, d3mul(d3mul(d3c(h->vr11[n * np2 + m]), a[(n + 1) * dim + m + 1]), d_));
        }
        d3 rr = d3div(rho_np1, eq_radius);
        a0 = d3add(a0, d3mul(rr, sum0));
        a1 = d3add(a1, d3mul(rr, sum1));
        a2 = d3add(a2, d3mul(rr, sum2));
        a3 = d3sub(a3, d3mul(rr, sum3));
    }
    d3 al[3] = { d3add(a0, d3mul(a3, s_)), d3add(a1, d3mul(a3, t_)), d3add(a2, d3mul(a3, u_)) };  /* :416 */
    /* :417-430 dx = dcm * real, grad = dcm * grad_local * dcm^T with dcm = body-fixed -> inertial = R^T */
    for (int i = 0; i < 3; ++i) acc[i] = (R[i] * al[0].v + R[3 + i] * al[1].v) + R[6 + i] * al[2].v;
    double tmp[9]; /* R^T G */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            tmp[3 * i + j] = (R[i] * al[0].d[j] + R[3 + i] * al[1].d[j]) + R[6 + i] * al[2].d[j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            grad[3 * i + j] = (tmp[3 * i] * R[j] + tmp[3 * i + 1] * R[3 + j]) + tmp[3 * i + 2] * R[6 + j];
}

/* ------------------------------------------------------------------------- */
/* dual_eom — spacecraft.rs:312-363 and what it calls                         */
/* ------------------------------------------------------------------------- */
typedef struct {
    const nyxb_dynamics* dyn;
    nyx_oracle_grav* grav;
    d3* grav_scratch;
    double dry_mass, extra_mass, srp_area, drag_area;
    int64_t n_rhs;
} od_ctx;

#define AU_KM 149597870.700
#define SPEED_OF_LIGHT_M_S (299792.458 * 1e3)

/* y[9] (osculating: Cr clamped by the caller) at absolute time t_ns -> dx[9], grad[81] row-major 9x9 */
static int dual_eom(od_ctx* cx, int64_t t_ns, const double y[9], double dx[9], double grad[81]) {
    const nyxb_dynamics* dyn = cx->dyn;
    if (dyn->drag) return NYXB_ERR_PROP_MATH; /* PartialsUndefined (drag.rs:109-118, 286-295); rejected earlier */
    memset(grad, 0, sizeof(double) * 81);
    for (int e = 0; e < 9; ++e) dx[e] = 0.0;
    const double* r = y;
    /* ---- OrbitalDynamics::dual_eom, orbital.rs:116-172 */
    d3 radius[3] = { d3var(r[0], 0), d3var(r[1], 1), d3var(r[2], 2) };
    d3 rmag = d3norm(radius);
    d3 fac = d3div(d3c(-dyn->mu_central_km3_s2), d3powi(rmag, 3));
    for (int i = 0; i < 3; ++i) {
        d3 ba = d3mul(radius[i], fac);
        dx[i] = y[3 + i];
        dx[3 + i] = ba.v;
        grad[i * 9 + 3 + i] = 1.0;                     /* velocity[i][j]: d(v_i)/d(v_i) */
        for (int j = 0; j < 3; ++j) grad[(3 + i) * 9 + j] = ba.d[j];
    }
    double bpos[NYXB_MAX_BODIES][3];
    for (int j = 0; j < dyn->n_bodies; ++j)
        if (nyx_oracle_body_position(&dyn->bodies[j], t_ns, bpos[j])) return NYXB_ERR_EPHEMERIS;
    /* ---- PointMasses::gradient, orbital.rs:249-307 (r_ij carries identity partials, as coded) */
    if (dyn->point_mass_mask) {   /* NB: ascending body index; the STM path takes one PointMasses list in almanac order */
        double fx[3] = {0, 0, 0}, g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = 0; j < dyn->n_bodies; ++j) {
            if (!((dyn->point_mass_mask >> j) & 1u)) continue;
            d3 gm_d = d3c(-dyn->bodies[j].mu_km3_s2);
            d3 r_ij[3] = { d3var(bpos[j][0], 0), d3var(bpos[j][1], 1), d3var(bpos[j][2], 2) };
            d3 r_ij3 = d3powi(d3norm(r_ij), 3);
            d3 r_j[3];
            for (int i = 0; i < 3; ++i) { r_j[i] = d3sub(radius[i], r_ij[i]); r_j[i].d[i] = 1.0; }  /* :287-291 */
            d3 r_j3 = d3powi(d3norm(r_j), 3);
            for (int i = 0; i < 3; ++i) {
                d3 t = d3add(d3div(r_j[i], r_j3), d3div(r_ij[i], r_ij3));
                t = d3mul(t, gm_d);
                fx[i] += t.v;
                for (int q = 0; q < 3; ++q) g[3 * i + q] += t.d[q];
            }
        }
        for (int i = 0; i < 3; ++i) {                  /* orbital.rs:159-167 */
            dx[3 + i] += fx[i];
            for (int q = 0; q < 3; ++q) grad[(3 + i) * 9 + q] += g[3 * i + q];
        }
    }
    if (cx->grav) {
        double ga[3], gg[9];
        grav_gradient(cx->grav, t_ns, r, cx->grav_scratch, ga, gg);
        for (int i = 0; i < 3; ++i) {
            dx[3 + i] += ga[i];
            for (int q = 0; q < 3; ++q) grad[(3 + i) * 9 + q] += gg[3 * i + q];
        }
    }
    /* ---- force models, spacecraft.rs:340-360 */
    if (dyn->srp) {
        /* SolarPressure::gradient, solarpressure.rs:167-233 */
        const nyxb_srp* sp = dyn->srp;
        double total_mass = cx->dry_mass + y[8] + cx->extra_mass;
        double cr = y[6];
        const double* sun = bpos[sp->sun_body];
        double r_sun[3] = { r[0] - sun[0], r[1] - sun[1], r[2] - sun[2] };
        d3 r_sun_d[3] = { d3var(r_sun[0], 0), d3var(r_sun[1], 1), d3var(r_sun[2], 2) };
        d3 n_d = d3norm(r_sun_d);
        d3 unit[3] = { d3div(r_sun_d[0], n_d), d3div(r_sun_d[1], n_d), d3div(r_sun_d[2], n_d) };
        double occult = 0.0;
        double r_ls[3] = { -r_sun[0], -r_sun[1], -r_sun[2] };
        for (int q = 0; q < sp->n_shadow; ++q) {
            int bi = sp->shadow_body[q];
            double r_eb[3], rad;
            if (bi == NYXB_CENTRAL_BODY) { r_eb[0] = r[0]; r_eb[1] = r[1]; r_eb[2] = r[2]; rad = dyn->central_radius_km; }
            else { r_eb[0] = r[0] - bpos[bi][0]; r_eb[1] = r[1] - bpos[bi][1]; r_eb[2] = r[2] - bpos[bi][2]; rad = dyn->bodies[bi].radius_km; }
            double p = nyx_oracle_occultation(r_eb, r_ls, dyn->bodies[sp->sun_body].radius_km, rad);
            if (p > occult) occult = p;
        }
        double k = fabs(occult - 1.0);
        d3 r_sun_au = d3divs(n_d, AU_KM);
        d3 inv = d3div(d3c(1.0), r_sun_au);
        d3 flux = d3mul(d3c(k * sp->phi_w_m2 / SPEED_OF_LIGHT_M_S), d3powi(inv, 2));
        d3 scal = d3c(1e-3 * cr * cx->srp_area);
        /* eom (solarpressure.rs:135-165) for the Cr partial: wrt_cr = eom / Cr */
        double n_sun = norm3(r_sun);
        double r_au = n_sun / AU_KM, inv_s = 1.0 / r_au;
        double flux_s = (k * sp->phi_w_m2 / SPEED_OF_LIGHT_M_S) * (inv_s * inv_s);
        double scal_s = 1e-3 * cr * cx->srp_area * flux_s;
        for (int i = 0; i < 3; ++i) {
            d3 f = d3mul(d3mul(scal, flux), unit[i]);
            dx[3 + i] += f.v / total_mass;
            for (int q = 0; q < 3; ++q) grad[(3 + i) * 9 + q] += f.d[q] / total_mass;
            if (sp->estimate) {
                double wrt_cr = (scal_s * (r_sun[i] / n_sun)) / cr;
                grad[(3 + i) * 9 + 6] += wrt_cr / total_mass;
            }
        }
    }
    return 0;
}

/* SpacecraftDynamics::eom on the 90-vector with ctx.stm = Some (spacecraft.rs:191-227).
 * ctx_stm: the 9x9 STM of the context state, column-major like the vector tail. */
static int eom90(od_ctx* cx, int64_t epoch_ns, double delta_t_s, const double yv[90], const double ctx_stm[81], double dy[90]) {
    cx->n_rhs++;
    int64_t t_ns = epoch_ns + nyx_oracle_dur_from_seconds(delta_t_s);
    double y[9];
    for (int e = 0; e < 9; ++e) y[e] = yv[e];
    y[6] = y[6] < 0.0 ? 0.0 : (y[6] > 2.0 ? 2.0 : y[6]);           /* cosmic/spacecraft.rs:494 */
    if (cx->dyn->srp) {
        double mass = cx->dry_mass + y[8] + cx->extra_mass;
        if (!(mass > 0.0)) return NYXB_ERR_MASSLESS;                /* spacecraft.rs:201-203 */
    }
    double dx[9], grad[81];
    int rc = dual_eom(cx, t_ns, y, dx, grad);
    if (rc) return rc;
    for (int e = 0; e < 9; ++e) dy[e] = dx[e];
    /* stm_dt = stm * grad (spacecraft.rs:213), written column-major (:220-222) */
    for (int c = 0; c < 9; ++c)
        for (int r = 0; r < 9; ++r) {
            double s = 0.0;
            for (int k = 0; k < 9; ++k) s += ctx_stm[k * 9 + r] * grad[k * 9 + c];
            dy[9 + c * 9 + r] = s;
        }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* PropInstance on the 90-vector — propagators/instance.rs                     */
/* ------------------------------------------------------------------------- */
struct nyx_oracle_inst {
    nyxb_dynamics dyn;
    nyxb_integ_opts opts;
    int order, stages;
    const double *ta, *tb;
    od_ctx cx;
    double y[90];
    int64_t epoch_ns, step_ns;
    int fixed, status;
    nyxb_details det;
};

#define MAX_STAGES 16
#define VL 90

static int derive90(nyx_oracle_inst* in, int64_t* dt_ns, double next[VL]) {
    static _Thread_local double k[MAX_STAGES][VL];
    const nyxb_integ_opts* o = &in->opts;
    const double* y = in->y;
    int S = in->stages;
    in->det.attempts = 1;
    double h = nyx_oracle_dur_to_seconds(in->step_ns);
    const double min_s = nyx_oracle_dur_to_seconds(o->min_step_ns);
    const double max_s = nyx_oracle_dur_to_seconds(o->max_step_ns);
    const double* ctx_stm = in->y + 9;                              /* instance.rs:364 state_ctx = &self.state */
    for (;;) {
        int rc = eom90(&in->cx, in->epoch_ns, 0.0, y, ctx_stm, k[0]);
        if (rc) return rc;
        int a_idx = 0;
        for (int i = 0; i < S - 1; ++i) {
            double ci = 0.0, wi[VL];
            for (int e = 0; e < VL; ++e) wi[e] = 0.0;
            for (int j = 0; j <= i; ++j) {
                double a_ij = in->ta[a_idx++];
                ci += a_ij;
                for (int e = 0; e < VL; ++e) wi[e] += a_ij * k[j][e];
            }
            double ys[VL];
            for (int e = 0; e < VL; ++e) ys[e] = y[e] + h * wi[e];
            rc = eom90(&in->cx, in->epoch_ns, ci * h, ys, ctx_stm, k[i + 1]);
            if (rc) return rc;
        }
        double err_est[VL];
        for (int e = 0; e < VL; ++e) { err_est[e] = 0.0; next[e] = y[e]; }
        for (int i = 0; i < S; ++i) {
            double b_i = in->tb[i];
            if (!in->fixed) {
                double cf = h * (b_i - in->tb[i + S]);
                for (int e = 0; e < VL; ++e) err_est[e] += cf * k[i][e];
            }
            double cb = h * b_i;
            for (int e = 0; e < VL; ++e) next[e] += cb * k[i][e];
        }
        if (in->fixed) {
            in->det.step_ns = in->step_ns;
            *dt_ns = in->step_ns;
            return 0;
        }
        /* only the Cartesian controls are accepted: they read components 0..5 (error_ctrl.rs:89-122) */
        in->det.error = nyx_oracle_error_estimate(o->error_ctrl, err_est, next, y);
        if (nyx_oracle_error_scale_ != 1.0) in->det.error *= nyx_oracle_error_scale_;   /* sensitivity probe only (nyx_oracle.c) */
        if (in->det.error <= o->tolerance || h <= min_s || in->det.attempts >= o->attempts) {
            for (int e = 0; e < VL; ++e)
                if (next[e] != next[e]) return NYXB_ERR_PROP_MATH;
            if (in->det.attempts >= o->attempts) in->status |= NYXB_WARN_MAX_ATTEMPTS;
            in->det.step_ns = nyx_oracle_dur_from_seconds(h);
            if (in->det.error < o->tolerance) {
                double proposed = 0.9 * h * pow(o->tolerance / in->det.error, 1.0 / (double)in->order);
                if (fabs(proposed) > fabs(max_s)) {
                    double sg = (proposed != proposed) ? proposed : (signbit(proposed) ? -1.0 : 1.0);
                    h = max_s * sg;
                } else {
                    h = proposed;
                }
            }
            in->step_ns = nyx_oracle_dur_from_seconds(h);
            int64_t ab = in->step_ns < 0 ? -in->step_ns : in->step_ns;
            if (ab < o->min_step_ns) in->step_ns = (in->step_ns < 0) ? -o->min_step_ns : o->min_step_ns;
            *dt_ns = in->det.step_ns;
            return 0;
        }
        in->det.attempts += 1;
        in->det.n_rejected += 1;
        double proposed = 0.9 * h * pow(o->tolerance / in->det.error, 1.0 / (double)(in->order - 1));
        h = (proposed < min_s) ? min_s : proposed;
    }
}

static int single_step90(nyx_oracle_inst* in) {
    int64_t dt; double next[VL];
    int rc = derive90(in, &dt, next);
    if (rc) return rc;
    in->epoch_ns += dt;
    for (int e = 0; e < VL; ++e) in->y[e] = next[e];
    in->y[6] = in->y[6] < 0.0 ? 0.0 : (in->y[6] > 2.0 ? 2.0 : in->y[6]);
    in->det.n_steps += 1;
    return (in->y[8] < 0.0) ? NYXB_ERR_FUEL_EXHAUSTED : 0;
}

static int propagate90(nyx_oracle_inst* in, int64_t duration_ns) {
    if (duration_ns == 0) return 0;
    int64_t stop = in->epoch_ns + duration_ns;
    if (in->y[8] < 0.0) return NYXB_ERR_FUEL_EXHAUSTED;
    int backprop = duration_ns < 0;
    if (backprop) in->step_ns = -in->step_ns;
    for (;;) {
        int64_t epoch = in->epoch_ns;
        if ((!backprop && epoch + in->step_ns > stop) || (backprop && epoch + in->step_ns <= stop)) {
            if (stop == epoch) return 0;
            int64_t prev_step = in->step_ns; int prev_fixed = in->fixed;
            in->step_ns = stop - epoch; in->fixed = 1;
            int rc = single_step90(in);
            if (rc) return rc;
            in->step_ns = prev_step; in->fixed = prev_fixed;
            if (backprop) in->step_ns = -in->step_ns;
            return 0;
        }
        int rc = single_step90(in);
        if (rc) return rc;
    }
}

/* ---- handle API (used by oracle/pyoracle_od.py) ---- */
nyx_oracle_inst* nyx_oracle_inst_new(const nyxb_dynamics* dyn, const nyxb_integ_opts* opts, const double y9[9],
                                     const double consts[4], int64_t epoch_ns) {
    if (dyn->drag) return NULL;
    if (dyn->n_gravity > 1 || (dyn->gravity && dyn->gravity->body != NYXB_CENTRAL_BODY) || opts->state_center) return NULL;   /* one central field, no frame swap */
    if (!opts->fixed_step && opts->error_ctrl != NYXB_RSS_CARTESIAN_STATE && opts->error_ctrl != NYXB_RSS_CARTESIAN_STEP) return NULL;
    nyx_oracle_inst* in = (nyx_oracle_inst*)calloc(1, sizeof(*in));
    in->dyn = *dyn; in->opts = *opts;
    if (nyx_oracle_tableau(opts->method, &in->order, &in->stages, &in->ta, &in->tb)) { free(in); return NULL; }
    in->cx.dyn = &in->dyn;
    in->cx.grav = dyn->gravity ? nyx_oracle_grav_new(dyn->gravity) : NULL;
    in->cx.grav_scratch = in->cx.grav ? (d3*)malloc(sizeof(d3) * (size_t)in->cx.grav->dim * in->cx.grav->dim) : NULL;
    in->cx.dry_mass = consts[0]; in->cx.extra_mass = consts[1]; in->cx.srp_area = consts[2]; in->cx.drag_area = consts[3];
    for (int e = 0; e < 9; ++e) in->y[e] = y9[e];
    for (int c = 0; c < 9; ++c) in->y[9 + c * 9 + c] = 1.0;         /* with_stm(): identity */
    in->epoch_ns = epoch_ns;
    in->step_ns = opts->init_step_ns; in->fixed = opts->fixed_step;
    in->det.step_ns = opts->init_step_ns; in->det.attempts = 1;
    return in;
}
void nyx_oracle_inst_free(nyx_oracle_inst* in) {
    if (!in) return;
    nyx_oracle_grav_free(in->cx.grav);
    free(in->cx.grav_scratch);
    free(in);
}
int nyx_oracle_inst_for_duration(nyx_oracle_inst* in, int64_t duration_ns) { return propagate90(in, duration_ns); }
void nyx_oracle_inst_get(const nyx_oracle_inst* in, double y[90], int64_t* epoch_ns, int64_t* step_ns, int* fixed, nyxb_details* det) {
    memcpy(y, in->y, sizeof(double) * 90);
    *epoch_ns = in->epoch_ns; *step_ns = in->step_ns; *fixed = in->fixed;
    if (det) { *det = in->det; det->n_rhs = in->cx.n_rhs; }
}
void nyx_oracle_inst_set(nyx_oracle_inst* in, const double y[90], int64_t epoch_ns) {
    memcpy(in->y, y, sizeof(double) * 90);
    in->epoch_ns = epoch_ns;
}
void nyx_oracle_inst_set_step(nyx_oracle_inst* in, int64_t step_ns, int fixed) { in->step_ns = step_ns; in->fixed = fixed; }

/* one evaluation of dual_eom for unit tests: y[9], consts[4] -> dx[9], grad[81] row-major */
int nyx_oracle_dual_eom(const nyxb_dynamics* dyn, int64_t t_ns, const double y[9], const double consts[4], double dx[9], double grad[81]) {
    nyxb_integ_opts o; memset(&o, 0, sizeof(o)); o.method = NYXB_RK4; o.fixed_step = 1; o.init_step_ns = 1;
    nyx_oracle_inst* in = nyx_oracle_inst_new(dyn, &o, y, consts, t_ns);
    if (!in) return -1;
    double yy[9];
    for (int e = 0; e < 9; ++e) yy[e] = y[e];
    yy[6] = yy[6] < 0.0 ? 0.0 : (yy[6] > 2.0 ? 2.0 : yy[6]);
    int rc = dual_eom(&in->cx, t_ns, yy, dx, grad);
    nyx_oracle_inst_free(in);
    return rc;
}

/* Batch STM propagation: same contract as nyxb_propagate_batch_stm (include/nyxb.h). */
int nyx_oracle_propagate_batch_stm(const nyxb_dynamics* dyn, const nyxb_integ_opts* opts, size_t n,
                                   const double* state_soa, const double* consts_soa,
                                   const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                                   const double* stm_in_soa, double* out_state_soa, int64_t* out_epoch_ns,
                                   double* out_stm_soa, nyxb_details* out_details, int32_t* out_status, int n_threads) {
    if (dyn->drag) return -1;
    if (!opts->fixed_step && opts->error_ctrl != NYXB_RSS_CARTESIAN_STATE && opts->error_ctrl != NYXB_RSS_CARTESIAN_STEP) return -1;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (long long ii = 0; ii < (long long)n; ++ii) {
        size_t i = (size_t)ii;
        double y9[9], cs[4];
        for (int e = 0; e < 9; ++e) y9[e] = state_soa[(size_t)e * n + i];
        for (int e = 0; e < 4; ++e) cs[e] = consts_soa[(size_t)e * n + i];
        nyx_oracle_inst* in = nyx_oracle_inst_new(dyn, opts, y9, cs, epoch0_ns[i]);
        if (stm_in_soa)
            for (int e = 0; e < 81; ++e) in->y[9 + e] = stm_in_soa[(size_t)e * n + i];
        if (step_ns) in->step_ns = step_ns[i];
        int rc = propagate90(in, end_epoch_ns - in->epoch_ns);
        for (int e = 0; e < 9; ++e) out_state_soa[(size_t)e * n + i] = in->y[e];
        for (int e = 0; e < 81; ++e) out_stm_soa[(size_t)e * n + i] = in->y[9 + e];
        out_epoch_ns[i] = in->epoch_ns;
        if (step_ns) step_ns[i] = in->step_ns;
        in->det.n_rhs = in->cx.n_rhs;
        if (out_details) out_details[i] = in->det;
        out_status[i] = (in->status & NYXB_WARN_MAX_ATTEMPTS) | rc;
        nyx_oracle_inst_free(in);
    }
    return 0;
}
