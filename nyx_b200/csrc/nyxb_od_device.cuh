// nyxb_od_device.cuh — device code shared by the per-thread (nyxb_od.cu) and the warp-cooperative (nyxb_od_coop.cu)
// STM / filter kernels: 3-partial dual numbers, the dual right-hand side (spacecraft.rs:312-363 and what it calls) and
// the tracking geometry (trk_device.rs:150-200).
#pragma once
#include "nyxb_od.cuh"

// ------------------------------------------------------------------------- dual numbers (value + d/dx, d/dy, d/dz)
struct D3 { double v, x, y, z; };
__device__ __forceinline__ D3 dc(double v) { return D3{v, 0.0, 0.0, 0.0}; }
__device__ __forceinline__ D3 dvar(double v, int i) { return D3{v, i == 0 ? 1.0 : 0.0, i == 1 ? 1.0 : 0.0, i == 2 ? 1.0 : 0.0}; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return D3{a.v + b.v, a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return D3{a.v - b.v, a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 operator*(D3 a, D3 b) {
    return D3{a.v * b.v, b.v * a.x + a.v * b.x, b.v * a.y + a.v * b.y, b.v * a.z + a.v * b.z};
}
__device__ __forceinline__ D3 operator/(D3 a, D3 b) {
#if NYXB_STRICT
    double den = b.v * b.v;
    return D3{a.v / b.v, (b.v * a.x - a.v * b.x) / den, (b.v * a.y - a.v * b.y) / den, (b.v * a.z - a.v * b.z) / den};
#else
    const double r = 1.0 / b.v, q = a.v * r;   // FAST: one division; d(a/b) = (da - (a/b) db) / b
    return D3{q, (a.x - q * b.x) * r, (a.y - q * b.y) * r, (a.z - q * b.z) * r};
#endif
}
__device__ __forceinline__ D3 dscale(D3 a, double c) { return D3{a.v * c, a.x * c, a.y * c, a.z * c}; }
__device__ __forceinline__ D3 ddivs(D3 a, double c) { return D3{a.v / c, a.x / c, a.y / c, a.z / c}; }
__device__ __forceinline__ D3 dsq(D3 a) { double p = 2.0 * a.v; return D3{a.v * a.v, p * a.x, p * a.y, p * a.z}; }
__device__ __forceinline__ D3 dcube(D3 a) { double p = 3.0 * (a.v * a.v); return D3{(a.v * a.v) * a.v, p * a.x, p * a.y, p * a.z}; }
__device__ __forceinline__ D3 dsqrt(D3 a) {
    double r = sqrt(a.v), dd = 1.0 / (2.0 * r);
    return D3{r, a.x * dd, a.y * dd, a.z * dd};
}
__device__ __forceinline__ D3 dnorm(D3 a, D3 b, D3 c) { return dsqrt(((dc(0.0) + dsq(a)) + dsq(b)) + dsq(c)); }
__device__ __forceinline__ double dpart(const D3& a, int j) { return j == 0 ? a.x : (j == 1 ? a.y : a.z); }

// ------------------------------------------------------------------------- GravityField::gradient (gravity_field.rs:273-431)
// rolling rows of the derived-Legendre triangle in dual numbers: P = row n, Q = row n-1 -> row n+1
__device__ static void grav_gradient(const DevGrav& g, long long t_ns, const double r_in[3], double acc[3], double G[9]) {
    const int N = g.N, M = g.M;
    double R[9];
    rotation_dcm(g.rot, t_ns, R);
    double rb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) rb[i] = (R[3 * i] * r_in[0] + R[3 * i + 1] * r_in[1]) + R[3 * i + 2] * r_in[2];
    D3 rx = dvar(rb[0], 0), ry = dvar(rb[1], 1), rz = dvar(rb[2], 2);
    D3 r_ = dnorm(rx, ry, rz);
    D3 s_ = rx / r_, t_ = ry / r_, u_ = rz / r_;
    D3 rowA[NYXB_MAX_DEGREE + 3], rowB[NYXB_MAX_DEGREE + 3];
    D3 r_m[NYXB_MAX_DEGREE + 1], i_m[NYXB_MAX_DEGREE + 1];
    D3* P = rowA;
    D3* Q = rowB;
    for (int m = 0; m <= N + 2; ++m) { rowA[m] = dc(0.0); rowB[m] = dc(0.0); }
    Q[0] = dc(1.0);
    P[0] = dscale(u_, sqrt(3.0));
    P[1] = dc(__ldg(g.a_diag + 1));
    const int mm = N < M ? N : M;
    r_m[0] = dc(1.0); i_m[0] = dc(0.0);
    for (int m = 1; m <= mm; ++m) {
        r_m[m] = s_ * r_m[m - 1] - t_ * i_m[m - 1];
        i_m[m] = s_ * i_m[m - 1] + t_ * r_m[m - 1];
    }
    D3 eq_radius = dc(g.r_eq);
    D3 rho = eq_radius / r_;
    D3 rho_np1 = (dc(g.mu) / r_) * rho;
    D3 a0 = dc(0.0), a1 = dc(0.0), a2 = dc(0.0), a3 = dc(0.0);
    const D3 sqrt2 = dc(sqrt(2.0));
    for (int n = 1; n <= N; ++n) {
        {   // row n+1 into Q (holds row n-1): gravity_field.rs:305-317
            const int np1 = n + 1;
            const DevHarm* trow = g.tab + tri(np1, 0);
            int mrec = np1 - 2;
            if (mrec > M + 1) mrec = M + 1;
            for (int m = 0; m <= mrec; ++m) {
                double bb = __ldg(&trow[m].b), cc = __ldg(&trow[m].c);
                Q[m] = (u_ * dc(bb)) * P[m] - dc(cc) * Q[m];
            }
            for (int m = mrec + 1; m <= np1 - 2; ++m) Q[m] = dc(0.0);
            Q[n] = (dc(__ldg(g.offdiag + n)) * u_) * dc(__ldg(g.a_diag + n));
            Q[np1] = dc(__ldg(g.a_diag + np1));
        }
        D3 sum0 = dc(0.0), sum1 = dc(0.0), sum2 = dc(0.0), sum3 = dc(0.0);
        rho_np1 = rho_np1 * rho;
        const DevHarm* trow = g.tab + tri(n, 0);
        int mtop = n < M ? n : M;
        for (int m = 0; m <= mtop; ++m) {
            D3 cv = dc(__ldg(&trow[m].cbar)), sv = dc(__ldg(&trow[m].sbar));
            D3 d_ = (cv * r_m[m] + sv * i_m[m]) * sqrt2;
            D3 e_ = dc(0.0), f_ = dc(0.0);
            if (m != 0) {
                e_ = (cv * r_m[m - 1] + sv * i_m[m - 1]) * sqrt2;
                f_ = (sv * r_m[m - 1] - cv * i_m[m - 1]) * sqrt2;
            }
            D3 mf = dc((double)m);
            sum0 = sum0 + (mf * P[m]) * e_;
            sum1 = sum1 + (mf * P[m]) * f_;
            sum2 = sum2 + (dc(__ldg(&trow[m].vr01)) * P[m + 1]) * d_;
            sum3 = sum3 + (dc(__ldg(&trow[m].vr11)) * Q[m + 1]) * d_;
        }
        D3 rr = rho_np1 / eq_radius;
        a0 = a0 + rr * sum0;
        a1 = a1 + rr * sum1;
        a2 = a2 + rr * sum2;
        a3 = a3 - rr * sum3;
        D3* tmp = P; P = Q; Q = tmp;
    }
    D3 al[3] = { a0 + a3 * s_, a1 + a3 * t_, a2 + a3 * u_ };
#pragma unroll
    for (int i = 0; i < 3; ++i) acc[i] = (R[i] * al[0].v + R[3 + i] * al[1].v) + R[6 + i] * al[2].v;
    double tmp9[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            tmp9[3 * i + j] = (R[i] * dpart(al[0], j) + R[3 + i] * dpart(al[1], j)) + R[6 + i] * dpart(al[2], j);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            G[3 * i + j] = (tmp9[3 * i] * R[j] + tmp9[3 * i + 1] * R[3 + j]) + tmp9[3 * i + 2] * R[6 + j];
}

// ------------------------------------------------------------------------- dual_eom (spacecraft.rs:312-363)
// y[9] with Cr already clamped; outputs: acc[3], G[9] = d(acc)/d(r) row-major, gcr[3] = d(acc)/d(Cr)
template <bool WITH_GRAV>
__device__ static int dual_eom_dev(const DevSetup& S, long long t_ns, const double y[9], double total_mass, double srp_area,
                                   double acc[3], double G[9], double gcr[3]) {
    // OrbitalDynamics::dual_eom, orbital.rs:116-172
    D3 rad[3] = { dvar(y[0], 0), dvar(y[1], 1), dvar(y[2], 2) };
    D3 rmag = dnorm(rad[0], rad[1], rad[2]);
    D3 fac = dc(-S.mu_central) / dcube(rmag);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        D3 ba = rad[i] * fac;
        acc[i] = ba.v;
        G[3 * i] = ba.x; G[3 * i + 1] = ba.y; G[3 * i + 2] = ba.z;
        gcr[i] = 0.0;
    }
    double bpos[NYXB_MAX_BODIES][3];
    for (int j = 0; j < S.n_bodies; ++j)
        if (!body_position(S.bodies[j], t_ns, bpos[j])) return NYXB_ERR_EPHEMERIS;
    if (S.point_mass_mask) {  // PointMasses::gradient, orbital.rs:249-307 (r_ij carries identity partials, as coded)
        double fx[3] = {0.0, 0.0, 0.0}, gp[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int j = 0; j < S.n_bodies; ++j) {
            if (!((S.point_mass_mask >> j) & 1u)) continue;
            D3 gm_d = dc(-S.bodies[j].mu);
            D3 rij[3] = { dvar(bpos[j][0], 0), dvar(bpos[j][1], 1), dvar(bpos[j][2], 2) };
            D3 rij3 = dcube(dnorm(rij[0], rij[1], rij[2]));
            D3 rj[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) { rj[i] = rad[i] - rij[i]; }
            rj[0].x = 1.0; rj[1].y = 1.0; rj[2].z = 1.0;
            D3 rj3 = dcube(dnorm(rj[0], rj[1], rj[2]));
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                D3 t = (rj[i] / rj3 + rij[i] / rij3) * gm_d;
                fx[i] += t.v;
                gp[3 * i] += t.x; gp[3 * i + 1] += t.y; gp[3 * i + 2] += t.z;
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) acc[i] += fx[i];
#pragma unroll
        for (int q = 0; q < 9; ++q) G[q] += gp[q];
    }
    if (WITH_GRAV && S.has_grav) {
        double ga[3], gg[9];
        grav_gradient(S.grav, t_ns, y, ga, gg);
#pragma unroll
        for (int i = 0; i < 3; ++i) acc[i] += ga[i];
#pragma unroll
        for (int q = 0; q < 9; ++q) G[q] += gg[q];
    }
    if (S.has_srp) {  // SolarPressure::gradient, solarpressure.rs:167-233
        const double cr = y[6];
        const double* sun = bpos[S.srp.sun_body];
        double rs[3] = { y[0] - sun[0], y[1] - sun[1], y[2] - sun[2] };
        D3 rsd[3] = { dvar(rs[0], 0), dvar(rs[1], 1), dvar(rs[2], 2) };
        D3 n_d = dnorm(rsd[0], rsd[1], rsd[2]);
        double occult = 0.0;
        double r_ls[3] = { -rs[0], -rs[1], -rs[2] };
        for (int q = 0; q < S.srp.n_shadow; ++q) {
            int bi = S.srp.shadow_body[q];
            double r_eb[3], radb;
            if (bi == NYXB_CENTRAL_BODY) { r_eb[0] = y[0]; r_eb[1] = y[1]; r_eb[2] = y[2]; radb = S.central_radius; }
            else { r_eb[0] = y[0] - bpos[bi][0]; r_eb[1] = y[1] - bpos[bi][1]; r_eb[2] = y[2] - bpos[bi][2]; radb = S.bodies[bi].radius; }
            double p = occultation(r_eb, r_ls, S.bodies[S.srp.sun_body].radius, radb);
            if (p > occult) occult = p;
        }
        double k = fabs(occult - 1.0);
        D3 r_sun_au = ddivs(n_d, NYXB_AU_KM);
        D3 inv = dc(1.0) / r_sun_au;
        D3 flux = dc(k * S.srp.phi / NYXB_C_M_S) * dsq(inv);
        D3 scal = dc(1e-3 * cr * srp_area);
        double n_sun = norm3(rs[0], rs[1], rs[2]);
        double r_au = n_sun / NYXB_AU_KM, inv_s = 1.0 / r_au;
        double flux_s = (k * S.srp.phi / NYXB_C_M_S) * (inv_s * inv_s);
        double scal_s = 1e-3 * cr * srp_area * flux_s;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            D3 f = (scal * flux) * (rsd[i] / n_d);
            acc[i] += f.v / total_mass;
            G[3 * i] += f.x / total_mass; G[3 * i + 1] += f.y / total_mass; G[3 * i + 2] += f.z / total_mass;
            if (S.srp.estimate) gcr[i] += ((scal_s * (rs[i] / n_sun)) / cr) / total_mass;
        }
    }
    return 0;
}

// ------------------------------------------------------------------------- tracking geometry
// trk_device.rs:150-152 `location`: antenna position / velocity in the integration frame and the inertial zenith
__device__ static bool station_state(const DevSetup& S, const DevStation& st, long long t_ns, double r[3], double v[3], double up[3]) {
    double R[9];
    rotation_dcm(st.rot, t_ns, R);
    const double wdot = st.rot.kind ? st.rot.wdot : 0.0;
    double vf[3] = { -(wdot * st.pos[1]), wdot * st.pos[0], 0.0 };
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        r[i] = (R[i] * st.pos[0] + R[3 + i] * st.pos[1]) + R[6 + i] * st.pos[2];
        v[i] = (R[i] * vf[0] + R[3 + i] * vf[1]) + R[6 + i] * vf[2];
        up[i] = (R[i] * st.up[0] + R[3 + i] * st.up[1]) + R[6 + i] * st.up[2];
    }
    if (st.body != NYXB_CENTRAL_BODY) {
        double bp[3], bv[3];
        if (!body_position(S.bodies[st.body], t_ns, bp) || !body_velocity(S.bodies[st.body], t_ns, bv)) return false;
#pragma unroll
        for (int i = 0; i < 3; ++i) { r[i] += bp[i]; v[i] += bv[i]; }
    }
    return true;
}

// ------------------------------------------------------------------------- one residual window of a measurement
// What od/process/mod.rs:270-352 prepares before `measurement_update`: the window's measurement types, the real observation,
// the tracker geometry (range, range rate, elevation mask, line-of-sight obstruction), `h_tilde` (msr/sensitivity.rs:88-239:
// identity rows unless the type is in msr.data; the observed range / Doppler in the denominators, as coded), the measurement
// noise and the computed observation minus the device bias.  Shared by the per-thread and the warp-cooperative filter kernels.
struct OdWindow {
    int ncur;
    int cur[2];
    bool avail[2];
    double real_obs[2], Rk[2], comp[2];
    double H[2][9];
};
enum { OD_WIN_OK = 0, OD_WIN_EMPTY = 1, OD_WIN_UNAVAILABLE = 2, OD_WIN_NOT_VISIBLE = 3, OD_WIN_EPHEMERIS = 4 };

__device__ static int od_window_setup(const DevSetup& S, const DevStation& gs, int M, int wno, const double o[2], long long t_k,
                                      const double y[9], OdWindow& w) {
    w.ncur = 0;
    for (int q = wno * M; q < (wno + 1) * M && q < gs.n_types; ++q) w.cur[w.ncur++] = gs.types[q];
    if (w.ncur == 0) return OD_WIN_EMPTY;
    bool any = false;
    w.avail[0] = w.avail[1] = false;
    for (int q = 0; q < w.ncur; ++q) { w.avail[q] = (o[w.cur[q]] == o[w.cur[q]]); any = any || w.avail[q]; }
    if (!any) return OD_WIN_UNAVAILABLE;
    w.real_obs[0] = w.real_obs[1] = 0.0;
    for (int q = 0; q < w.ncur; ++q) if (w.avail[q]) w.real_obs[q] = o[w.cur[q]];
    double r_tx[3], v_tx[3], up[3];
    if (!station_state(S, gs, t_k, r_tx, v_tx, up)) return OD_WIN_EPHEMERIS;
    const double dr[3] = { y[0] - r_tx[0], y[1] - r_tx[1], y[2] - r_tx[2] };
    const double dv[3] = { y[3] - v_tx[0], y[4] - v_tx[1], y[5] - v_tx[2] };
    const double rng = sqrt((dr[0] * dr[0] + dr[1] * dr[1]) + dr[2] * dr[2]);
    const double rr = ((dr[0] * dv[0] + dr[1] * dv[1]) + dr[2] * dv[2]) / rng;
    const double elev = asin(((dr[0] * up[0] + dr[1] * up[1]) + dr[2] * up[2]) / rng) * (180.0 / 3.14159265358979323846);
    bool visible = !(elev - gs.mask_deg < 0.0);
    if (visible && gs.body != NYXB_CENTRAL_BODY && gs.body_radius > 0.0) {   // Vallado SIGHT (anise line_of_sight_obstructed)
        double r1sq = (y[0] * y[0] + y[1] * y[1]) + y[2] * y[2];
        double r2sq = (r_tx[0] * r_tx[0] + r_tx[1] * r_tx[1]) + r_tx[2] * r_tx[2];
        double r12 = (y[0] * r_tx[0] + y[1] * r_tx[1]) + y[2] * r_tx[2];
        double tau = (r1sq - r12) / (r1sq + r2sq - 2.0 * r12);
        if (tau >= 0.0 && tau <= 1.0 && (1.0 - tau) * r1sq + r12 * tau <= gs.body_radius * gs.body_radius) visible = false;
    }
    if (!visible) return OD_WIN_NOT_VISIBLE;   // device.measure() -> None (process/mod.rs:386-392)
    for (int q = 0; q < 2; ++q)
        for (int c = 0; c < 9; ++c) w.H[q][c] = (q == c) ? 1.0 : 0.0;
    w.Rk[0] = w.Rk[1] = 0.0; w.comp[0] = w.comp[1] = 0.0;
    for (int q = 0; q < w.ncur; ++q) {
        const int slot = wno * M + q;  // position of the type in the device's list
        w.Rk[q] = gs.noise_var[slot];
        w.comp[q] = ((w.cur[q] == NYXB_MSR_RANGE) ? rng : rr) - gs.bias[slot];
        if (!w.avail[q]) continue;
        if (w.cur[q] == NYXB_MSR_DOPPLER) {
            const double rho = rng, rho_dot = o[NYXB_MSR_DOPPLER], rho2 = rho * rho;
            w.H[q][0] = dv[0] / rho - rho_dot * dr[0] / rho2;
            w.H[q][1] = dv[1] / rho - rho_dot * dr[1] / rho2;
            w.H[q][2] = dv[2] / rho - rho_dot * dr[2] / rho2;
            w.H[q][3] = dr[0] / rho; w.H[q][4] = dr[1] / rho; w.H[q][5] = dr[2] / rho;
            w.H[q][6] = 0.0; w.H[q][7] = 0.0; w.H[q][8] = 0.0;
        } else {
            const double rho = o[NYXB_MSR_RANGE];
            w.H[q][0] = dr[0] / rho; w.H[q][1] = dr[1] / rho; w.H[q][2] = dr[2] / rho;
            for (int c = 3; c < 9; ++c) w.H[q][c] = 0.0;
        }
    }
    return OD_WIN_OK;
}

// Innovation statistics (filtering.rs:152-167): S = H P H^T + R (given), Cholesky of S (fallback: of R), whitened residual ratio.
// Returns false on SingularNoiseRk.
__device__ static bool od_ratio(int M, const double Sk[2][2], const double Rk[2], const double pre[2], double& ratio) {
    double L00 = 1.0, L10 = 0.0, L11 = 1.0;
    bool chol_ok = Sk[0][0] > 0.0;
    if (chol_ok) {
        L00 = sqrt(Sk[0][0]);
        if (M == 2) {
            L10 = Sk[1][0] / L00;
            const double d = Sk[1][1] - L10 * L10;
            if (d > 0.0) L11 = sqrt(d); else chol_ok = false;
        }
    }
    double W00 = L00, W10 = L10, W11 = L11;
    if (!chol_ok) {
        if (!(Rk[0] > 0.0) || (M == 2 && !(Rk[1] > 0.0))) return false;
        W00 = sqrt(Rk[0]); W10 = 0.0; W11 = (M == 2) ? sqrt(Rk[1]) : 1.0;
    }
    const double w0 = pre[0] / W00, w1 = (M == 2) ? (pre[1] - W10 * w0) / W11 : 0.0;
    ratio = sqrt(((M == 2) ? (w0 * w0 + w1 * w1) : (w0 * w0)) / (double)M);
    return true;
}

// S^-1 for M = 1, 2 (the gain K = P H^T S^-1, filtering.rs:206-231); false on SingularKalmanGain
__device__ static bool od_sinv(int M, const double Sk[2][2], double Si[2][2]) {
    if (M == 1) { Si[0][0] = 1.0 / Sk[0][0]; Si[0][1] = Si[1][0] = 0.0; Si[1][1] = 0.0; return true; }
    const double det = Sk[0][0] * Sk[1][1] - Sk[0][1] * Sk[1][0];
    if (det == 0.0 || det != det) return false;
    Si[0][0] = Sk[1][1] / det; Si[0][1] = -Sk[0][1] / det; Si[1][0] = -Sk[1][0] / det; Si[1][1] = Sk[0][0] / det;
    return true;
}
