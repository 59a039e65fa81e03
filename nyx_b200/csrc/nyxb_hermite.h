// nyxb_hermite.h — window selection of Traj::at (md/trajectory/traj.rs:83-126) and the Hermite interpolation of
// `Interpolatable for Spacecraft` (md/trajectory/interpolatable.rs:53-108) as host/device inline functions.
// Used by the resampling kernel (nyxb_traj.cu); plain C++ when not compiled by nvcc so that a CPU test can check the very same
// arithmetic against the numpy restatement (tests/cpp/hermite_core_shim.cpp, tests/test_trajectory.py).
//
// anise's `hermite_eval` is not in the reference tree: the interpolant here is the textbook divided-difference (Newton) form
// through value/derivative pairs — the algorithm of NAIF HRMINT / SPK type 13 — evaluated with sub, div, mul, add only, in the
// operation order of nyx_b200/trajectory.py::hermite_eval (bit-identical when built without FMA contraction).
#pragma once
#include <math.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define NYXB_HD __host__ __device__ __forceinline__
#else
#define NYXB_HD inline
#endif

#define NYXB_INTERP_SAMPLES 13 /* interpolatable.rs:22 */

// Samples [first, last) used for a query whose insertion index in the ascending epoch list of `cnt` records is `idx`
// (0 < idx < cnt), traj.rs:108-117: 6 to the left, 13 in total; at the right edge the reference falls back to 12.
NYXB_HD void nyxb_hermite_window(long long idx, long long cnt, long long* first, long long* last) {
    const long long num_left = NYXB_INTERP_SAMPLES / 2;
    long long f = idx - num_left;
    if (f < 0) f = 0;
    long long l = f + NYXB_INTERP_SAMPLES;
    if (l > cnt) l = cnt;
    if (l == cnt) {
        f = l - 2 * num_left;
        if (f < 0) f = 0;
    }
    *first = f;
    *last = l;
}

// Value and derivative at x of the Hermite interpolant through nw (2..13) knots ts[] (strictly monotonic) with values ys[] and
// derivatives yd[].  q[] is scratch of 2 * nw doubles.
NYXB_HD void nyxb_hermite_eval(int nw, const double* ts, const double* ys, const double* yd, double x, double* q,
                               double* val_out, double* der_out) {
    const int m = 2 * nw;   // doubled knots z[2k] = z[2k+1] = ts[k]
    for (int k = 0; k < nw; ++k) { q[2 * k] = ys[k]; q[2 * k + 1] = ys[k]; }
    // first-order differences: the derivative at doubled knots, a plain quotient between distinct ones
    for (int i = m - 1; i >= 1; --i) {
        if (i & 1) q[i] = yd[i >> 1];
        else q[i] = (q[i] - q[i - 1]) / (ts[i >> 1] - ts[(i - 1) >> 1]);
    }
    // higher orders, in place from the bottom: after pass j, q[i] (i >= j) holds f[z_{i-j} .. z_i]
    for (int j = 2; j < m; ++j)
        for (int i = m - 1; i >= j; --i) q[i] = (q[i] - q[i - 1]) / (ts[i >> 1] - ts[(i - j) >> 1]);
    // Horner on the Newton form, value and derivative together
    double val = q[m - 1], der = 0.0;
    for (int k = m - 2; k >= 0; --k) {
        const double dx = x - ts[k >> 1];
        der = der * dx + val;
        val = val * dx + q[k];
    }
    *val_out = val;
    *der_out = der;
}

// ---- Traj::at for trajectory i of a step-major SoA recording (include/nyxb.h: nyxb_traj_sink) — the whole per-(query, trajectory)
// work of the resampling kernel, host/device so that the CPU test exercises the same indexing.
struct NyxbTrajView {
    long long cap;
    const long long* epoch;  // [cap][n]
    const double* state;     // [6][cap][n]
    const long long* count;  // [n]
};

// returns 0 (NYXB_TRAJ_OK) and rv[6], or 1 (NYXB_TRAJ_NO_DATA: TrajError::NoInterpolationData, traj.rs:84-86) with rv untouched
NYXB_HD int nyxb_traj_at(const NyxbTrajView& tv, size_t n, size_t i, long long qe, double rv[6]) {
    long long cnt = tv.count[i];
    if (cnt > tv.cap) cnt = tv.cap;
    if (cnt <= 0) return 1;
    // records are in step order: ascending epochs for a forward propagation, descending for a backward one (Traj::finalize sorts)
    const bool asc = cnt < 2 || tv.epoch[(size_t)(cnt - 1) * n + i] >= tv.epoch[i];
#define NYXB_REC(s) ((size_t)(asc ? (s) : cnt - 1 - (s)))
#define NYXB_EP(s) tv.epoch[NYXB_REC(s) * n + i]
    if (qe < NYXB_EP(0) || qe > NYXB_EP(cnt - 1)) return 1;
    long long lo = 0, hi = cnt;   // lower bound: first record with epoch >= qe
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (NYXB_EP(mid) < qe) lo = mid + 1; else hi = mid;
    }
    const long long idx = lo;
    if (NYXB_EP(idx) == qe) {     // "we actually had this exact state" traj.rs:92-95
        for (int c = 0; c < 6; ++c) rv[c] = tv.state[((size_t)c * tv.cap + NYXB_REC(idx)) * n + i];
        return 0;
    }
    long long first, last;
    nyxb_hermite_window(idx, cnt, &first, &last);
    const int nw = (int)(last - first);
    double ts[NYXB_INTERP_SAMPLES], ys[NYXB_INTERP_SAMPLES], yd[NYXB_INTERP_SAMPLES], q[2 * NYXB_INTERP_SAMPLES];
    const long long t0 = NYXB_EP(first);
    for (int k = 0; k < nw; ++k) ts[k] = (double)(NYXB_EP(first + k) - t0) * 1e-9;
    const double x = (double)(qe - t0) * 1e-9;
    for (int c = 0; c < 3; ++c) {
        for (int k = 0; k < nw; ++k) {
            const size_t r = NYXB_REC(first + k);
            ys[k] = tv.state[((size_t)c * tv.cap + r) * n + i];
            yd[k] = tv.state[((size_t)(3 + c) * tv.cap + r) * n + i];
        }
        nyxb_hermite_eval(nw, ts, ys, yd, x, q, &rv[c], &rv[3 + c]);
    }
#undef NYXB_EP
#undef NYXB_REC
    return 0;
}

// ---- Event location inside the bracketing step (propagators/event.rs:166-211): the propagation kernels stop at the end of the
// step in which the event scalar crossed zero for the trigger-th time; the root is searched on the Hermite-interpolated
// recording between the last two records with Brent's method (Brent 1973, ch. 4; the reference calls anise's `brent_solver`,
// not in the tree: restated from the published algorithm, parity unpinned).  Operation order of nyx_b200/event.py.

// event scalar minus the desired value: the closed set of include/nyxb.h (enum nyxb_event_kind), reference operation order
NYXB_HD double nyxb_event_scalar(int kind, double value, const double rv[6]) {
    double s;
    switch (kind) {
    case 1: s = sqrt((rv[0] * rv[0] + rv[1] * rv[1]) + rv[2] * rv[2]); break;        // NYXB_EVENT_RMAG
    case 2: s = (rv[0] * rv[3] + rv[1] * rv[4]) + rv[2] * rv[5]; break;              // NYXB_EVENT_RDOTV
    case 3: s = rv[0]; break;                                                        // NYXB_EVENT_X
    case 4: s = rv[1]; break;                                                        // NYXB_EVENT_Y
    case 5: s = rv[2]; break;                                                        // NYXB_EVENT_Z
    default: s = sqrt((rv[3] * rv[3] + rv[4] * rv[4]) + rv[5] * rv[5]); break;       // NYXB_EVENT_VMAG
    }
    return s - value;
}

// f(dt) of the search: the event scalar on the interpolated state dt seconds after t0 (rounded to integer ns, half to even)
NYXB_HD int nyxb_event_f(const NyxbTrajView& tv, size_t n, size_t i, int kind, double value, long long t0, double dt_s, double* f) {
    double rv[6];
    if (nyxb_traj_at(tv, n, i, t0 + llrint(dt_s * 1e9), rv)) return 1;
    *f = nyxb_event_scalar(kind, value, rv);
    return 0;
}

// returns 0 and (event epoch, interpolated state), 1 = no bracket in the recording (fewer than two records / no data),
// 2 = the last step does not bracket a root (same sign at both ends)
NYXB_HD int nyxb_event_locate_one(const NyxbTrajView& tv, size_t n, size_t i, int kind, double value, long long precision_ns,
                                  long long* ev_epoch, double rv[6]) {
    long long cnt = tv.count[i];
    if (cnt > tv.cap) cnt = tv.cap;
    if (cnt < 2) return 1;
    // the last step taken, in recording order (forward: the two largest epochs; backward: the two smallest)
    const long long e1 = tv.epoch[(size_t)(cnt - 2) * n + i], e2 = tv.epoch[(size_t)(cnt - 1) * n + i];
    const long long t0 = e1 < e2 ? e1 : e2, t1 = e1 < e2 ? e2 : e1;
    const double xtol = (double)precision_ns * 1e-9;
    double a = 0.0, b = (double)(t1 - t0) * 1e-9, fa, fb;
    if (nyxb_event_f(tv, n, i, kind, value, t0, a, &fa) || nyxb_event_f(tv, n, i, kind, value, t0, b, &fb)) return 1;
    double root = b;
    if (fa == 0.0) root = a;
    else if (fb == 0.0) root = b;
    else if (fa * fb > 0.0) return 2;
    else {
        double c = a, fc = fa, d = b - a, e = b - a;
        for (int it = 0; it < 100; ++it) {
            if (fb * fc > 0.0) { c = a; fc = fa; d = b - a; e = d; }
            if (fabs(fc) < fabs(fb)) { a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
            const double tol = 2.0 * 2.220446049250313e-16 * fabs(b) + 0.5 * xtol;
            const double m = 0.5 * (c - b);
            if (fabs(m) <= tol || fb == 0.0) break;
            if (fabs(e) >= tol && fabs(fa) > fabs(fb)) {
                const double s = fb / fa;
                double p, q;
                if (a == c) { p = 2.0 * m * s; q = 1.0 - s; }
                else {
                    const double qq = fa / fc, r = fb / fc;
                    p = s * (2.0 * m * qq * (qq - r) - (b - a) * (r - 1.0));
                    q = (qq - 1.0) * (r - 1.0) * (s - 1.0);
                }
                if (p > 0.0) q = -q;
                p = fabs(p);
                const double lim1 = 3.0 * m * q - fabs(tol * q), lim2 = fabs(e * q);
                if (2.0 * p < (lim1 < lim2 ? lim1 : lim2)) { e = d; d = p / q; }
                else { d = m; e = m; }
            } else { d = m; e = m; }
            a = b; fa = fb;
            b = (fabs(d) > tol) ? b + d : b + copysign(tol, m);
            if (nyxb_event_f(tv, n, i, kind, value, t0, b, &fb)) return 1;
        }
        root = b;
    }
    *ev_epoch = t0 + llrint(root * 1e9);
    return nyxb_traj_at(tv, n, i, *ev_epoch, rv);
}
