// nyxb_coop.cu — host side of the lane-cooperative kernel (nyxb_coop_kernel.cuh): column -> lane schedule,
// record table, dispatch on the lane count.  Kernel summary: G lanes of one warp integrate
// ONE (or two) trajectories.  The spherical-harmonic double sum (gravity_field.rs:217-249), which is >98 % of
// the arithmetic for a 21x21 field, is split across the lanes by COLUMNS of the derived-Legendre
// triangle: every A[n][m] is produced by its own column recursion (gravity_field.rs:175-181) in a
// register, and the four partial sums are regrouped so that each A[n][m] is consumed exactly once,
// by the lane that produced it:
//     X += rr_n   m A[n][m] E(n,m)              Y += rr_n m A[n][m] F(n,m)
//     Z += rr_n   vr01[n][m-1] A[n][m] D(n,m-1)  W -= rr_{n-1} vr11[n-1][m-1] A[n][m] D(n-1,m-1)
// (all three of E/F/D use the same (cos,sin)((m-1) lambda) pair, a per-column constant), so there is
// no A matrix in memory, no cross-lane traffic inside the sum, and one butterfly reduction at the end.
// RK stage vectors live in shared memory ([stage][6] per trajectory), lane c < 6 owns state component c;
// the error norm and the step-size controller are evaluated redundantly by every lane of the group
// (identical inputs -> identical decisions, no broadcast).  HBM is touched only to read the initial
// state and write the final one; coefficient records stream from L1/L2 (17.7 KB for 21x21).
//
// Reference behaviour: instance.rs:87-262, 343-352, 358-493 (propagate/single_step/derive) and
// spacecraft.rs:191-310 (eom).  FMA contraction and the regrouped summation make this a
// tolerance-parity path (tests assert < 1e-6 km, the north-star's sub-mm bound).
#include <algorithm>
#include <cmath>
#include <numeric>

#include "nyxb_coop.h"

// ------------------------------------------------------------------------------------------------
// host: column -> lane schedule (longest-processing-time greedy) and record table
// ------------------------------------------------------------------------------------------------
void nyxb_coop_build_host(int N, int M, const double* c_nm, const double* s_nm, int G, CoopHost& out) {
    const double sqrt2 = std::sqrt(2.0);
    auto C = [&](int n, int m) { return (n <= N && m <= M && m <= n) ? c_nm[(size_t)n * (N + 1) + m] : 0.0; };
    auto Sx = [&](int n, int m) { return (n <= N && m <= M && m <= n) ? s_nm[(size_t)n * (N + 1) + m] : 0.0; };
    auto vr01 = [&](int n, int m) {
        double nf = n, mf = m;
        double v = std::sqrt((nf - mf) * (nf + mf + 1.0));
        return m == 0 ? v / sqrt2 : v;
    };
    auto vr11 = [&](int n, int m) {
        double nf = n, mf = m;
        double v = std::sqrt(((2.0 * nf + 1.0) * (nf + mf + 2.0) * (nf + mf + 1.0)) / (2.0 * nf + 3.0));
        return m == 0 ? v / sqrt2 : v;
    };
    auto bnm = [&](int n, int m) {
        double nf = n, mf = m;
        return std::sqrt(((2.0 * nf + 1.0) * (2.0 * nf - 1.0)) / ((nf + mf) * (nf - mf)));
    };
    auto cnm = [&](int n, int m) {
        double nf = n, mf = m;
        return std::sqrt(((2.0 * nf + 1.0) * (nf + mf - 1.0) * (nf - mf - 1.0)) / ((nf - mf) * (nf + mf) * (2.0 * nf - 3.0)));
    };
    const int mcols = std::min(M + 1, N + 1);  // columns m = 1..mcols
    // LPT assignment
    std::vector<int> order(mcols);
    std::iota(order.begin(), order.end(), 1);  // already sorted by decreasing length (N + 2 - m)
    std::vector<int> load(G, 0);
    std::vector<std::vector<int>> cols(G);
    for (int m : order) {
        int best = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        cols[best].push_back(m);
        load[best] += N + 2 - m;
    }
    out.G = G;
    out.L = *std::max_element(load.begin(), load.end());
    out.kmax = 1;
    for (auto& c : cols) out.kmax = std::max(out.kmax, (int)c.size());
    out.recs.assign((size_t)(out.L + 1) * G, DevCoopRec{0, 0, 0, 0, 0, 0, 0, 0});  // +1: prefetch pad
    out.col_start.assign((size_t)G * out.kmax, out.L + 1);
    out.col_m.assign((size_t)G * out.kmax, 1);
    for (int lane = 0; lane < G; ++lane) {
        int e = 0;
        for (size_t k = 0; k < cols[lane].size(); ++k) {
            int m = cols[lane][k];
            out.col_start[(size_t)lane * out.kmax + k] = e;
            out.col_m[(size_t)lane * out.kmax + k] = m;
            for (int n = m; n <= N + 1; ++n, ++e) {
                DevCoopRec r{0, 0, 0, 0, 0, 0, 0, 0};
                if (n <= N) {
                    r.p1 = sqrt2 * (double)m * C(n, m);
                    r.p2 = sqrt2 * (double)m * Sx(n, m);
                    r.p3 = sqrt2 * vr01(n, m - 1) * C(n, m - 1);
                    r.p4 = sqrt2 * vr01(n, m - 1) * Sx(n, m - 1);
                }
                if (n >= 2) {
                    r.p5 = sqrt2 * vr11(n - 1, m - 1) * C(n - 1, m - 1);
                    r.p6 = sqrt2 * vr11(n - 1, m - 1) * Sx(n - 1, m - 1);
                }
                if (n <= N) {
                    if (n == m) { r.bq = std::sqrt(2.0 * (double)m + 3.0); r.cq = 0.0; }  // gravity_field.rs:168-173
                    else { r.bq = bnm(n + 1, m); r.cq = cnm(n + 1, m); }                    // gravity_field.rs:175-181
                }
                // device layout: [entry][quarter][lane] of 16-byte pieces (one coalesced 16*G-byte segment per load)
                double* base = reinterpret_cast<double*>(out.recs.data()) + (size_t)e * G * 8;
                const double q[8] = {r.p1, r.p2, r.p3, r.p4, r.p5, r.p6, r.bq, r.cq};
                for (int k = 0; k < 4; ++k) { base[(k * G + lane) * 2] = q[2 * k]; base[(k * G + lane) * 2 + 1] = q[2 * k + 1]; }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// dispatch: one translation unit per lane count (nyxb_coop_g{8,16,32}.cu) so that they build in parallel
// ------------------------------------------------------------------------------------------------
#define NYXB_COOP_DECL(G) \
    cudaError_t nyxb_launch_coop_g##G(const DevSetup*, const DevCoop*, int, size_t, const double*, const double*, const long long*, \
                                      long long, long long*, double*, long long*, nyxb_details*, int*, cudaStream_t);
NYXB_COOP_DECL(8)
NYXB_COOP_DECL(16)
NYXB_COOP_DECL(32)

extern "C" cudaError_t nyxb_launch_coop(const DevSetup* S, const DevCoop* Cp, int T, size_t n, const double* state,
                                        const double* consts, const long long* epoch0, long long end_epoch,
                                        long long* step_io, double* out_state, long long* out_epoch,
                                        nyxb_details* out_details, int* out_status, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (T != 1 && T != 2) return cudaErrorInvalidValue;
    switch (Cp->G) {
    case 8: return nyxb_launch_coop_g8(S, Cp, T, n, state, consts, epoch0, end_epoch, step_io, out_state, out_epoch, out_details, out_status, stream);
    case 16: return nyxb_launch_coop_g16(S, Cp, T, n, state, consts, epoch0, end_epoch, step_io, out_state, out_epoch, out_details, out_status, stream);
    case 32: return nyxb_launch_coop_g32(S, Cp, T, n, state, consts, epoch0, end_epoch, step_io, out_state, out_epoch, out_details, out_status, stream);
    default: return cudaErrorInvalidValue;
    }
}
