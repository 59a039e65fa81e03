// nyxb_coop.cu — lane-cooperative propagation kernel (FAST mode): G lanes of one warp integrate
// ONE trajectory.  The spherical-harmonic double sum (gravity_field.rs:217-249), which is >98 % of
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
// device
// ------------------------------------------------------------------------------------------------
#ifndef COOP_CTA
#define COOP_CTA 128  /* 5 CTAs x 128 threads per SM: 10 000 x 8 lanes fit in ONE wave of 148 SMs */
#endif
#ifndef COOP_MINB
#define COOP_MINB 5
#endif
#define COOP_SM_FIXED 120  // kst[16*6] + ys[6] + ycur[6] + nxt[6] + er[6]

// doubles of shared memory per trajectory group, padded to 8 (mod 16) doubles: the groups of one warp then
// start 64 B apart modulo the 128-B bank row instead of on the same banks
__host__ __device__ inline int coop_group_stride(int N) {
    int s = COOP_SM_FIXED + 3 * (N + 3);
    return s + ((8 - (s & 15)) & 15);
}
// bytes of the CTA-shared table region: records [(L+1)][4][G] double2, a_diag[N+3], col_start/col_m [G][kmax]
__host__ __device__ inline size_t coop_rec_bytes(int L, int G) { return (size_t)(L + 1) * G * 64; }
__host__ __device__ inline size_t coop_meta_bytes(int N, int G, int kmax) {
    size_t b = (size_t)(N + 3) * 8 + (size_t)2 * G * (kmax + 1) * 4;
    return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ double shfl_d(unsigned mask, double v, int src, int width) {
    return __shfl_sync(mask, v, src, width);
}
__device__ __forceinline__ double shfl_xor_d(unsigned mask, double v, int lanemask, int width) {
    return __shfl_xor_sync(mask, v, lanemask, width);
}

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds_s32(unsigned addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

struct GroupCtx {
    double* kst; double* ys; double* ycur; double* nxt; double* er;
    double* rm; double* im; double* rp;
    int lane;
    unsigned gmask;
    double dry_mass, extra_mass, srp_area, drag_area;
    double cr, cd, pm;  // y[6..8]: constant without guidance (spacecraft.rs:248)
    double hz;          // h * 0.0 of the current attempt: NaN-propagating stand-in for y[6..8] + h*0 (instance.rs:394)
};

// sin/cos of the three orientation angles at the step epoch + their rates: the per-stage DCM is
// obtained by a first-order update of the (slow) pole angles and an exact angle addition for W.
struct RotBase {
    double sa, ca, sd, cd, sw, cw;
    double ra_dot, dec_dot, w_dot;  // rad/s
};

// third bodies + SRP + drag for the cooperative kernel: a few hundred flops, evaluated redundantly by every
// lane, kept out of line so that the ephemeris scratch does not inflate the register count of the sum
__device__ __noinline__ int coop_extra(const DevSetup& S, const GroupCtx& g, long long t_ns, const double y[9], double acc[3]) {
    double mass = g.dry_mass + y[8] + g.extra_mass;
    const bool has_force = S.has_srp || S.has_drag;
    if (has_force && !(mass > 0.0)) return NYXB_ERR_MASSLESS;
    double bpos[NYXB_MAX_BODIES][3];
    int rc = accel_point_masses(S, t_ns, y, bpos, acc);
    if (rc) return rc;
    if (has_force) accel_post(S, t_ns, y, bpos, mass, g.srp_area, g.drag_area, acc);
    return 0;
}

// Cooperative SpacecraftDynamics::eom at the stage state held in g.ys; lane c < 6 receives dy[c].
//   recs      : record table ([entry][quarter][lane] 16-byte pieces), shared or global
//   a_diag    : [N+3];  cs/cm: column start entries / orders of this lane ([kmax+1], sentinel-terminated)
template <int G>
__device__ __forceinline__ int coop_rhs(const DevSetup& S, const double2* __restrict__ recs, int L,
                                        const double* __restrict__ a_diag, const int* __restrict__ cs,
                                        const int* __restrict__ cm, const GroupCtx& g, const RotBase& rb,
                                        double dt_s, long long t_ns, double& dyc) {
    const DevGrav& gv = S.grav;
    double inv_r, rho, ub;
    {
        // ---- inertial -> body-fixed DCM at the stage time (angle addition from the step-epoch base)
        double R[9];
        if (gv.rot.kind == 0) {
            R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        } else {
            const double da = rb.ra_dot * dt_s, dd = rb.dec_dot * dt_s, dw = rb.w_dot * dt_s;
            const double sa = fma(rb.ca, da, rb.sa), ca = fma(-rb.sa, da, rb.ca);
            const double sd = fma(rb.cd, dd, rb.sd), cd = fma(-rb.sd, dd, rb.cd);
            double sdl, cdl;
            if (fabs(dw) < 0.02) {
                const double z = dw * dw;
                sdl = dw * fma(z, fma(z, 1.0 / 120.0, -1.0 / 6.0), 1.0);
                cdl = fma(z, fma(z, fma(z, -1.0 / 720.0, 1.0 / 24.0), -0.5), 1.0);
            } else {
                det_sincos(dw, sdl, cdl);
            }
            const double sw = fma(rb.sw, cdl, rb.cw * sdl), cw = fma(rb.cw, cdl, -(rb.sw * sdl));
            const double b00 = -sa, b01 = ca;
            const double b10 = -(sd * ca), b11 = -(sd * sa), b12 = cd;
            R[0] = fma(cw, b00, sw * b10); R[1] = fma(cw, b01, sw * b11); R[2] = sw * b12;
            R[3] = fma(cw, b10, -(sw * b00)); R[4] = fma(cw, b11, -(sw * b01)); R[5] = cw * b12;
            R[6] = cd * ca; R[7] = cd * sa; R[8] = sd;
        }
        const double y0 = g.ys[0], y1 = g.ys[1], y2 = g.ys[2];
        const double rb0 = fma(R[2], y2, fma(R[1], y1, R[0] * y0));
        const double rb1 = fma(R[5], y2, fma(R[4], y1, R[3] * y0));
        const double rb2 = fma(R[8], y2, fma(R[7], y1, R[6] * y0));
        const double r_ = norm3(rb0, rb1, rb2);
        inv_r = 1.0 / r_;
        rho = gv.r_eq * inv_r;
        ub = (rb2 * inv_r) * rho;
        // park the DCM in the group's scratch (nxt/er are idle during the stages): it is only needed again after
        // the column walk, and keeping it in registers would push the walk's live set past the occupancy target
        if (g.lane == 0) {
#pragma unroll
            for (int q = 0; q < 9; ++q) g.nxt[q] = R[q];
        }
        // power-table seeds
        double zr = 1.0, zi = 0.0, pr = 1.0;
        double bzr = rb0 * inv_r, bzi = rb1 * inv_r, bp = rho;
#pragma unroll
        for (int bit = 1; bit < G; bit <<= 1) {
            if (g.lane & bit) {
                const double nzr = fma(zr, bzr, -(zi * bzi));
                zi = fma(zr, bzi, zi * bzr);
                zr = nzr;
                pr *= bp;
            }
            const double nb = fma(bzr, bzr, -(bzi * bzi));
            bzi = 2.0 * bzr * bzi;
            bzr = nb;
            bp *= bp;
        }
        const int top = gv.N + 1;
        for (int k = g.lane; k <= top; k += G) {
            g.rm[k] = zr; g.im[k] = zi; g.rp[k] = pr * a_diag[k];  // rho^k * A[k][k]: the seed of column k
            const double nzr = fma(zr, bzr, -(zi * bzi));
            zi = fma(zr, bzi, zi * bzr);
            zr = nzr;
            pr *= bp;
        }
    }
    __syncwarp(g.gmask);

    // ---- column walk; the (A, cos, sin) seed of the NEXT column is prefetched one column ahead
    double r2 = rho * rho;
    // keep the loop invariants in registers: ptxas otherwise rematerialises them (two extra DMULs per entry and
    // ~30 integer instructions of shared-memory address arithmetic per column start)
    asm volatile("" : "+d"(r2), "+d"(ub));
    unsigned a_rm = smem_u32(g.rm), a_cs = smem_u32(cs);
    asm volatile("" : "+r"(a_rm), "+r"(a_cs));
    const unsigned pw8 = (unsigned)(gv.N + 3) * 8u;   // rm -> im -> rp stride in bytes
    const unsigned cm_off = (unsigned)((cm - cs) * 4);
    double X = 0.0, Y = 0.0, Z = 0.0, W = 0.0, A = 0.0, Ap = 0.0, rr = 0.0, ii = 0.0;
    int ci = 0;
    int next_start = lds_s32(a_cs);
    int mn = lds_s32(a_cs + cm_off);
    double An0 = lds_f64(a_rm + 2 * pw8 + mn * 8);
    double rrn = lds_f64(a_rm + mn * 8 - 8), iin = lds_f64(a_rm + pw8 + mn * 8 - 8);
    const double2* rec = recs + g.lane;
    double2 n0 = rec[0], n1 = rec[G], n2 = rec[2 * G], n3 = rec[3 * G];
    for (int e = 0; e < L; ++e) {
        const double2 q0 = n0, q1 = n1, q2 = n2, q3 = n3;
        rec += G * 4;
        n0 = rec[0]; n1 = rec[G]; n2 = rec[2 * G]; n3 = rec[3 * G];  // software prefetch (table padded by one entry)
        if (e == next_start) {
            A = An0; rr = rrn; ii = iin; Ap = 0.0;
            ++ci;
            next_start = lds_s32(a_cs + ci * 4);          // sentinel L+1 after the last column
            mn = lds_s32(a_cs + cm_off + ci * 4);         // sentinel column 1
            An0 = lds_f64(a_rm + 2 * pw8 + mn * 8);
            rrn = lds_f64(a_rm + mn * 8 - 8); iin = lds_f64(a_rm + pw8 + mn * 8 - 8);
        }
        const double t1 = fma(q0.y, ii, q0.x * rr);
        const double t2 = fma(q0.y, rr, -(q0.x * ii));
        const double t3 = fma(q1.y, ii, q1.x * rr);
        const double t4 = fma(q2.y, ii, q2.x * rr);
        X = fma(A, t1, X);
        Y = fma(A, t2, Y);
        Z = fma(A, t3, Z);
        W = fma(A, t4, W);
        const double An = fma(ub * q3.x, A, -((r2 * q3.y) * Ap));
        Ap = A;
        A = An;
    }
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) {
        X += shfl_xor_d(g.gmask, X, off, G);
        Y += shfl_xor_d(g.gmask, Y, off, G);
        Z += shfl_xor_d(g.gmask, Z, off, G);
        W += shfl_xor_d(g.gmask, W, off, G);
    }
    // ---- reload the stage state and the DCM, assemble the acceleration
    double y[9];
#pragma unroll
    for (int e = 0; e < 6; ++e) y[e] = g.ys[e];
    y[6] = g.cr + g.hz; y[7] = g.cd + g.hz; y[8] = g.pm + g.hz;
    double R[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) R[q] = g.nxt[q];
    const double s_ = fma(R[2], y[2], fma(R[1], y[1], R[0] * y[0])) * inv_r;
    const double t_ = fma(R[5], y[2], fma(R[4], y[1], R[3] * y[0])) * inv_r;
    const double u_ = fma(R[8], y[2], fma(R[7], y[1], R[6] * y[0])) * inv_r;
    // rr_n A[n][m] = K0 rho (rho^n A),  rr_{n-1} A[n][m] = K0 (rho^n A),  K0 = mu / (r R_eq)
    const double K0 = gv.mu * inv_r / gv.r_eq;
    const double K1 = K0 * rho;
    const double aw = -K0 * W;
    const double ab0 = fma(aw, s_, K1 * X), ab1 = fma(aw, t_, K1 * Y), ab2 = fma(aw, u_, K1 * Z);
    // two-body (orbital.rs:86-92) from the same 1/r
    const double fac = -S.mu_central * inv_r * inv_r * inv_r;
    double acc[3];
    acc[0] = fma(fac, y[0], fma(R[6], ab2, fma(R[3], ab1, R[0] * ab0)));
    acc[1] = fma(fac, y[1], fma(R[7], ab2, fma(R[4], ab1, R[1] * ab0)));
    acc[2] = fma(fac, y[2], fma(R[8], ab2, fma(R[5], ab1, R[2] * ab0)));
    if (S.n_bodies > 0 || S.has_srp || S.has_drag) {
        int rc = coop_extra(S, g, t_ns, y, acc);
        if (rc) return rc;
    }
    double out = y[3];
    if (g.lane == 1) out = y[4];
    else if (g.lane == 2) out = y[5];
    else if (g.lane == 3) out = acc[0];
    else if (g.lane == 4) out = acc[1];
    else if (g.lane == 5) out = acc[2];
    dyc = out;
    return 0;
}

template <int G, bool SMEM_TABLE>
__global__ void __launch_bounds__(COOP_CTA, COOP_MINB)
nyxb_k_coop(const __grid_constant__ DevSetup S, const __grid_constant__ DevCoop Cp, size_t n,
            const double* __restrict__ state, const double* __restrict__ consts,
            const long long* __restrict__ epoch0, long long end_epoch, long long* __restrict__ step_io,
            double* __restrict__ out_state, long long* __restrict__ out_epoch,
            nyxb_details* __restrict__ out_details, int* __restrict__ out_status) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long tma_bar;
    const int tid = threadIdx.x;
    const int lane = tid % G, grp = tid / G;
    const int N = S.grav.N;

    // ---- CTA-shared tables: records via one TMA bulk copy (SMEM_TABLE), small metadata via plain loads
    const size_t rec_bytes = SMEM_TABLE ? coop_rec_bytes(Cp.L, G) : 0;
    unsigned char* meta = smem_raw + rec_bytes;
    double* sm_adiag = reinterpret_cast<double*>(meta);
    int* sm_cs = reinterpret_cast<int*>(meta + (size_t)(N + 3) * 8);
    int* sm_cm = sm_cs + G * (Cp.kmax + 1);
    if (SMEM_TABLE) {
        if (tid == 0) mbar_init(&tma_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&tma_bar, (unsigned)rec_bytes);
            tma_bulk_g2s(smem_raw, Cp.recs, (unsigned)rec_bytes, &tma_bar);
        }
    }
    for (int k = tid; k < N + 3; k += COOP_CTA) sm_adiag[k] = __ldg(S.grav.a_diag + k);
    for (int k = tid; k < G * (Cp.kmax + 1); k += COOP_CTA) {
        const int l = k / (Cp.kmax + 1), q = k % (Cp.kmax + 1);
        sm_cs[k] = (q < Cp.kmax) ? __ldg(Cp.col_start + l * Cp.kmax + q) : Cp.L + 1;
        sm_cm[k] = (q < Cp.kmax) ? __ldg(Cp.col_m + l * Cp.kmax + q) : 1;
    }
    if (SMEM_TABLE) mbar_wait(&tma_bar, 0);
    __syncthreads();
    const double2* recs = SMEM_TABLE ? reinterpret_cast<const double2*>(smem_raw) : reinterpret_cast<const double2*>(Cp.recs);
    const int* cs = sm_cs + lane * (Cp.kmax + 1);
    const int* cm = sm_cm + lane * (Cp.kmax + 1);

    double* sm = reinterpret_cast<double*>(meta + coop_meta_bytes(N, G, Cp.kmax)) + (size_t)grp * coop_group_stride(N);
    const int pw = N + 3;
    GroupCtx g;
    g.kst = sm; g.ys = sm + 96; g.ycur = sm + 102; g.nxt = sm + 108; g.er = sm + 114;
    g.rm = sm + COOP_SM_FIXED; g.im = g.rm + pw; g.rp = g.im + pw;
    g.lane = lane;
    g.hz = 0.0;
    const unsigned lw = tid & 31;
    g.gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lw - lane));
    const size_t traj = (size_t)blockIdx.x * (COOP_CTA / G) + grp;
    if (traj >= n) return;  // uniform per group; no block-wide barrier below this point

    // every lane of the group reads the same addresses (broadcast within the request)
    const int cidx = lane < 6 ? lane : 0;
    double yc = state[(size_t)cidx * n + traj];
    g.cr = state[6 * n + traj]; g.cd = state[7 * n + traj]; g.pm = state[8 * n + traj];
    g.dry_mass = consts[traj]; g.extra_mass = consts[n + traj]; g.srp_area = consts[2 * n + traj]; g.drag_area = consts[3 * n + traj];
    long long epoch = epoch0[traj];
    long long step_ns = step_io ? step_io[traj] : S.init_step_ns;
    int fixed = S.fixed_step;
    int status = 0, rc = 0;
    long long det_step = S.init_step_ns, n_steps = 0, n_rej = 0, n_rhs = 0;
    double det_error = 0.0;
    int det_attempts = 1;
    if (lane < 6) g.ycur[lane] = yc;
    __syncwarp(g.gmask);

    RotBase rbase;
    rbase.sa = 0; rbase.ca = 1; rbase.sd = 1; rbase.cd = 0; rbase.sw = 0; rbase.cw = 1;
    rbase.ra_dot = S.grav.rot.ra1 * NYXB_DEG2RAD / (36525.0 * 86400.0);
    rbase.dec_dot = S.grav.rot.dec1 * NYXB_DEG2RAD / (36525.0 * 86400.0);
    rbase.w_dot = S.grav.rot.wdot;

    const int stages = S.tb.stages;
    const long long duration = end_epoch - epoch;
    const long long stop = end_epoch;
    const bool backprop = duration < 0;
    bool done = (duration == 0);
    if (!done && g.pm < 0.0) { rc = NYXB_ERR_FUEL_EXHAUSTED; done = true; }
    if (!done && backprop) step_ns = -step_ns;

    while (!done) {
        // ---- instance.rs:149-196: pick this step (regular, or the final fixed step to the stop time)
        bool last = false;
        const long long prev_step = step_ns;
        const int prev_fixed = fixed;
        if ((!backprop && epoch + step_ns > stop) || (backprop && epoch + step_ns <= stop)) {
            if (stop == epoch) break;
            step_ns = stop - epoch;
            fixed = 1;
            last = true;
        }
        // ---- orientation angles at the step epoch: lanes 0..2 evaluate one sin/cos pair each
        if (S.grav.rot.kind != 0) {
            const double t_s = dur_to_seconds(epoch);
            const double d = t_s / 86400.0;
            const double T = d / 36525.0;
            double ang;
            if (lane == 0) ang = (S.grav.rot.ra0 + S.grav.rot.ra1 * T) * NYXB_DEG2RAD;
            else if (lane == 1) ang = (S.grav.rot.dec0 + S.grav.rot.dec1 * T) * NYXB_DEG2RAD;
            else ang = fmod(S.grav.rot.w0 + S.grav.rot.w1 * d, 360.0) * NYXB_DEG2RAD;
            double sv, cv;
            det_sincos(ang, sv, cv);
            rbase.sa = shfl_d(g.gmask, sv, 0, G); rbase.ca = shfl_d(g.gmask, cv, 0, G);
            rbase.sd = shfl_d(g.gmask, sv, 1, G); rbase.cd = shfl_d(g.gmask, cv, 1, G);
            rbase.sw = shfl_d(g.gmask, sv, 2, G); rbase.cw = shfl_d(g.gmask, cv, 2, G);
        }
        // ---- derive(): instance.rs:358-493
        det_attempts = 1;
        double h = dur_to_seconds(step_ns);
        long long dt_ns = 0;
        double nx = 0.0;
        for (;;) {
            double dyc;
            for (int i = 0; i < stages; ++i) {
                // stage state y + h * sum_j a_ij k_j (instance.rs:376-394); stage 0 is y itself
                if (lane < 6) {
                    double ysv = yc;
                    if (i > 0) {
                        const double* arow = &S.tb.a[(i - 1) * NYXB_MAX_STAGES];
                        double w = 0.0;
                        for (int j = 0; j < i; ++j) {
                            const double a_ij = arow[j];
                            if (a_ij != 0.0) w = fma(a_ij, g.kst[j * 6 + lane], w);
                        }
                        ysv = fma(h, w, yc);
                    }
                    g.ys[lane] = ysv;
                }
                __syncwarp(g.gmask);
                g.hz = (i > 0) ? h * 0.0 : 0.0;
                const long long off_ns = (i > 0) ? dur_from_seconds(S.tb.c[i - 1] * h) : 0;  // stage epoch is ns-truncated
                rc = coop_rhs<G>(S, recs, Cp.L, sm_adiag, cs, cm, g, rbase, (double)off_ns * 1e-9, epoch + off_ns, dyc);
                ++n_rhs;
                if (rc) break;
                if (lane < 6) g.kst[i * 6 + lane] = dyc;
            }
            if (rc) break;
            double er = 0.0;
            nx = yc;
            __syncwarp(g.gmask);  // the last stage's readers of the parked DCM (nxt) are done
            if (lane < 6) {
                for (int i = 0; i < stages; ++i) {
                    const double ki = g.kst[i * 6 + lane];
                    if (!fixed) er = fma(h * S.tb.e[i], ki, er);
                    nx = fma(h * S.tb.b[i], ki, nx);
                }
                g.nxt[lane] = nx;
                g.er[lane] = er;
            }
            __syncwarp(g.gmask);
            if (fixed) { det_step = step_ns; dt_ns = step_ns; break; }
            double e9[9], c9[9], y9[9];
#pragma unroll
            for (int e = 0; e < 6; ++e) { e9[e] = g.er[e]; c9[e] = g.nxt[e]; y9[e] = g.ycur[e]; }
            e9[6] = e9[7] = e9[8] = 0.0;
            y9[6] = g.cr; y9[7] = g.cd; y9[8] = g.pm;
            c9[6] = g.cr + g.hz; c9[7] = g.cd + g.hz; c9[8] = g.pm + g.hz;
            det_error = error_estimate(S.error_ctrl, e9, c9, y9);
            if (det_error <= S.tolerance || h <= S.min_step_s || det_attempts >= S.attempts) {
                bool bad = false;
#pragma unroll
                for (int e = 0; e < 9; ++e) bad |= (c9[e] != c9[e]);
                if (bad) { rc = NYXB_ERR_PROP_MATH; break; }
                if (det_attempts >= S.attempts) status |= NYXB_WARN_MAX_ATTEMPTS;
                det_step = dur_from_seconds(h);
                if (det_error < S.tolerance) {
                    const double proposed = 0.9 * h * pow_inv_int(S.tolerance / det_error, S.tb.order);
                    if (fabs(proposed) > fabs(S.max_step_s)) {
                        const double sg = (proposed != proposed) ? proposed : (signbit(proposed) ? -1.0 : 1.0);
                        h = S.max_step_s * sg;
                    } else {
                        h = proposed;
                    }
                }
                step_ns = dur_from_seconds(h);
                const long long ab = step_ns < 0 ? -step_ns : step_ns;
                if (ab < S.min_step_ns) step_ns = (step_ns < 0) ? -S.min_step_ns : S.min_step_ns;
                dt_ns = det_step;
                break;
            }
            det_attempts += 1;
            n_rej += 1;
            const double proposed = 0.9 * h * pow_inv_int(S.tolerance / det_error, S.tb.order - 1);
            h = (proposed < S.min_step_s) ? S.min_step_s : proposed;
            __syncwarp(g.gmask);  // everyone has read nxt/er before the retry overwrites ys
        }
        if (rc) break;
        // ---- single_step(): instance.rs:343-352
        epoch += dt_ns;
        __syncwarp(g.gmask);  // all lanes are done reading ycur/nxt
        if (lane < 6) { yc = nx; g.ycur[lane] = nx; }
        g.cr = g.cr < 0.0 ? 0.0 : (g.cr > 2.0 ? 2.0 : g.cr);  // cosmic/spacecraft.rs:494
        n_steps += 1;
        if (g.pm < 0.0) { rc = NYXB_ERR_FUEL_EXHAUSTED; break; }
        if (last) {
            step_ns = prev_step;
            fixed = prev_fixed;
            if (backprop) step_ns = -step_ns;
            break;
        }
    }
    __syncwarp(g.gmask);
    if (lane < 6) out_state[(size_t)lane * n + traj] = g.ycur[lane];
    if (lane == 6) {
        out_state[6 * n + traj] = g.cr; out_state[7 * n + traj] = g.cd; out_state[8 * n + traj] = g.pm;
        out_epoch[traj] = epoch;
        if (step_io) step_io[traj] = step_ns;
        out_status[traj] = (status & NYXB_WARN_MAX_ATTEMPTS) | rc;
    }
    if (lane == 7 && out_details) {
        nyxb_details d;
        d.step_ns = det_step; d.error = det_error; d.attempts = det_attempts; d._pad = 0;
        d.n_steps = n_steps; d.n_rejected = n_rej; d.n_rhs = n_rhs;
        out_details[traj] = d;
    }
}

template <int G>
static cudaError_t launch_g(const DevSetup* S, const DevCoop* Cp, size_t n, const double* state, const double* consts,
                            const long long* epoch0, long long end_epoch, long long* step_io, double* out_state,
                            long long* out_epoch, nyxb_details* out_details, int* out_status, cudaStream_t stream) {
    const int groups = COOP_CTA / G;
    const size_t grp_bytes = (size_t)groups * coop_group_stride(S->grav.N) * sizeof(double) + coop_meta_bytes(S->grav.N, G, Cp->kmax);
    const size_t with_table = grp_bytes + coop_rec_bytes(Cp->L, G);
    // stage the record table in shared memory when at least two CTAs still fit per SM (227 KB usable)
    const bool smem_table = with_table * 2 <= 227 * 1024;
    const size_t smem = smem_table ? with_table : grp_bytes;
    unsigned grid = (unsigned)((n + groups - 1) / groups);
    cudaError_t e;
    if (smem_table) {
        e = cudaFuncSetAttribute(nyxb_k_coop<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        nyxb_k_coop<G, true><<<grid, COOP_CTA, smem, stream>>>(*S, *Cp, n, state, consts, epoch0, end_epoch, step_io,
                                                                out_state, out_epoch, out_details, out_status);
    } else {
        e = cudaFuncSetAttribute(nyxb_k_coop<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        nyxb_k_coop<G, false><<<grid, COOP_CTA, smem, stream>>>(*S, *Cp, n, state, consts, epoch0, end_epoch, step_io,
                                                                 out_state, out_epoch, out_details, out_status);
    }
    return cudaGetLastError();
}

extern "C" cudaError_t nyxb_launch_coop(const DevSetup* S, const DevCoop* Cp, size_t n, const double* state,
                                        const double* consts, const long long* epoch0, long long end_epoch,
                                        long long* step_io, double* out_state, long long* out_epoch,
                                        nyxb_details* out_details, int* out_status, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    switch (Cp->G) {
    case 8: return launch_g<8>(S, Cp, n, state, consts, epoch0, end_epoch, step_io, out_state, out_epoch, out_details, out_status, stream);
    case 16: return launch_g<16>(S, Cp, n, state, consts, epoch0, end_epoch, step_io, out_state, out_epoch, out_details, out_status, stream);
    case 32: return launch_g<32>(S, Cp, n, state, consts, epoch0, end_epoch, step_io, out_state, out_epoch, out_details, out_status, stream);
    default: return cudaErrorInvalidValue;
    }
}
