/*
 * nyx_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the nyx-space/nyx propagation hot path, used ONLY as
 * the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs.  The product (nyx_b200/csrc) never links or calls it.
 *
 * Why a restatement: the reference is Rust and no Rust toolchain exists in this
 * image, so the reference itself cannot be compiled or run here.
 *
 * Parity pinning (tests/test_oracle_golden.py):
 *   - integrator + controller + two-body: PINNED bit-exactly against the
 *     reference's own golden vectors (tests/propagation/propagators.rs:43-52,
 *     105-144, 320-369; tests/mission_design/orbitaldyn.rs:112-119).
 *   - spherical harmonics: recursion/normalisation pinned loosely (J2 vs JPL Monte,
 *     orbitaldyn.rs:863-929, 20 m; 70x70 vs GMAT, orbitaldyn.rs:1021-1069, 200 m).
 *     The frame rotation is anise 0.10.2 + pck08.pca (absent): "parity unpinned"
 *     at that boundary — the rotation here is the documented IAU model of nyxb.h.
 *   - third bodies / SRP / eclipse: arithmetic lives in anise 0.10.2 (absent) and
 *     needs DE440s (absent): "parity unpinned"; the eclipse geometry restates
 *     anise's published `Almanac::occultation` algorithm from memory.
 *   - drag: no asserting test exists in the reference itself: "parity unpinned".
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/nyx-core/src).  Build: -O2 -ffp-contract=off (rustc never
 * contracts a*b+c into an FMA; neither may we).
 */
#include "nyx_oracle.h"
#include "nyx_oracle_priv.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* Butcher tableaux — propagators/rk_methods/{rk,dormand,verner}.rs          */
/* Flat lower-triangular a (row i has i+1 entries), then b followed by b*.   */
/* The numbers are the published GMAT / Dormand-Prince / Verner / Cash-Karp  */
/* coefficients; expressions are constant-folded in IEEE double exactly as   */
/* rustc's const evaluation does.                                            */
/* ------------------------------------------------------------------------- */
#define Q6 2.449489742783178 /* rk.rs:83 SQRT6 literal */

/* rk.rs:94-215 (RK89, order 9, 16 stages) */
static const double RK89_A[120] = {
 /* 1*/ 1.0/12.0,
 /* 2*/ 1.0/27.0, 2.0/27.0,
 /* 3*/ 1.0/24.0, 0.0, 1.0/8.0,
 /* 4*/ (4.0+94.0*Q6)/375.0, 0.0, (-94.0-84.0*Q6)/125.0, (328.0+208.0*Q6)/375.0,
 /* 5*/ (9.0-Q6)/150.0, 0.0, 0.0, (312.0+32.0*Q6)/1425.0, (69.0+29.0*Q6)/570.0,
 /* 6*/ (927.0-347.0*Q6)/1250.0, 0.0, 0.0, (-16248.0+7328.0*Q6)/9375.0, (-489.0+179.0*Q6)/3750.0, (14268.0-5798.0*Q6)/9375.0,
 /* 7*/ 2.0/27.0, 0.0, 0.0, 0.0, 0.0, (16.0-Q6)/54.0, (16.0+Q6)/54.0,
 /* 8*/ 19.0/256.0, 0.0, 0.0, 0.0, 0.0, (118.0-23.0*Q6)/512.0, (118.0+23.0*Q6)/512.0, -9.0/256.0,
 /* 9*/ 11.0/144.0, 0.0, 0.0, 0.0, 0.0, (266.0-Q6)/864.0, (266.0+Q6)/864.0, -1.0/16.0, -8.0/27.0,
 /*10*/ (5034.0-271.0*Q6)/61440.0, 0.0, 0.0, 0.0, 0.0, 0.0, (7859.0-1626.0*Q6)/10240.0, (-2232.0+813.0*Q6)/20480.0, (-594.0+271.0*Q6)/960.0, (657.0-813.0*Q6)/5120.0,
 /*11*/ (5996.0-3794.0*Q6)/405.0, 0.0, 0.0, 0.0, 0.0, (-4342.0-338.0*Q6)/9.0, (154922.0-40458.0*Q6)/135.0, (-4176.0+3794.0*Q6)/45.0, (-340864.0+242816.0*Q6)/405.0, (26304.0-15176.0*Q6)/45.0, -26624.0/81.0,
 /*12*/ (3793.0+2168.0*Q6)/103680.0, 0.0, 0.0, 0.0, 0.0, (4042.0+2263.0*Q6)/13824.0, (-231278.0+40717.0*Q6)/69120.0, (7947.0-2168.0*Q6)/11520.0, (1048.0-542.0*Q6)/405.0, (-1383.0+542.0*Q6)/720.0, 2624.0/1053.0, 3.0/1664.0,
 /*13*/ -137.0/1296.0, 0.0, 0.0, 0.0, 0.0, (5642.0-337.0*Q6)/864.0, (5642.0+337.0*Q6)/864.0, -299.0/48.0, 184.0/81.0, -44.0/9.0, -5120.0/1053.0, -11.0/468.0, 16.0/9.0,
 /*14*/ (33617.0-2168.0*Q6)/518400.0, 0.0, 0.0, 0.0, 0.0, (-3846.0+31.0*Q6)/13824.0, (155338.0-52807.0*Q6)/345600.0, (-12537.0+2168.0*Q6)/57600.0, (92.0+542.0*Q6)/2025.0, (-1797.0-542.0*Q6)/3600.0, 320.0/567.0, -1.0/1920.0, 4.0/105.0, 0.0,
 /*15*/ (-36487.0-30352.0*Q6)/279600.0, 0.0, 0.0, 0.0, 0.0, (-29666.0-4499.0*Q6)/7456.0, (2779182.0-615973.0*Q6)/186400.0, (-94329.0+91056.0*Q6)/93200.0, (-232192.0+121408.0*Q6)/17475.0, (101226.0-22764.0*Q6)/5825.0, -169984.0/9087.0, -87.0/30290.0, 492.0/1165.0, 0.0, 1260.0/233.0,
};
/* rk.rs:216-251: b then b* (b* = b + GMAT error weights) */
static const double RK89_B[32] = {
 23.0/525.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 171.0/1400.0, 86.0/525.0, 93.0/280.0, -2048.0/6825.0, -3.0/18200.0, 39.0/175.0, 0.0, 9.0/25.0, 233.0/4200.0,
 23.0/525.0+7.0/400.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 171.0/1400.0-63.0/200.0, 86.0/525.0+14.0/25.0, 93.0/280.0-21.0/20.0, -2048.0/6825.0+1024.0/975.0, -3.0/18200.0+21.0/36400.0, 39.0/175.0+3.0/25.0, 9.0/280.0, 0.0, 0.0,
};

/* dormand.rs:74-152 (Dormand-Prince 7(8), order 8, 13 stages) */
static const double DP78_A[78] = {
 /* 1*/ 1.0/18.0,
 /* 2*/ 1.0/48.0, 1.0/16.0,
 /* 3*/ 1.0/32.0, 0.0, 3.0/32.0,
 /* 4*/ 5.0/16.0, 0.0, -75.0/64.0, 75.0/64.0,
 /* 5*/ 3.0/80.0, 0.0, 0.0, 3.0/16.0, 3.0/20.0,
 /* 6*/ 29443841.0/614563906.0, 0.0, 0.0, 77736538.0/692538347.0, -28693883.0/1125000000.0, 23124283.0/1800000000.0,
 /* 7*/ 16016141.0/946692911.0, 0.0, 0.0, 61564180.0/158732637.0, 22789713.0/633445777.0, 545815736.0/2771057229.0, -180193667.0/1043307555.0,
 /* 8*/ 39632708.0/573591083.0, 0.0, 0.0, -433636366.0/683701615.0, -421739975.0/2616292301.0, 100302831.0/723423059.0, 790204164.0/839813087.0, 800635310.0/3783071287.0,
 /* 9*/ 246121993.0/1340847787.0, 0.0, 0.0, -37695042795.0/15268766246.0, -309121744.0/1061227803.0, -12992083.0/490766935.0, 6005943493.0/2108947869.0, 393006217.0/1396673457.0, 123872331.0/1001029789.0,
 /*10*/ -1028468189.0/846180014.0, 0.0, 0.0, 8478235783.0/508512852.0, 1311729495.0/1432422823.0, -10304129995.0/1701304382.0, -48777925059.0/3047939560.0, 15336726248.0/1032824649.0, -45442868181.0/3398467696.0, 3065993473.0/597172653.0,
 /*11*/ 185892177.0/718116043.0, 0.0, 0.0, -3185094517.0/667107341.0, -477755414.0/1098053517.0, -703635378.0/230739211.0, 5731566787.0/1027545527.0, 5232866602.0/850066563.0, -4093664535.0/808688257.0, 3962137247.0/1805957418.0, 65686358.0/487910083.0,
 /*12*/ 403863854.0/491063109.0, 0.0, 0.0, -5068492393.0/434740067.0, -411421997.0/543043805.0, 652783627.0/914296604.0, 11173962825.0/925320556.0, -13158990841.0/6184727034.0, 3936647629.0/1978049680.0, -160528059.0/685178525.0, 248638103.0/1413531060.0, 0.0,
};
/* dormand.rs:153-182 */
static const double DP78_B[26] = {
 14005451.0/335480064.0, 0.0, 0.0, 0.0, 0.0, -59238493.0/1068277825.0, 181606767.0/758867731.0, 561292985.0/797845732.0, -1041891430.0/1371343529.0, 760417239.0/1151165299.0, 118820643.0/751138087.0, -528747749.0/2220607170.0, 0.25,
 13451932.0/455176623.0, 0.0, 0.0, 0.0, 0.0, -808719846.0/976000145.0, 1757004468.0/5645159321.0, 656045339.0/265891186.0, -3867574721.0/1518517206.0, 465885868.0/322736535.0, 53011238.0/667516719.0, 2.0/45.0, 0.0,
};

/* dormand.rs:26-48 (Dormand-Prince 4(5), order 5, 7 stages) */
static const double DP45_A[21] = {
 1.0/5.0,
 3.0/40.0, 9.0/40.0,
 44.0/45.0, -56.0/15.0, 32.0/9.0,
 19372.0/6561.0, -25360.0/2187.0, 64448.0/6561.0, -212.0/729.0,
 9017.0/3168.0, -355.0/33.0, 46732.0/5247.0, 49.0/176.0, -5103.0/18656.0,
 35.0/384.0, 0.0, 500.0/1113.0, 125.0/192.0, -2187.0/6784.0, 11.0/84.0,
};
/* dormand.rs:49-64 */
static const double DP45_B[14] = {
 35.0/384.0, 0.0, 500.0/1113.0, 125.0/192.0, -2187.0/6784.0, 11.0/84.0, 0.0,
 5179.0/57600.0, 0.0, 7571.0/16695.0, 393.0/640.0, -92097.0/339200.0, 187.0/2100.0, 1.0/40.0,
};

/* rk.rs:69 (classic RK4; b duplicated so the error estimate is zero, rk.rs:70-80) */
static const double RK4_A[6] = { 0.5, 0.0, 0.5, 0.0, 0.0, 1.0 };
static const double RK4_B[8] = { 1.0/6.0, 1.0/3.0, 1.0/3.0, 1.0/6.0, 1.0/6.0, 1.0/3.0, 1.0/3.0, 1.0/6.0 };

/* rk.rs:26-42 (Cash-Karp 4(5), order 5, 6 stages) */
static const double CK45_A[15] = {
 1.0/5.0,
 3.0/40.0, 9.0/40.0,
 3.0/10.0, -9.0/10.0, 6.0/5.0,
 -11.0/54.0, 5.0/2.0, -70.0/27.0, 35.0/27.0,
 1631.0/55296.0, 175.0/512.0, 575.0/13824.0, 44275.0/110592.0, 253.0/4096.0,
};
/* rk.rs:43-56 */
static const double CK45_B[12] = {
 37.0/378.0, 0.0, 250.0/621.0, 125.0/594.0, 0.0, 512.0/1771.0,
 2825.0/27648.0, 0.0, 18575.0/48384.0, 13525.0/55296.0, 277.0/14336.0, 1.0/4.0,
};

/* verner.rs:31-59 (Verner 5(6), order 6, 8 stages) */
static const double V56_A[28] = {
 1.0/6.0,
 4.0/75.0, 16.0/75.0,
 5.0/6.0, -8.0/3.0, 5.0/2.0,
 -165.0/64.0, 55.0/6.0, -425.0/64.0, 85.0/96.0,
 -8263.0/15000.0, 124.0/75.0, -643.0/680.0, -81.0/250.0, 2484.0/10625.0,
 3501.0/1720.0, -300.0/43.0, 297275.0/52632.0, -319.0/2322.0, 24068.0/84065.0, 3850.0/26703.0,
 12.0/5.0, -8.0, 4015.0/612.0, -11.0/36.0, 88.0/255.0, 0.0, 0.0,
};
/* verner.rs:61-78 */
static const double V56_B[16] = {
 3.0/40.0, 0.0, 875.0/2244.0, 23.0/72.0, 264.0/1955.0, 125.0/11592.0, 43.0/616.0, 0.0,
 13.0/160.0, 0.0, 2375.0/5984.0, 5.0/16.0, 12.0/85.0, 0.0, 0.0, 3.0/44.0,
};

typedef struct { int order, stages; const double *a, *b; } tableau_t;

/* rk_methods/mod.rs:81-133 */
static int tableau_for(int method, tableau_t* t) {
    switch (method) {
    case NYXB_RK89: *t = (tableau_t){9, 16, RK89_A, RK89_B}; return 0;
    case NYXB_DP78: *t = (tableau_t){8, 13, DP78_A, DP78_B}; return 0;
    case NYXB_DP45: *t = (tableau_t){5, 7, DP45_A, DP45_B}; return 0;
    case NYXB_RK4:  *t = (tableau_t){4, 4, RK4_A, RK4_B}; return 0;
    case NYXB_CK45: *t = (tableau_t){5, 6, CK45_A, CK45_B}; return 0;
    case NYXB_V56:  *t = (tableau_t){6, 8, V56_A, V56_B}; return 0;
    default: return -1;
    }
}

int nyx_oracle_tableau(int method, int* order, int* stages, const double** a, const double** b) {
    tableau_t t;
    if (tableau_for(method, &t)) return -1;
    *order = t.order; *stages = t.stages; *a = t.a; *b = t.b;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* hifitime 4.3.0 (not in the reference tree) Duration semantics, pinned by   */
/* the bit-exact golden vectors:                              */
/*   f64 * Unit::Second -> Duration : trunc(s * 1e9) ns  (`as i128` cast)     */
/*   Duration::to_seconds()         : whole seconds + subsec_ns * 1e-9,       */
/*                                    plus centuries * 3155760000.0 if != 0   */
/* ------------------------------------------------------------------------- */
#define NS_PER_S 1000000000LL
#define NS_PER_CENTURY 3155760000000000000LL

double nyx_oracle_dur_to_seconds(int64_t total_ns) {
    int64_t cent = total_ns / NS_PER_CENTURY;
    if (total_ns % NS_PER_CENTURY < 0) cent -= 1; /* floor division: centuries may be -1 */
    int64_t nanos = total_ns - cent * NS_PER_CENTURY; /* in [0, NS_PER_CENTURY) */
    int64_t sec = nanos / NS_PER_S;
    int64_t sub = nanos % NS_PER_S;
    if (cent == 0) return (double)sec + (double)sub * 1e-9;
    return (double)cent * 3155760000.0 + (double)sec + (double)sub * 1e-9;
}

int64_t nyx_oracle_dur_from_seconds(double s) {
    double ns = s * 1e9;
    if (ns != ns) return 0;                 /* Rust `NaN as i128` == 0 */
    if (ns >= 9.2e18) return INT64_MAX;     /* saturating cast */
    if (ns <= -9.2e18) return INT64_MIN;
    return (int64_t)ns;                     /* truncation toward zero */
}

/* ------------------------------------------------------------------------- */
/* Deterministic sin/cos (no libm): identical operation sequence on CPU and   */
/* GPU so that frame rotations are bit-reproducible.  Cody-Waite reduction by */
/* pi/2 (three 33-bit parts, fdlibm constants) + fdlibm kernel polynomials.   */
/* Valid for |x| < ~1e5; accuracy < 2 ulp (tests/test_oracle_units.py).       */
/* ------------------------------------------------------------------------- */
void nyx_oracle_sincos(double x, double* s, double* c) {
    const double two_over_pi = 6.36619772367581382433e-01;
    const double p1 = 1.57079632673412561417e+00;
    const double p2 = 6.07710050630396597660e-11;
    const double p3 = 2.02226624879595063154e-21;
    double kf = rint(x * two_over_pi);
    double r = ((x - kf * p1) - kf * p2) - kf * p3;
    double z = r * r;
    double ps = -1.66666666666666324348e-01 + z * (8.33333333332248946124e-03 + z * (-1.98412698298579493134e-04 + z * (2.75573137070700676789e-06 + z * (-2.50507602534068634195e-08 + z * 1.58969099521155010221e-10))));
    double sn = r + (r * z) * ps;
    double pc = 4.16666666666666019037e-02 + z * (-1.38888888888741095749e-03 + z * (2.48015872894767294178e-05 + z * (-2.75573143513906633035e-07 + z * (2.08757232129817482790e-09 + z * -1.13596475577881948265e-11))));
    double cs = (1.0 - 0.5 * z) + (z * z) * pc;
    long long k = (long long)kf;
    switch (k & 3) {
    case 0: *s = sn; *c = cs; break;
    case 1: *s = cs; *c = -sn; break;
    case 2: *s = -sn; *c = -cs; break;
    default: *s = -cs; *c = sn; break;
    }
}

#define DEG2RAD 1.7453292519943295e-2

/* Orientation model of nyxb.h (stands in for anise `Almanac::rotate`,
 * gravity_field.rs:258-265).  R maps inertial -> body-fixed; wdot in rad/s. */
void nyx_oracle_rotation(const nyxb_rotation* rot, int64_t t_ns, double R[9], double* wdot) {
    if (rot->kind == 0) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        *wdot = 0.0;
        return;
    }
    double t_s = nyx_oracle_dur_to_seconds(t_ns);
    double d = t_s / 86400.0;
    double T = d / 36525.0;
    double ra = (rot->ra0_deg + rot->ra1_deg_cy * T) * DEG2RAD;
    double dec = (rot->dec0_deg + rot->dec1_deg_cy * T) * DEG2RAD;
    double w = fmod(rot->w0_deg + rot->w1_deg_day * d, 360.0) * DEG2RAD;
    double sa, ca, sd, cd, sw, cw;
    nyx_oracle_sincos(ra, &sa, &ca);
    nyx_oracle_sincos(dec, &sd, &cd);
    nyx_oracle_sincos(w, &sw, &cw);
    /* BA = R1(90deg - dec) * R3(90deg + ra) */
    double b00 = -sa, b01 = ca, b02 = 0.0;
    double b10 = -(sd * ca), b11 = -(sd * sa), b12 = cd;
    double b20 = cd * ca, b21 = cd * sa, b22 = sd;
    R[0] = cw * b00 + sw * b10; R[1] = cw * b01 + sw * b11; R[2] = cw * b02 + sw * b12;
    R[3] = cw * b10 - sw * b00; R[4] = cw * b11 - sw * b01; R[5] = cw * b12 - sw * b02;
    R[6] = b20; R[7] = b21; R[8] = b22;
    *wdot = rot->w1_deg_day * DEG2RAD / 86400.0;
}

/* ------------------------------------------------------------------------- */
/* Ephemeris: piecewise Chebyshev position (stands in for anise SPK           */
/* evaluation behind `almanac.transform`, orbital.rs:230-234).                */
/* ------------------------------------------------------------------------- */
int nyx_oracle_body_position(const nyxb_body* b, int64_t t_ns, double pos[3]) {
    int64_t dt = t_ns - b->t0_ns;
    if (dt < 0) return -1;
    int64_t idx = dt / b->interval_ns;
    if (idx >= b->n_intervals) return -1;
    int64_t off = dt - idx * b->interval_ns;
    double tau = 2.0 * ((double)off / (double)b->interval_ns) - 1.0;
    double tau2 = 2.0 * tau;
    int nc = b->n_coeffs;
    const double* c = b->coeffs + (size_t)idx * 3 * (size_t)nc;
    for (int ax = 0; ax < 3; ++ax) {
        const double* ca = c + ax * nc;
        double b1 = 0.0, b2 = 0.0;
        for (int k = nc - 1; k >= 1; --k) {
            double bk = (tau2 * b1 - b2) + ca[k];
            b2 = b1; b1 = bk;
        }
        pos[ax] = (tau * b1 - b2) + ca[0];
    }
    return 0;
}

/* nalgebra Vector3::norm(): sqrt((a*a + b*b) + c*c) — pinned by the golden vectors */
/* d/dt of the Chebyshev series (anise differentiates the SPK segment the same way): sum_k c_k T_k'(tau) * 2 / interval */
int nyx_oracle_body_velocity(const nyxb_body* b, int64_t t_ns, double vel[3]) {
    int64_t dt = t_ns - b->t0_ns;
    if (dt < 0) return 1;
    int64_t idx = dt / b->interval_ns;
    if (idx >= b->n_intervals) return 1;
    int64_t off = dt - idx * b->interval_ns;
    double tau = 2.0 * ((double)off / (double)b->interval_ns) - 1.0;
    double tau2 = 2.0 * tau;
    int nc = b->n_coeffs;
    const double* c = b->coeffs + (size_t)idx * 3 * (size_t)nc;
    double scale = 2.0 / ((double)b->interval_ns * 1e-9);
    for (int ax = 0; ax < 3; ++ax) {
        const double* ca = c + ax * nc;
        double b1 = 0.0, b2 = 0.0;
        for (int j = nc - 2; j >= 0; --j) {
            double bj = ((double)(j + 1) * ca[j + 1] + tau2 * b1) - b2;
            b2 = b1; b1 = bj;
        }
        vel[ax] = b1 * scale;
    }
    return 0;
}

static inline double norm3(const double v[3]) {
    return sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
}

/* ------------------------------------------------------------------------- */
/* GravityField — dynamics/gravity_field.rs                                   */
/* ------------------------------------------------------------------------- */
/* struct nyx_oracle_grav: see nyx_oracle_priv.h (shared with nyx_oracle_od.c) */

/* gravity_field.rs:52-132 (GravityField::new) */
nyx_oracle_grav* nyx_oracle_grav_new(const nyxb_gravity_field* g) {
    nyx_oracle_grav* h = (nyx_oracle_grav*)calloc(1, sizeof(*h));
    int N = g->degree, np2 = N + 2;
    h->n = N; h->m = g->order; h->dim = N + 3; h->mu = g->mu_km3_s2; h->r_eq = g->r_eq_km; h->rot = g->rot;
    h->a_diag = (double*)calloc((size_t)np2 + 1, sizeof(double));
    h->b_nm = (double*)calloc((size_t)np2 * np2, sizeof(double));
    h->c_nm = (double*)calloc((size_t)np2 * np2, sizeof(double));
    h->vr01 = (double*)calloc((size_t)np2 * np2, sizeof(double));
    h->vr11 = (double*)calloc((size_t)np2 * np2, sizeof(double));
    h->cbar = (double*)malloc(sizeof(double) * (size_t)(N + 1) * (N + 1));
    h->sbar = (double*)malloc(sizeof(double) * (size_t)(N + 1) * (N + 1));
    memcpy(h->cbar, g->c_nm, sizeof(double) * (size_t)(N + 1) * (N + 1));
    memcpy(h->sbar, g->s_nm, sizeof(double) * (size_t)(N + 1) * (N + 1));
    h->a_diag[0] = 1.0;
    for (int n = 1; n <= np2; ++n) {
        double nf = (double)n;
        h->a_diag[n] = sqrt(1.0 + 1.0 / (2.0 * nf)) * h->a_diag[n - 1];
    }
    for (int n = 0; n < np2; ++n) {
        for (int m = 0; m < np2; ++m) {
            double nf = (double)n, mf = (double)m;
            h->c_nm[n * np2 + m] = sqrt(((2.0 * nf + 1.0) * (nf + mf - 1.0) * (nf - mf - 1.0)) / ((nf - mf) * (nf + mf) * (2.0 * nf - 3.0)));
            h->b_nm[n * np2 + m] = sqrt(((2.0 * nf + 1.0) * (2.0 * nf - 1.0)) / ((nf + mf) * (nf - mf)));
            h->vr01[n * np2 + m] = sqrt((nf - mf) * (nf + mf + 1.0));
            h->vr11[n * np2 + m] = sqrt(((2.0 * nf + 1.0) * (nf + mf + 2.0) * (nf + mf + 1.0)) / (2.0 * nf + 3.0));
            if (m == 0) {
                h->vr01[n * np2 + m] /= sqrt(2.0);
                h->vr11[n * np2 + m] /= sqrt(2.0);
            }
        }
    }
    return h;
}

void nyx_oracle_grav_free(nyx_oracle_grav* h) {
    if (!h) return;
    free(h->a_diag); free(h->b_nm); free(h->c_nm); free(h->vr01); free(h->vr11); free(h->cbar); free(h->sbar);
    free(h);
}

/* gravity_field.rs:148-268 (GravityField::eom).  `scratch` holds (N+3)^2 doubles:
 * the per-call clone of a_nm (gravity_field.rs:165) without the heap allocation. */
void nyx_oracle_grav_accel(const nyx_oracle_grav* h, int64_t t_ns, const double r_in[3], double* scratch, double acc[3]) {
    int N = h->n, M = h->m, dim = h->dim, np2 = N + 2;
    double R[9], wdot;
    nyx_oracle_rotation(&h->rot, t_ns, R, &wdot);
    /* :150-154 transform_to(body-fixed): r_bf = R r */
    double rb[3];
    for (int i = 0; i < 3; ++i) rb[i] = (R[3 * i] * r_in[0] + R[3 * i + 1] * r_in[1]) + R[3 * i + 2] * r_in[2];
    /* :157-160 */
    double r_ = norm3(rb);
    double s_ = rb[0] / r_, t_ = rb[1] / r_, u_ = rb[2] / r_;
    /* :165 clone of the precomputed matrix: zeros + diagonal */
    double* a = scratch;
    memset(a, 0, sizeof(double) * (size_t)dim * dim);
    for (int k = 0; k <= np2; ++k) a[k * dim + k] = h->a_diag[k];
    /* :168-173 */
    a[1 * dim + 0] = u_ * sqrt(3.0);
    for (int n = 1; n <= N + 1; ++n) {
        double nf = (double)n;
        a[(n + 1) * dim + n] = sqrt(2.0 * nf + 3.0) * u_ * a[n * dim + n];
    }
    /* :175-181 */
    for (int m = 0; m <= M + 1; ++m)
        for (int n = m + 2; n <= N + 1; ++n)
            a[n * dim + m] = u_ * h->b_nm[n * np2 + m] * a[(n - 1) * dim + m] - h->c_nm[n * np2 + m] * a[(n - 2) * dim + m];
    /* :184-193 */
    int mm = N < M ? N : M;
    double r_m[mm + 2], i_m[mm + 2];
    r_m[0] = 1.0; i_m[0] = 0.0;
    for (int m = 1; m <= mm; ++m) {
        r_m[m] = s_ * r_m[m - 1] - t_ * i_m[m - 1];
        i_m[m] = s_ * i_m[m - 1] + t_ * r_m[m - 1];
    }
    /* :209-215 */
    double rho = h->r_eq / r_;
    double rho_np1 = h->mu / r_ * rho;
    double a4[4] = {0, 0, 0, 0};
    const double sqrt2 = sqrt(2.0);
    /* :217-249 */
    for (int n = 1; n <= N; ++n) {
        double sum[4] = {0, 0, 0, 0};
        rho_np1 *= rho;
        int mtop = n < M ? n : M;
        for (int m = 0; m <= mtop; ++m) {
            double cv = h->cbar[n * (N + 1) + m], sv = h->sbar[n * (N + 1) + m];
            double d_ = (cv * r_m[m] + sv * i_m[m]) * sqrt2;
            double e_ = (m == 0) ? 0.0 : (cv * r_m[m - 1] + sv * i_m[m - 1]) * sqrt2;
            double f_ = (m == 0) ? 0.0 : (sv * r_m[m - 1] - cv * i_m[m - 1]) * sqrt2;
            sum[0] += (double)m * a[n * dim + m] * e_;
            sum[1] += (double)m * a[n * dim + m] * f_;
            sum[2] += h->vr01[n * np2 + m] * a[n * dim + m + 1] * d_;
            sum[3] -= h->vr11[n * np2 + m] * a[(n + 1) * dim + m + 1] * d_;
        }
        double rr = rho_np1 / h->r_eq;
        for (int q = 0; q < 4; ++q) a4[q] += rr * sum[q];
    }
    /* :250-254 */
    double ab[3] = { a4[0] + a4[3] * s_, a4[1] + a4[3] * t_, a4[2] + a4[3] * u_ };
    /* :258-267 rotate back: R^T a */
    for (int i = 0; i < 3; ++i) acc[i] = (R[i] * ab[0] + R[3 + i] * ab[1]) + R[6 + i] * ab[2];
}

/* ------------------------------------------------------------------------- */
/* Eclipse — cosmic/eclipse.rs:69-83 calling anise 0.10.2                     */
/* `Almanac::solar_eclipsing` -> `occultation` (absent; restated from the     */
/* published algorithm: apparent-disk overlap).  Returns percentage/100.      */
/* ------------------------------------------------------------------------- */
static double circ_seg_area(double r, double d) {
    return (r * r) * acos(d / r) - d * sqrt(r * r - d * d);
}

double nyx_oracle_occultation(const double r_eb[3] /* eclipsing body -> observer */,
                              const double r_ls[3] /* observer -> light source */,
                              double light_radius_km, double body_radius_km) {
    double n_ls = norm3(r_ls), n_eb = norm3(r_eb);
    double r_ls_prime = (light_radius_km >= n_ls) ? light_radius_km : asin(light_radius_km / n_ls);
    double r_fobj_prime = (body_radius_km >= n_eb) ? body_radius_km : asin(body_radius_km / n_eb);
    double dot = (r_ls[0] * r_eb[0] + r_ls[1] * r_eb[1]) + r_ls[2] * r_eb[2];
    double d_prime = acos(-dot / (n_eb * n_ls));
    if (d_prime - r_ls_prime > r_fobj_prime) return 0.0;
    if (r_fobj_prime > d_prime + r_ls_prime) return 1.0;
    if (fabs(r_ls_prime - r_fobj_prime) < d_prime && d_prime < r_ls_prime + r_fobj_prime) {
        double d1 = (d_prime * d_prime - r_ls_prime * r_ls_prime + r_fobj_prime * r_fobj_prime) / (2.0 * d_prime);
        double d2 = (d_prime * d_prime + r_ls_prime * r_ls_prime - r_fobj_prime * r_fobj_prime) / (2.0 * d_prime);
        double shadow_area = circ_seg_area(r_fobj_prime, d1) + circ_seg_area(r_ls_prime, d2);
        if (shadow_area != shadow_area) return 1.0;
        double nominal_area = M_PI * (r_ls_prime * r_ls_prime);
        return shadow_area / nominal_area;
    }
    /* annular */
    return (r_fobj_prime * r_fobj_prime) / (r_ls_prime * r_ls_prime);
}

/* ------------------------------------------------------------------------- */
/* SpacecraftDynamics::eom — dynamics/spacecraft.rs:191-310 (stm = None,      */
/* guid_law = None branch) and everything it calls.                           */
/* ------------------------------------------------------------------------- */
typedef struct {
    const nyxb_dynamics* dyn;
    int n_grav;                          /* harmonic fields in accel-model order (orbital.rs:44-46, 102-107) */
    nyx_oracle_grav* grav[NYXB_MAX_FIELDS];
    int grav_body[NYXB_MAX_FIELDS];      /* NYXB_CENTRAL_BODY or the body the field belongs to */
    double* grav_scratch;                /* sized for the largest field */
    double dry_mass, extra_mass, srp_area, drag_area;
    int64_t n_rhs;
} eom_ctx;

/* fields of a dynamics descriptor -> ctx; returns the scratch size (doubles) */
static size_t ctx_fields_new(eom_ctx* cx, const nyxb_dynamics* dyn) {
    size_t need = 0;
    cx->n_grav = dyn->gravity ? (dyn->n_gravity > 0 ? dyn->n_gravity : 1) : 0;
    if (cx->n_grav > NYXB_MAX_FIELDS) cx->n_grav = NYXB_MAX_FIELDS;
    for (int f = 0; f < cx->n_grav; ++f) {
        cx->grav[f] = nyx_oracle_grav_new(&dyn->gravity[f]);
        cx->grav_body[f] = dyn->gravity[f].body;
        size_t d = (size_t)cx->grav[f]->dim * cx->grav[f]->dim;
        if (d > need) need = d;
    }
    return need;
}
static void ctx_fields_free(eom_ctx* cx) {
    for (int f = 0; f < cx->n_grav; ++f) nyx_oracle_grav_free(cx->grav[f]);
}

#define AU_KM 149597870.700                 /* cosmic/mod.rs:183 */
#define SPEED_OF_LIGHT_M_S (299792.458 * 1e3) /* cosmic/mod.rs:179-180 */

/* returns 0 or an nyxb_status error */
static int eom(eom_ctx* cx, int64_t epoch_ns, double delta_t_s, const double y[9], double dy[9]) {
    const nyxb_dynamics* dyn = cx->dyn;
    cx->n_rhs++;
    /* spacecraft.rs:199 + cosmic/mod.rs:94-104: stage epoch = ctx.epoch + delta_t_s (ns-truncated) */
    int64_t t_ns = epoch_ns + nyx_oracle_dur_from_seconds(delta_t_s);
    /* cosmic/spacecraft.rs:494: Cr clamped on every `set` */
    double cr = y[6] < 0.0 ? 0.0 : (y[6] > 2.0 ? 2.0 : y[6]);
    double cd = y[7];
    double mass = cx->dry_mass + y[8] + cx->extra_mass; /* Spacecraft::mass_kg cosmic/spacecraft.rs:301 */
    int has_force = (dyn->srp != NULL) || (dyn->drag != NULL);
    /* spacecraft.rs:201-203 */
    if (has_force && !(mass > 0.0)) return NYXB_ERR_MASSLESS;

    const double* r = y;
    const double* v = y + 3;
    /* orbital.rs:86-92: (-mu / r^3) * r_vec, r^3 = powi(3) = r*r*r */
    double rmag = norm3(r);
    double fac = -dyn->mu_central_km3_s2 / (rmag * rmag * rmag);
    double acc[3] = { fac * r[0], fac * r[1], fac * r[2] };

    /* body positions at the stage epoch (anise `transform`, orbital.rs:230-234) */
    double bpos[NYXB_MAX_BODIES][3];
    for (int j = 0; j < dyn->n_bodies; ++j)
        if (nyx_oracle_body_position(&dyn->bodies[j], t_ns, bpos[j])) return NYXB_ERR_EPHEMERIS;

    /* orbital.rs:102-107: accel models in order [PointMasses, GravityField] */
    /* PointMasses::eom orbital.rs:213-247 */
    if (dyn->point_mass_mask) {
        double dx[3] = {0, 0, 0};
        /* orbital.rs:217: `for third_body in &self.celestial_objects` — the caller's order when it is given */
        int order[NYXB_MAX_BODIES], n_pm = 0;
        if (dyn->n_point_masses > 0) { for (int q = 0; q < dyn->n_point_masses; ++q) order[n_pm++] = dyn->point_mass_order[q]; }
        else { for (int j = 0; j < dyn->n_bodies; ++j) if ((dyn->point_mass_mask >> j) & 1u) order[n_pm++] = j; }
        for (int q = 0; q < n_pm; ++q) {
            const int j = order[q];
            const double* r_ij = bpos[j];
            double n_ij = norm3(r_ij);
            double r_ij3 = n_ij * n_ij * n_ij;
            double r_j[3] = { r[0] - r_ij[0], r[1] - r_ij[1], r[2] - r_ij[2] };
            double n_j = norm3(r_j);
            double r_j3 = n_j * n_j * n_j;
            double nmu = -dyn->bodies[j].mu_km3_s2;
            for (int i = 0; i < 3; ++i) dx[i] += nmu * (r_j[i] / r_j3 + r_ij[i] / r_ij3);
        }
        for (int i = 0; i < 3; ++i) acc[i] += dx[i];
    }
    for (int f = 0; f < cx->n_grav; ++f) {
        /* gravity_field.rs:149-154: transform_to(self.grav_data.frame): translation to the field's body, then the body-fixed rotation */
        double ga[3], rel[3] = { r[0], r[1], r[2] };
        const int b = cx->grav_body[f];
        if (b >= 0) { rel[0] -= bpos[b][0]; rel[1] -= bpos[b][1]; rel[2] -= bpos[b][2]; }
        nyx_oracle_grav_accel(cx->grav[f], t_ns, rel, cx->grav_scratch, ga);
        for (int i = 0; i < 3; ++i) acc[i] += ga[i];
    }

    /* spacecraft.rs:238-243: force models in order [SolarPressure, Drag], each / mass */
    if (dyn->srp) {
        /* solarpressure.rs:135-165 */
        const nyxb_srp* sp = dyn->srp;
        const double* sun = bpos[sp->sun_body];
        double r_sun[3] = { r[0] - sun[0], r[1] - sun[1], r[2] - sun[2] }; /* s/c seen from the Sun */
        double n_sun = norm3(r_sun);
        double unit[3] = { r_sun[0] / n_sun, r_sun[1] / n_sun, r_sun[2] / n_sun };
        /* cosmic/eclipse.rs:69-83: max occultation over shadow bodies */
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
        double r_sun_au = n_sun / AU_KM;
        double inv = 1.0 / r_sun_au;
        double flux_pressure = (k * sp->phi_w_m2 / SPEED_OF_LIGHT_M_S) * (inv * inv);
        double scal = 1e-3 * cr * cx->srp_area * flux_pressure;
        for (int i = 0; i < 3; ++i) acc[i] += (scal * unit[i]) / mass;
    }
    if (dyn->drag) {
        /* drag.rs:181-284, as coded (incl. its unit/frame quirks) */
        const nyxb_drag* dg = dyn->drag;
        double R[9], wdot;
        nyx_oracle_rotation(&dg->rot, t_ns, R, &wdot);
        double rb[3], vb[3];
        for (int i = 0; i < 3; ++i) {
            rb[i] = (R[3 * i] * r[0] + R[3 * i + 1] * r[1]) + R[3 * i + 2] * r[2];
            vb[i] = (R[3 * i] * v[0] + R[3 * i + 1] * v[1]) + R[3 * i + 2] * v[2];
        }
        /* v_bf = R v + dR/dt r = R v - w x r_bf  (w = wdot * z) */
        vb[0] = vb[0] + wdot * rb[1];
        vb[1] = vb[1] - wdot * rb[0];
        double rho, vel[3];
        if (dg->density == NYXB_DENSITY_CONSTANT) {
            rho = dg->rho0;
            vel[0] = vb[0]; vel[1] = vb[1]; vel[2] = vb[2];               /* drag.rs:193-203 */
        } else {
            double rmag_bf = norm3(rb);
            if (dg->density == NYXB_DENSITY_EXPONENTIAL) {
                rho = dg->rho0 * exp(-(rmag_bf - (dg->r0 + dg->r_eq_km)) / dg->ref_alt_m); /* drag.rs:210-219 */
            } else {
                double alt = rmag_bf - dg->r_eq_km;                       /* drag.rs:241-264 */
                if (alt > dg->ref_alt_m / 1000.0) {
                    rho = pow(10.0, (-7e-5) * alt - 14.464);
                } else {
                    double sc = (alt - 526.8000) / 292.8563;
                    double s2 = sc * sc, s3 = s2 * sc, s4 = s3 * sc, s5 = s4 * sc, s6 = s5 * sc;
                    double logd = 0.34047 * s6 - 0.5889 * s5 - 0.5269 * s4 + 1.0036 * s3 + 0.60713 * s2 - 2.3024 * sc - 12.575;
                    rho = pow(10.0, logd);
                }
            }
            /* drag.rs:223-230: v(integration frame) - v(drag frame), mixed bases as coded */
            vel[0] = v[0] - vb[0]; vel[1] = v[1] - vb[1]; vel[2] = v[2] - vb[2];
        }
        double scal = -0.5 * 1e3 * rho * cd * cx->drag_area * norm3(vel);
        for (int i = 0; i < 3; ++i) acc[i] += (scal * vel[i]) / mass;
    }

    dy[0] = v[0]; dy[1] = v[1]; dy[2] = v[2];
    dy[3] = acc[0]; dy[4] = acc[1]; dy[5] = acc[2];
    dy[6] = 0.0; dy[7] = 0.0; dy[8] = 0.0; /* no guidance law: spacecraft.rs:248 */
    return 0;
}

/* ------------------------------------------------------------------------- */
/* ErrorControl::estimate — propagators/error_ctrl.rs:79-230                  */
/* ------------------------------------------------------------------------- */
static double rss_step3(const double* e, const double* cand, const double* cur) {
    double d[3] = { cand[0] - cur[0], cand[1] - cur[1], cand[2] - cur[2] };
    double mag = norm3(d), err = norm3(e);
    return (mag > sqrt(0.1)) ? err / mag : err;              /* :186-203 */
}
static double rss_state3(const double* e, const double* cand, const double* cur) {
    double s[3] = { cand[0] + cur[0], cand[1] + cur[1], cand[2] + cur[2] };
    double mag = 0.5 * norm3(s), err = norm3(e);
    return (mag > 0.1) ? err / mag : err;                    /* :217-230 */
}
/* nalgebra's generic `norm()` on the 90-vector whose entries 9.. are zero
 * (cosmic/spacecraft.rs:449-473, stm = None): 8 interleaved accumulators folded
 * as (a0+a4),(a1+a5),(a2+a6),(a3+a7) — restated from nalgebra 0.35 `dotx`
 * (not in the tree; unpinned by any reference vector). */
static double norm9_nalgebra(const double v[9]) {
    double acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = v[i] * v[i];
    acc[0] += v[8] * v[8];
    double res = 0.0;
    res += acc[0] + acc[4];
    res += acc[1] + acc[5];
    res += acc[2] + acc[6];
    res += acc[3] + acc[7];
    return sqrt(res);
}

double nyx_oracle_error_estimate(int ctrl, const double err[9], const double cand[9], const double cur[9]) {
    switch (ctrl) {
    case NYXB_RSS_CARTESIAN_STATE: { /* :89-105 */
        double a = rss_state3(err, cand, cur), b = rss_state3(err + 3, cand + 3, cur + 3);
        return (a > b || b != b) ? a : b; /* f64::max */
    }
    case NYXB_RSS_CARTESIAN_STEP: { /* :106-122 */
        double a = rss_step3(err, cand, cur), b = rss_step3(err + 3, cand + 3, cur + 3);
        return (a > b || b != b) ? a : b;
    }
    case NYXB_RSS_STATE: { /* :123-127 */
        double s[9]; for (int i = 0; i < 9; ++i) s[i] = cand[i] + cur[i];
        double mag = 0.5 * norm9_nalgebra(s), e = norm9_nalgebra(err);
        return (mag > 0.1) ? e / mag : e;
    }
    case NYXB_RSS_STEP: { /* :128-136 */
        double d[9]; for (int i = 0; i < 9; ++i) d[i] = cand[i] - cur[i];
        double mag = norm9_nalgebra(d), e = norm9_nalgebra(err);
        return (mag > sqrt(0.1)) ? e / mag : e;
    }
    case NYXB_LARGEST_ERROR: { /* :137-151 (signed comparison, as coded) */
        double max_err = 0.0;
        for (int i = 0; i < 9; ++i) {
            double delta = cand[i] - cur[i];
            double e = (delta > 0.1) ? fabs(err[i] / delta) : fabs(err[i]);
            if (e > max_err) max_err = e;
        }
        return max_err;
    }
    case NYXB_LARGEST_STATE: { /* :152-162 */
        double mag = 0.0, e = 0.0;
        for (int i = 0; i < 9; ++i) { mag += 0.5 * fabs(cand[i] + cur[i]); e += fabs(err[i]); }
        return (mag > 0.1) ? e / mag : e;
    }
    default: { /* LargestStep :163-173 */
        double mag = 0.0, e = 0.0;
        for (int i = 0; i < 9; ++i) { mag += fabs(cand[i] - cur[i]); e += fabs(err[i]); }
        return (mag > 0.1) ? e / mag : e;
    }
    }
}

/* ------------------------------------------------------------------------- */
/* PropInstance — propagators/instance.rs                                     */
/* ------------------------------------------------------------------------- */
typedef struct {
    double y[9];
    int64_t epoch_ns;
    int64_t step_ns;    /* instance.rs:56 step_size (Duration) */
    int fixed;          /* instance.rs:57 */
    nyxb_details det;
    int status;
    /* trajectory recording: the channel of for_duration_with_traj (instance.rs:297-326) */
    const nyxb_traj_sink* sink;
    size_t idx, n;
    /* stop condition of until_nth_event (event.rs:88-211) */
    const nyxb_event* ev;
    double ev_prev;
    int ev_count;
} inst_t;

/* Event scalar minus the desired value for the closed set of include/nyxb.h (anise's Event::eval with
 * Condition::Equals(value) is `scalar - value`; anise is not in the tree). */
static double event_eval(const nyxb_event* ev, const double* y) {
    double s;
    switch (ev->kind) {
    case NYXB_EVENT_RMAG: s = sqrt((y[0] * y[0] + y[1] * y[1]) + y[2] * y[2]); break;
    case NYXB_EVENT_RDOTV: s = (y[0] * y[3] + y[1] * y[4]) + y[2] * y[5]; break;
    case NYXB_EVENT_X: s = y[0]; break;
    case NYXB_EVENT_Y: s = y[1]; break;
    case NYXB_EVENT_Z: s = y[2]; break;
    default: s = sqrt((y[3] * y[3] + y[4] * y[4]) + y[5] * y[5]); break;
    }
    return s - ev->value;
}

static void record_state(const inst_t* in, int64_t s) {
    if (!in->sink || s >= in->sink->capacity) return;
    in->sink->epoch_ns[(size_t)s * in->n + in->idx] = in->epoch_ns;
    for (int c = 0; c < 6; ++c) in->sink->state[((size_t)c * in->sink->capacity + s) * in->n + in->idx] = in->y[c];
}

#define MAX_STAGES 16

/* Sensitivity probe (tests/test_oracle_sensitivity.py, scripts/oracle_sensitivity.py): the error norm of every adaptive attempt
 * is multiplied by this factor.  1.0 (the default) leaves the restatement untouched; 1 + 2^-52 asks "what does ONE ulp in the
 * reference's own error estimate do to the final state" -- the step-sequence sensitivity that bounds any tolerance-parity mode. */
double nyx_oracle_error_scale_ = 1.0;   /* shared with nyx_oracle_od.c */
void nyx_oracle_set_error_scale(double s) { nyx_oracle_error_scale_ = s; }

/* instance.rs:358-493 derive(); returns status, writes dt_ns and next[9] */
static int derive(inst_t* in, eom_ctx* cx, const nyxb_integ_opts* o, const tableau_t* tb, int64_t* dt_ns, double next[9]) {
    double k[MAX_STAGES][9];
    const double* y = in->y;
    int S = tb->stages;
    in->det.attempts = 1;                                        /* :365 */
    double h = nyx_oracle_dur_to_seconds(in->step_ns);           /* :367 */
    const double min_s = nyx_oracle_dur_to_seconds(o->min_step_ns);
    const double max_s = nyx_oracle_dur_to_seconds(o->max_step_ns);
    for (;;) {
        int rc = eom(cx, in->epoch_ns, 0.0, y, k[0]);            /* :369-374 */
        if (rc) return rc;
        int a_idx = 0;
        for (int i = 0; i < S - 1; ++i) {                        /* :376-400 */
            double ci = 0.0, wi[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j <= i; ++j) {
                double a_ij = tb->a[a_idx++];
                ci += a_ij;
                for (int e = 0; e < 9; ++e) wi[e] += a_ij * k[j][e];
            }
            double ys[9];
            for (int e = 0; e < 9; ++e) ys[e] = y[e] + h * wi[e];
            rc = eom(cx, in->epoch_ns, ci * h, ys, k[i + 1]);
            if (rc) return rc;
        }
        double err_est[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int e = 0; e < 9; ++e) next[e] = y[e];
        for (int i = 0; i < S; ++i) {                            /* :407-414 */
            double b_i = tb->b[i];
            if (!in->fixed) {
                double b_s = tb->b[i + S];
                double cf = h * (b_i - b_s);
                for (int e = 0; e < 9; ++e) err_est[e] += cf * k[i][e];
            }
            double cb = h * b_i;
            for (int e = 0; e < 9; ++e) next[e] += cb * k[i][e];
        }
        if (in->fixed) {                                         /* :416-419 */
            in->det.step_ns = in->step_ns;
            *dt_ns = in->step_ns;
            return 0;
        }
        in->det.error = nyx_oracle_error_estimate(o->error_ctrl, err_est, next, y); /* :422-426 */
        if (nyx_oracle_error_scale_ != 1.0) in->det.error *= nyx_oracle_error_scale_;   /* sensitivity probe only */
        if (in->det.error <= o->tolerance || h <= min_s || in->det.attempts >= o->attempts) { /* :428-431 */
            for (int e = 0; e < 9; ++e)
                if (next[e] != next[e]) return NYXB_ERR_PROP_MATH;   /* :432-439 */
            if (in->det.attempts >= o->attempts) in->status |= NYXB_WARN_MAX_ATTEMPTS; /* :440-445 */
            in->det.step_ns = nyx_oracle_dur_from_seconds(h);    /* :447 */
            if (in->det.error < o->tolerance) {                  /* :448-462 */
                double proposed = 0.9 * h * pow(o->tolerance / in->det.error, 1.0 / (double)tb->order);
                if (fabs(proposed) > fabs(max_s)) {
                    double sg = (proposed != proposed) ? proposed : (signbit(proposed) ? -1.0 : 1.0);
                    h = max_s * sg;
                } else {
                    h = proposed;
                }
            }
            in->step_ns = nyx_oracle_dur_from_seconds(h);        /* :464 */
            int64_t ab = in->step_ns < 0 ? -in->step_ns : in->step_ns;
            if (ab < o->min_step_ns)                             /* :465-473 */
                in->step_ns = (in->step_ns < 0) ? -o->min_step_ns : o->min_step_ns;
            *dt_ns = in->det.step_ns;
            return 0;
        }
        in->det.attempts += 1;                                   /* :475-490 */
        in->det.n_rejected += 1;
        double proposed = 0.9 * h * pow(o->tolerance / in->det.error, 1.0 / (double)(tb->order - 1));
        h = (proposed < min_s) ? min_s : proposed;
    }
}

/* spacecraft.rs:158-189 finally(): only the prop-mass check applies without guidance */
static int finally_check(const inst_t* in) {
    return (in->y[8] < 0.0) ? NYXB_ERR_FUEL_EXHAUSTED : 0;
}

/* instance.rs:343-352 single_step() */
static int single_step(inst_t* in, eom_ctx* cx, const nyxb_integ_opts* o, const tableau_t* tb) {
    int64_t dt; double next[9];
    int rc = derive(in, cx, o, tb, &dt, next);
    if (rc) return rc;
    in->epoch_ns += dt;
    for (int e = 0; e < 9; ++e) in->y[e] = next[e];
    /* State::set clamps Cr (cosmic/spacecraft.rs:494) */
    in->y[6] = in->y[6] < 0.0 ? 0.0 : (in->y[6] > 2.0 ? 2.0 : in->y[6]);
    in->det.n_steps += 1;
    record_state(in, in->det.n_steps);                           /* instance.rs:186-193, 255-259 */
    return finally_check(in);
}

/* instance.rs:87-262 propagate(); the channel is record_state() inside single_step, the stop condition is in->ev */
static int propagate(inst_t* in, eom_ctx* cx, const nyxb_integ_opts* o, const tableau_t* tb, int64_t duration_ns) {
    if (duration_ns == 0) return 0;                              /* :96-98 */
    int64_t stop = in->epoch_ns + duration_ns;
    int rc = finally_check(in);                                  /* :106-110 */
    if (rc) return rc;
    int backprop = duration_ns < 0;
    if (backprop) in->step_ns = -in->step_ns;                    /* :112-115 */
    for (;;) {
        int64_t epoch = in->epoch_ns;
        if ((!backprop && epoch + in->step_ns > stop) || (backprop && epoch + in->step_ns <= stop)) { /* :151-153 */
            if (stop == epoch) return 0;                         /* :156-179 */
            int64_t prev_step = in->step_ns; int prev_fixed = in->fixed; /* :182-184 */
            in->step_ns = stop - epoch; in->fixed = 1;
            rc = single_step(in, cx, o, tb);
            if (rc) return rc;
            in->step_ns = prev_step; in->fixed = prev_fixed;     /* :196 */
            if (backprop) in->step_ns = -in->step_ns;            /* :198-200 */
            return 0;
        }
        rc = single_step(in, cx, o, tb);                         /* :241 */
        if (rc) return rc;
        if (in->ev) {                                            /* :243-252 stop condition == event.rs:120-150 closure */
            double y_next = event_eval(in->ev, in->y);
            if (in->ev_prev * y_next < 0.0) in->ev_count += 1;   /* event.rs:141-144 (non-angle scalars) */
            in->ev_prev = y_next;
            if (in->ev_count >= in->ev->trigger) return 0;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Batch driver == MonteCarlo::resume_run_until_epoch's par_iter              */
/* (mc/montecarlo.rs:233-253): independent runs, OpenMP dynamic schedule.     */
/* ------------------------------------------------------------------------- */
int nyx_oracle_propagate_batch_event(const nyxb_dynamics* dyn, const nyxb_integ_opts* opts, size_t n,
                                     const double* state_soa, const double* consts_soa,
                                     const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                                     double* out_state_soa, int64_t* out_epoch_ns,
                                     nyxb_details* out_details, int32_t* out_status, const nyxb_traj_sink* sink,
                                     const nyxb_event* event, int n_threads) {
    tableau_t tb;
    if (event && event->kind == NYXB_EVENT_NONE) event = NULL;
    if (tableau_for(opts->method, &tb)) return -1;
    if (dyn->n_bodies > NYXB_MAX_BODIES) return -1;
    if (opts->state_center < 0 || opts->state_center > dyn->n_bodies) return -1;
    const nyxb_body* shift = opts->state_center ? &dyn->bodies[opts->state_center - 1] : NULL;
    eom_ctx proto;
    const size_t scratch_n = ctx_fields_new(&proto, dyn);
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
#pragma omp parallel num_threads(n_threads)
    {
        double* scratch = scratch_n ? (double*)malloc(sizeof(double) * scratch_n) : NULL;
#pragma omp for schedule(dynamic, 1)
        for (long long ii = 0; ii < (long long)n; ++ii) {
            size_t i = (size_t)ii;
            inst_t in;
            memset(&in, 0, sizeof(in));
            for (int e = 0; e < 9; ++e) in.y[e] = state_soa[(size_t)e * n + i];
            in.epoch_ns = epoch0_ns[i];
            int shift_rc = 0;
            if (shift) {   /* instance.rs:117-142: transform the state into the integration frame (same axes: a translation) */
                double bp[3], bv[3];
                if (nyx_oracle_body_position(shift, in.epoch_ns, bp) || nyx_oracle_body_velocity(shift, in.epoch_ns, bv)) shift_rc = NYXB_ERR_EPHEMERIS;
                else for (int c = 0; c < 3; ++c) { in.y[c] = in.y[c] + 1.0 * bp[c]; in.y[3 + c] = in.y[3 + c] + 1.0 * bv[c]; }
            }
            /* propagator.rs:88-108 with(): step = opts.init_step, fixed = opts.fixed_step */
            in.step_ns = step_ns ? step_ns[i] : opts->init_step_ns;
            in.fixed = opts->fixed_step;
            in.det.step_ns = opts->init_step_ns; in.det.error = 0.0; in.det.attempts = 1;
            in.sink = (sink && sink->capacity > 0) ? sink : NULL; in.idx = i; in.n = n;
            record_state(&in, 0);                                /* start state: instance.rs:307, 321 */
            in.ev = event; in.ev_count = 0;
            if (event) in.ev_prev = event_eval(event, in.y);     /* event.rs:111-113 */
            eom_ctx cx = proto;
            cx.dyn = dyn; cx.grav_scratch = scratch; cx.n_rhs = 0;
            cx.dry_mass = consts_soa[0 * n + i]; cx.extra_mass = consts_soa[1 * n + i];
            cx.srp_area = consts_soa[2 * n + i]; cx.drag_area = consts_soa[3 * n + i];
            int rc = propagate(&in, &cx, opts, &tb, end_epoch_ns - in.epoch_ns); /* instance.rs:279-282 */
            if (shift) {   /* instance.rs:167-176, 211-220: transform back at the epoch the run ended at */
                double bp[3], bv[3];
                if (nyx_oracle_body_position(shift, in.epoch_ns, bp) || nyx_oracle_body_velocity(shift, in.epoch_ns, bv)) { if (!rc) rc = NYXB_ERR_EPHEMERIS; }
                else for (int c = 0; c < 3; ++c) { in.y[c] = in.y[c] + -1.0 * bp[c]; in.y[3 + c] = in.y[3 + c] + -1.0 * bv[c]; }
            }
            (void)shift_rc;
            if (event) {
                event->crossings[i] = in.ev_count;
                if (rc == 0 && in.ev_count < event->trigger) rc = NYXB_ERR_EVENT_NOT_FOUND; /* event.rs:177-182 */
            }
            in.status = (in.status & NYXB_WARN_MAX_ATTEMPTS) | rc;
            in.det.n_rhs = cx.n_rhs;
            for (int e = 0; e < 9; ++e) out_state_soa[(size_t)e * n + i] = in.y[e];
            out_epoch_ns[i] = in.epoch_ns;
            if (step_ns) step_ns[i] = in.step_ns;
            if (out_details) out_details[i] = in.det;
            out_status[i] = in.status;
            if (in.sink) in.sink->count[i] = (in.det.n_steps + 1 < in.sink->capacity) ? in.det.n_steps + 1 : in.sink->capacity;
        }
        free(scratch);
    }
    ctx_fields_free(&proto);
    return 0;
}

int nyx_oracle_propagate_batch_traj(const nyxb_dynamics* dyn, const nyxb_integ_opts* opts, size_t n,
                                    const double* state_soa, const double* consts_soa,
                                    const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                                    double* out_state_soa, int64_t* out_epoch_ns,
                                    nyxb_details* out_details, int32_t* out_status, const nyxb_traj_sink* sink, int n_threads) {
    return nyx_oracle_propagate_batch_event(dyn, opts, n, state_soa, consts_soa, epoch0_ns, end_epoch_ns, step_ns, out_state_soa,
                                            out_epoch_ns, out_details, out_status, sink, NULL, n_threads);
}

int nyx_oracle_propagate_batch(const nyxb_dynamics* dyn, const nyxb_integ_opts* opts, size_t n,
                               const double* state_soa, const double* consts_soa,
                               const int64_t* epoch0_ns, int64_t end_epoch_ns, int64_t* step_ns,
                               double* out_state_soa, int64_t* out_epoch_ns,
                               nyxb_details* out_details, int32_t* out_status, int n_threads) {
    return nyx_oracle_propagate_batch_traj(dyn, opts, n, state_soa, consts_soa, epoch0_ns, end_epoch_ns, step_ns, out_state_soa,
                                           out_epoch_ns, out_details, out_status, NULL, n_threads);
}

/* Direct RHS access for unit tests (one evaluation of SpacecraftDynamics::eom). */
int nyx_oracle_eom(const nyxb_dynamics* dyn, int64_t epoch_ns, double delta_t_s, const double y[9],
                   const double consts[4], double dy[9]) {
    eom_ctx cx;
    const size_t scratch_n = ctx_fields_new(&cx, dyn);
    double* scratch = scratch_n ? (double*)malloc(sizeof(double) * scratch_n) : NULL;
    cx.dyn = dyn; cx.grav_scratch = scratch; cx.n_rhs = 0;
    cx.dry_mass = consts[0]; cx.extra_mass = consts[1]; cx.srp_area = consts[2]; cx.drag_area = consts[3];
    int rc = eom(&cx, epoch_ns, delta_t_s, y, dy);
    free(scratch);
    ctx_fields_free(&cx);
    return rc;
}

int nyx_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
