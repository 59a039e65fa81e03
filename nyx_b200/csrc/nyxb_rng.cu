// nyxb_rng.cu — the reference's Monte Carlo dispersion stream, restated (host code).
//
// `MonteCarlo::generate_states` (mc/montecarlo.rs:277-296) draws every run's nine standard normals from ONE serial
// `Pcg64Mcg::new(seed)` through `rand_distr`'s `StandardNormal` (mc/multivariate.rs:298-302).  Neither crate is in the reference tree
// (rand_pcg 0.10, rand_distr 0.6: nyx-core/Cargo.toml:34-37); both algorithms are published:
//   * Pcg64Mcg = Mcg128Xsl64 (O'Neill's PCG family): state <- state * 0x2360ED051FC65DA44385DF649FCCF645 mod 2^128 (state odd),
//     output = rotr64((state >> 64) ^ state, state >> 122).  PINNED: the generator's official known-answer vector (seed 42) is
//     reproduced bit for bit (tests/test_reference_rng.py).
//   * StandardNormal = the ZIGNOR ziggurat (Doornik 2005) with 256 layers, R = 3.6541528853610088, V = 0.00492867323399; layer index
//     from the low 8 bits of one u64, the variate from its high 52 bits mapped to [-1, 1); wedge test with a fresh 53-bit uniform;
//     tail by Marsaglia's method from two open-interval uniforms.  The layer tables are REGENERATED here from the published
//     generator formulas (x_0 = V / f(R), x_1 = R, x_i = f^-1(V / x_{i-1} + f(x_{i-1})), f = exp(-x^2 / 2)); their first and last
//     entries equal the published table's (3.910757959537090045, 3.654152885361008796, 3.449278298560964462, ...,
//     0.215241895913273806, 0).  Parity of the normals is therefore "unpinned" at the last-ulp level of libm's exp / log.
#include <cmath>
#include <cstdint>
#include <mutex>

#include "../../include/nyxb.h"

namespace {
typedef unsigned __int128 u128;
struct Pcg64Mcg {
    u128 state;
    explicit Pcg64Mcg(u128 seed) : state(seed | 1) {}
    uint64_t next_u64() {
        const u128 mult = ((u128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;
        state *= mult;
        const unsigned rot = (unsigned)(state >> 122);
        const uint64_t xsl = (uint64_t)(state >> 64) ^ (uint64_t)state;
        return (xsl >> rot) | (xsl << ((64 - rot) & 63));
    }
};

constexpr double ZIG_R = 3.6541528853610088, ZIG_V = 0.00492867323399;
double zig_x[257], zig_f[257];
std::once_flag zig_once;
void zig_init() {
    auto f = [](double x) { return std::exp(-x * x / 2.0); };
    zig_x[0] = ZIG_V / f(ZIG_R);
    zig_x[1] = ZIG_R;
    for (int i = 2; i < 256; ++i) zig_x[i] = std::sqrt(-2.0 * std::log(ZIG_V / zig_x[i - 1] + f(zig_x[i - 1])));
    zig_x[256] = 0.0;
    for (int i = 0; i <= 256; ++i) zig_f[i] = f(zig_x[i]);
}

// the 52 low bits of `v` as the mantissa of a double with the given exponent: a value in [2^e, 2^(e+1))
inline double float_with_exponent(uint64_t v52, int e) {
    const uint64_t bits = ((uint64_t)(1023 + e) << 52) | (v52 & ((1ULL << 52) - 1));
    double d;
    __builtin_memcpy(&d, &bits, 8);
    return d;
}
inline double open01(Pcg64Mcg& g) { return float_with_exponent(g.next_u64() >> 12, 0) - (1.0 - 2.220446049250313e-16 / 2.0); }
inline double std_uniform(Pcg64Mcg& g) { return (double)(g.next_u64() >> 11) * (1.0 / 9007199254740992.0); }

double standard_normal(Pcg64Mcg& g) {
    for (;;) {
        const uint64_t bits = g.next_u64();
        const int i = (int)(bits & 0xff);
        const double u = float_with_exponent(bits >> 12, 1) - 3.0;   // [-1, 1)
        const double x = u * zig_x[i];
        if (std::fabs(x) < zig_x[i + 1]) return x;
        if (i == 0) {   // tail
            double xt = 1.0, yt = 0.0;
            while (-2.0 * yt < xt * xt) {
                const double a = open01(g), b = open01(g);
                xt = std::log(a) / ZIG_R;
                yt = std::log(b);
            }
            return u < 0.0 ? xt - ZIG_R : ZIG_R - xt;
        }
        if (zig_f[i + 1] + (zig_f[i] - zig_f[i + 1]) * std_uniform(g) < std::exp(-x * x / 2.0)) return x;
    }
}
}  // namespace

extern "C" int32_t nyxb_pcg64mcg_u64(uint64_t seed_lo, uint64_t seed_hi, size_t n, uint64_t* out) {
    if (!out) return NYXB_RC_BAD_ARG;
    Pcg64Mcg g(((u128)seed_hi << 64) | seed_lo);
    for (size_t i = 0; i < n; ++i) out[i] = g.next_u64();
    return NYXB_RC_OK;
}

extern "C" int32_t nyxb_reference_normals(uint64_t seed_lo, uint64_t seed_hi, uint64_t skip, size_t n, double* out_z) {
    if (!out_z && n) return NYXB_RC_BAD_ARG;
    std::call_once(zig_once, zig_init);
    Pcg64Mcg g(((u128)seed_hi << 64) | seed_lo);
    for (uint64_t s = 0; s < skip; ++s)   // `.skip(skip)` of the sample iterator: whole runs are drawn and dropped
        for (int c = 0; c < 9; ++c) (void)standard_normal(g);
    for (size_t i = 0; i < n; ++i)
        for (int c = 0; c < 9; ++c) out_z[i * 9 + c] = standard_normal(g);
    return NYXB_RC_OK;
}

extern "C" int32_t nyxb_ziggurat_tables(double* x257, double* f257) {
    std::call_once(zig_once, zig_init);
    for (int i = 0; i <= 256; ++i) { if (x257) x257[i] = zig_x[i]; if (f257) f257[i] = zig_f[i]; }
    return NYXB_RC_OK;
}
