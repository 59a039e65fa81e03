/*
 * nyx_oracle_mvn.c — TEST INFRASTRUCTURE.  CPU restatement of the on-device dispersion stream of
 * nyx_b200/csrc/nyxb_mvn.cu: `MvnSpacecraft::sample` (mc/multivariate.rs:298-331,
 * x = sqrt_s_v * z + mean added to the template) with z from Philox4x32-10 (Salmon et al., SC'11, the published
 * constants) keyed by (seed, run index) and a Box-Muller pair per call.  PARITY UNPINNED with respect to the
 * reference's own RNG (rand_pcg Pcg64Mcg + rand_distr ziggurat, not in the tree; the reference's MC tests assert no
 * numbers): this file pins the DEVICE stream, tests/test_mvn.py checks its integers against the Random123
 * known-answer vectors and its moments against the requested covariance.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

void nyx_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* same contract as nyxb_mvn_sample (include/nyxb.h), host arrays */
int nyx_oracle_mvn_sample(uint64_t seed, uint64_t first_index, size_t n, const double templ[9], const double mean[9],
                          const double L[81], double* out_state_soa, double* out_disp_soa) {
    const double TWO_PI = 6.283185307179586476925286766559;
    for (size_t i = 0; i < n; ++i) {
        uint64_t g = first_index + i;
        double z[10];
        for (int j = 0; j < 5; ++j) {
            uint32_t ctr[4] = { (uint32_t)g, (uint32_t)(g >> 32), (uint32_t)j, 0u }, key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) }, r[4];
            nyx_oracle_philox4x32_10(ctr, key, r);
            uint64_t k1 = ((uint64_t)r[0] << 21) ^ (uint64_t)(r[1] >> 11);
            uint64_t k2 = ((uint64_t)r[2] << 21) ^ (uint64_t)(r[3] >> 11);
            double u1 = (double)(k1 + 1) * 1.1102230246251565e-16;
            double u2 = (double)k2 * 1.1102230246251565e-16;
            double rad = sqrt(-2.0 * log(u1));
            z[2 * j] = rad * cos(TWO_PI * u2);
            z[2 * j + 1] = rad * sin(TWO_PI * u2);
        }
        for (int r = 0; r < 9; ++r) {
            double x = 0.0;
            for (int c = 0; c < 9; ++c) x += L[r * 9 + c] * z[c];
            x += mean ? mean[r] : 0.0;
            out_state_soa[(size_t)r * n + i] = templ[r] + x;
            if (out_disp_soa) out_disp_soa[(size_t)r * n + i] = x;
        }
    }
    return 0;
}
