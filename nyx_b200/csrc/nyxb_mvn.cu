// nyxb_mvn.cu — on-device Monte Carlo dispersions: `MvnSpacecraft::sample`
// (mc/multivariate.rs:298-331) for a whole ensemble in one launch: x = sqrt_s_v * z + mean added to the template's
// [r, v, Cr, Cd, prop mass], z ~ N(0, I_9).
//
// The reference draws z from ONE serial `Pcg64Mcg` stream through rand_distr's ziggurat (mc/montecarlo.rs:277-296);
// neither crate is in the tree and a rejection sampler cannot be jumped ahead, so the device stream is counter based
// instead: run index g (the reference's `Run.index` = draw order) keys its own draws,
//   Philox4x32-10(key = seed, counter = (g, j)) -> 2 x 53-bit uniforms -> Box-Muller pair,   j = 0..4
// which makes the ensemble independent of how it is sharded over GPUs.  Draw-for-draw parity with the reference RNG
// is out of scope (its MC tests assert no numbers); oracle/nyx_oracle_mvn.c restates THIS stream.
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/nyxb.h"

struct MvnParams {
    double templ[9], mean[9], L[81];
    unsigned long long seed, first;
};

__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void nyxb_k_mvn(const __grid_constant__ MvnParams P, size_t n, double* __restrict__ out_state, double* __restrict__ out_disp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long g = P.first + i;
    double z[10];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        unsigned r[4];
        philox4x32_10((unsigned)g, (unsigned)(g >> 32), (unsigned)j, 0u, (unsigned)P.seed, (unsigned)(P.seed >> 32), r);
        unsigned long long k1 = ((unsigned long long)r[0] << 21) ^ (unsigned long long)(r[1] >> 11);
        unsigned long long k2 = ((unsigned long long)r[2] << 21) ^ (unsigned long long)(r[3] >> 11);
        double u1 = (double)(k1 + 1ULL) * 1.1102230246251565e-16;   // (0, 1]
        double u2 = (double)k2 * 1.1102230246251565e-16;            // [0, 1)
        double rad = sqrt(-2.0 * log(u1));
        double s, c;
        sincospi(2.0 * u2, &s, &c);
        z[2 * j] = rad * c;
        z[2 * j + 1] = rad * s;
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        double x = 0.0;
#pragma unroll
        for (int c = 0; c < 9; ++c) x += P.L[r * 9 + c] * z[c];
        x += P.mean[r];
        out_state[(size_t)r * n + i] = P.templ[r] + x;
        if (out_disp) out_disp[(size_t)r * n + i] = x;
    }
}

extern "C" cudaError_t nyxb_launch_mvn(unsigned long long seed, unsigned long long first, size_t n, const double* templ,
                                       const double* mean, const double* L, double* out_state, double* out_disp, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    MvnParams P;
    for (int i = 0; i < 9; ++i) { P.templ[i] = templ[i]; P.mean[i] = mean ? mean[i] : 0.0; }
    for (int i = 0; i < 81; ++i) P.L[i] = L[i];
    P.seed = seed; P.first = first;
    const int block = 128;
    nyxb_k_mvn<<<(unsigned)((n + block - 1) / block), block, 0, st>>>(P, n, out_state, out_disp);
    return cudaGetLastError();
}
