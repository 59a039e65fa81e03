// walk_probe.cu — what bounds the column walk of the transposed kernel (nyxb_tx.cu: NYXB_TX_ENTRY) on B200?
// The walk issues 12 FP64 instructions (7 DFMA, 3 DADD, 2 DMUL) and 2.5 warp-uniform shared-memory loads per entry; the FP64 pipe
// takes 2 warp-instructions per clock and SM.  Variants (same dependency structure as the kernel's loop):
//   0  12 independent DFMA per entry, no loads                          (what the pipe delivers at this occupancy)
//   1  the walk's instruction stream, record held in registers          (the FP64 stream alone)
//   2  the walk as built: records from shared memory, warp-uniform LDS.128 / LDS.64, software-pipelined over two register sets
//   3  as 2 with two trajectories per lane (TT = 2)
//   4  as 1 with the DADD / DMUL written as DFMA                        (do add / mul issue at the DFMA rate?)
//   5  as 2 with the kappa stream packed to one LDS.128 per two entries (2.25 loads per entry)
// Output: FP64 warp-instructions per clock and SM (peak 2.0) at 4, 8, 12, 16 warps per SM (one CTA per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/walk_probe scripts/walk_probe.cu && /tmp/walk_probe
#include <cstdio>
#include <cuda_runtime.h>

#define ENTRIES 244   // 21x21: padded entries of all positions

template <int TT>
struct Lane { double Q[TT], c1[TT], g[TT], m2[TT], d[TT], S1[TT], S2[TT], S3[TT], S4[TT], S5[TT], S6[TT], dc[TT], dg[TT]; };

#define ENTRY(P01, P23, PK)                                          \
    _Pragma("unroll") for (int u = 0; u < TT; ++u) {                 \
        const double Qn = fma(L.c1[u], L.Q[u], -L.m2[u]);            \
        L.c1[u] += L.dc[u]; L.d[u] += L.g[u]; L.g[u] += L.dg[u];     \
        L.m2[u] = L.d[u] * L.Q[u];                                   \
        L.S1[u] = fma(L.Q[u], (P01).x, L.S1[u]);                     \
        L.S2[u] = fma(L.Q[u], (P01).y, L.S2[u]);                     \
        L.S3[u] = fma(L.Q[u], (P23).x, L.S3[u]);                     \
        L.S4[u] = fma(L.Q[u], (P23).y, L.S4[u]);                     \
        const double wv = (PK) * Qn;                                 \
        L.S5[u] = fma(wv, (P23).x, L.S5[u]);                         \
        L.S6[u] = fma(wv, (P23).y, L.S6[u]);                         \
        L.Q[u] = Qn;                                                 \
    }
#define ENTRY_FMA(P01, P23, PK)                                      \
    _Pragma("unroll") for (int u = 0; u < TT; ++u) {                 \
        const double Qn = fma(L.c1[u], L.Q[u], -L.m2[u]);            \
        L.c1[u] = fma(one, L.dc[u], L.c1[u]); L.d[u] = fma(one, L.g[u], L.d[u]); L.g[u] = fma(one, L.dg[u], L.g[u]); \
        L.m2[u] = fma(L.d[u], L.Q[u], zero);                         \
        L.S1[u] = fma(L.Q[u], (P01).x, L.S1[u]);                     \
        L.S2[u] = fma(L.Q[u], (P01).y, L.S2[u]);                     \
        L.S3[u] = fma(L.Q[u], (P23).x, L.S3[u]);                     \
        L.S4[u] = fma(L.Q[u], (P23).y, L.S4[u]);                     \
        const double wv = fma((PK), Qn, zero);                       \
        L.S5[u] = fma(wv, (P23).x, L.S5[u]);                         \
        L.S6[u] = fma(wv, (P23).y, L.S6[u]);                         \
        L.Q[u] = Qn;                                                 \
    }

template <int VAR, int TT>
__global__ void __launch_bounds__(512, 1) probe(double* out, long long* clk, int passes, double ub0, double one, double zero) {
    extern __shared__ __align__(16) double stab[];   // recA [ENTRIES+2][4] | recK [ENTRIES+4]
    double2* recA = reinterpret_cast<double2*>(stab);
    double* recK = stab + (ENTRIES + 2) * 4;
    for (int i = threadIdx.x; i < (ENTRIES + 2) * 4 + ENTRIES + 4; i += blockDim.x) stab[i] = 1e-3 + 1e-9 * (i % 977);
    __syncthreads();
    Lane<TT> L;
#pragma unroll
    for (int u = 0; u < TT; ++u) {
        L.Q[u] = 1.0 + 1e-3 * (threadIdx.x + u); L.c1[u] = 3.0 * ub0; L.g[u] = 3e-3; L.m2[u] = 0.0; L.d[u] = 0.0;
        L.S1[u] = L.S2[u] = L.S3[u] = L.S4[u] = L.S5[u] = L.S6[u] = 0.0;
        L.dc[u] = 2.0 * ub0 * 1e-3; L.dg[u] = 2e-6;
    }
    const long long t0 = clock64();
    if (VAR == 0) {
        double a[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) a[k] = 1.0 + 1e-3 * k + 1e-6 * threadIdx.x;
        const double x = 1.0 - 1e-9 * ub0, y = 1e-9;
        for (int p = 0; p < passes; ++p)
#pragma unroll 4
            for (int e = 0; e < ENTRIES; ++e) {
#pragma unroll
                for (int k = 0; k < 12; ++k) a[k] = fma(a[k], x, y);
            }
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 12; ++k) s += a[k];
        L.S1[0] = s;
    } else if (VAR == 1 || VAR == 4) {
        const double2 a01 = recA[threadIdx.x & 1], a23 = recA[2 + (threadIdx.x & 1)];
        const double kk = recK[threadIdx.x & 3];
        for (int p = 0; p < passes; ++p) {
#pragma unroll 4
            for (int e = 0; e < ENTRIES; ++e) {
                if (VAR == 1) { ENTRY(a01, a23, kk) } else { ENTRY_FMA(a01, a23, kk) }
            }
#pragma unroll
            for (int u = 0; u < TT; ++u) { L.Q[u] = 1.0; L.c1[u] = 3.0 * ub0; L.d[u] = 0.0; L.g[u] = 3e-3; L.m2[u] = 0.0; }
        }
    } else if (VAR == 2 || VAR == 3) {
        for (int p = 0; p < passes; ++p) {
            const double2* A = recA;
            const double* K = recK;
            double2 a01 = A[0], a23 = A[1], b01, b23;
            double kk = K[0], bk;
#pragma unroll 2
            for (int e = ENTRIES; e > 0; e -= 2) {
                b01 = A[2]; b23 = A[3]; bk = K[1];
                ENTRY(a01, a23, kk)
                a01 = A[4]; a23 = A[5]; kk = K[2];
                ENTRY(b01, b23, bk)
                A += 4; K += 2;
            }
#pragma unroll
            for (int u = 0; u < TT; ++u) { L.Q[u] = 1.0; L.c1[u] = 3.0 * ub0; L.d[u] = 0.0; L.g[u] = 3e-3; L.m2[u] = 0.0; }
        }
    } else {   // VAR 5: kappa pairs by LDS.128
        const double2* K2 = reinterpret_cast<const double2*>(recK);
        for (int p = 0; p < passes; ++p) {
            const double2* A = recA;
            const double2* K = K2;
            double2 a01 = A[0], a23 = A[1], b01, b23, kp = K[0];
#pragma unroll 2
            for (int e = ENTRIES; e > 0; e -= 2) {
                b01 = A[2]; b23 = A[3];
                ENTRY(a01, a23, kp.x)
                a01 = A[4]; a23 = A[5];
                const double bk = kp.y;
                kp = K[1];
                ENTRY(b01, b23, bk)
                A += 4; K += 1;
            }
#pragma unroll
            for (int u = 0; u < TT; ++u) { L.Q[u] = 1.0; L.c1[u] = 3.0 * ub0; L.d[u] = 0.0; L.g[u] = 3e-3; L.m2[u] = 0.0; }
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < TT; ++u) s += L.S1[u] + L.S2[u] + L.S3[u] + L.S4[u] + L.S5[u] + L.S6[u] + L.Q[u];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int VAR, int TT>
void run(const char* name, int sms) {
    double* out; long long* clk;
    cudaMalloc(&out, 8); cudaMalloc(&clk, 8 * sms);
    const int passes = 200;
    const size_t smem = ((ENTRIES + 2) * 4 + ENTRIES + 4) * 8;
    printf("%-66s", name);
    for (int warps : {4, 8, 12, 16}) {
        probe<VAR, TT><<<sms, warps * 32, smem>>>(out, clk, 2, 0.3, 1.0, 0.0);
        probe<VAR, TT><<<sms, warps * 32, smem>>>(out, clk, passes, 0.3, 1.0, 0.0);
        cudaDeviceSynchronize();
        long long h[256];
        cudaMemcpy(h, clk, 8 * sms, cudaMemcpyDeviceToHost);
        double mean = 0; for (int i = 0; i < sms; ++i) mean += (double)h[i]; mean /= sms;
        const double instr = 12.0 * TT * ENTRIES * passes * warps;
        printf("  %2dw %.3f", warps, instr / mean);
    }
    printf("\n");
    cudaFree(out); cudaFree(clk);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs; FP64 warp-instructions per clock and SM (pipe peak 2.0) at 4 / 8 / 12 / 16 warps per SM\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    run<0, 1>("0 independent DFMA x12, no loads", sms);
    run<1, 1>("1 walk stream, record in registers", sms);
    run<4, 1>("4 walk stream, DADD/DMUL as DFMA, record in registers", sms);
    run<2, 1>("2 walk as built (uniform LDS.128 x2 + LDS.64 per entry)", sms);
    run<5, 1>("5 walk, kappa pairs by LDS.128 (2.25 loads per entry)", sms);
    run<3, 2>("3 walk as built, two trajectories per lane", sms);
    run<1, 2>("1' walk stream, two trajectories per lane, record in registers", sms);
    return 0;
}
