// opnd_probe.cu — operand-delivery micro-benchmark for the harmonic column walk.
// Question: with 13 FP64 instructions per 40-byte coefficient record and entry, which delivery path lets the FP64 pipe
// (64 lanes/clk/SM) run at its peak?  Variants, all with the walk's real arithmetic (same dependency structure):
//   0  shared memory, lane-varying addresses l & 7 (K2 today: four 8-lane groups read the same 128 B)   2 x LDS.128 + LDS.64
//   1  shared memory, warp-uniform address (transposed mapping: lane = trajectory)
//   2  __constant__ bank, warp-uniform runtime index (LDC / ULDC)
//   3  global memory through the read-only path, warp-uniform address
//   4  as 0 with every record applied to TWO trajectories per lane (T = 2)
//   5  as 1 with T = 2
//   6  as 1 with the recursion coefficients (2n+1), (n+m)(n-m) delivered with the record too (56 B, 11 FP64 + 1)
//   7  as 2 launched as one-warp CTAs whose start entries are staggered by blockIdx (constant-cache behaviour when the
//      warps of an SM are at different places of the table); 1/5/6 stagger by warp index
// Output: SM clocks per warp-entry (per trajectory-entry for T = 2) at 4, 8 and 16 resident warps per SM; the FP64 floor is
// 13 instr x 0.5 clk = 6.5 clk (11 x 0.5 = 5.5 where the recursion coefficients come with the record).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/opnd_probe scripts/opnd_probe.cu && /tmp/opnd_probe
#include <cstdio>
#include <cuda_runtime.h>

#define ENTRIES 256           // records per pass (21x21: 253)
__constant__ double ctab[ENTRIES * 5 + 16];

__device__ __forceinline__ double2 lds128(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

template <int VAR, int T>
__global__ void __launch_bounds__(128) probe(const double* __restrict__ gtab, double* out, int passes, int entries, double ub0, double r20) {
    extern __shared__ __align__(16) double stab[];   // [entry][5][8 lanes] for VAR 0/4, [entry][5] (+pad to 6) otherwise
    const int l = threadIdx.x & 31;
    const int per = (VAR == 0 || VAR == 4) ? 8 : (VAR == 6 ? 2 : 1);
    for (int i = threadIdx.x; i < entries * 6 * per; i += blockDim.x) stab[i] = 1e-3 + 1e-9 * (i % 977);
    __syncthreads();
    double S1[T], S2[T], S3[T], S4[T], S5[T], S6[T], Q1[T], Q2[T], ub[T], r2[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        S1[t] = S2[t] = S3[t] = S4[t] = S5[t] = S6[t] = 0.0;
        Q1[t] = 1.0 + 1e-3 * (threadIdx.x + t); Q2[t] = 0.5;
        ub[t] = ub0 * (1.0 + 1e-6 * threadIdx.x); r2[t] = r20;
    }
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(stab);
    const int wid = (VAR == 7) ? (int)blockIdx.x : (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
    const int e0 = (VAR == 1 || VAR == 5 || VAR == 6 || VAR == 7) ? (wid * 37) % entries : 0;
    for (int p = 0; p < passes; ++p) {
        double al = 3.0, be = 0.0;
#pragma unroll 2
        for (int ee = 0; ee < entries; ++ee) {
            int e = ee + e0;
            if (e >= entries) e -= entries;
            double p1, p2, p3, p4, kk, cal = 0.0, cbe = 0.0;
            if (VAR == 0 || VAR == 4) {
                const unsigned a = sbase + (unsigned)e * (6 * 8 * 8) + (unsigned)(l & 7) * 16;
                const double2 a0 = lds128(a), a1 = lds128(a + 128);
                kk = lds64(sbase + (unsigned)e * (6 * 8 * 8) + 256 + (unsigned)(l & 7) * 8);
                p1 = a0.x; p2 = a0.y; p3 = a1.x; p4 = a1.y;
            } else if (VAR == 1 || VAR == 5) {
                const unsigned a = sbase + (unsigned)e * 48;
                const double2 a0 = lds128(a), a1 = lds128(a + 16);
                kk = lds64(a + 32);
                p1 = a0.x; p2 = a0.y; p3 = a1.x; p4 = a1.y;
            } else if (VAR == 6) {
                const unsigned a = sbase + (unsigned)e * 64;
                const double2 a0 = lds128(a), a1 = lds128(a + 16), a2 = lds128(a + 32);
                kk = lds64(a + 48);
                p1 = a0.x; p2 = a0.y; p3 = a1.x; p4 = a1.y; cal = a2.x; cbe = a2.y;
            } else if (VAR == 2 || VAR == 7) {
                p1 = ctab[e * 5]; p2 = ctab[e * 5 + 1]; p3 = ctab[e * 5 + 2]; p4 = ctab[e * 5 + 3]; kk = ctab[e * 5 + 4];
            } else {
                const double2 a0 = __ldg(reinterpret_cast<const double2*>(gtab + e * 6));
                const double2 a1 = __ldg(reinterpret_cast<const double2*>(gtab + e * 6 + 2));
                kk = __ldg(gtab + e * 6 + 4);
                p1 = a0.x; p2 = a0.y; p3 = a1.x; p4 = a1.y;
            }
#pragma unroll
            for (int t = 0; t < T; ++t) {
                S1[t] = fma(Q1[t], p1, S1[t]);
                S2[t] = fma(Q1[t], p2, S2[t]);
                S3[t] = fma(Q1[t], p3, S3[t]);
                S4[t] = fma(Q1[t], p4, S4[t]);
                const double Qa = (VAR == 6) ? fma(cal * ub[t], Q1[t], -((cbe * r2[t]) * Q2[t]))
                                             : fma(al * ub[t], Q1[t], -((be * r2[t]) * Q2[t]));
                const double wa = kk * Qa;
                S5[t] = fma(wa, p3, S5[t]);
                S6[t] = fma(wa, p4, S6[t]);
                Q2[t] = Q1[t];
                Q1[t] = Qa * 1e-3;   // keep the recursion bounded (one more FP64 op than the real walk, counted below)
            }
            be += al;
            al += 2.0;
        }
    }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < T; ++t) s += S1[t] + S2[t] + S3[t] + S4[t] + S5[t] + S6[t] + Q1[t];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int VAR, int T>
static void run(const char* what, int sms, double ghz, const double* gtab, double* out) {
    const int passes = 400, entries = ENTRIES;
    const size_t smem = (size_t)entries * 6 * ((VAR == 0 || VAR == 4) ? 8 : (VAR == 6 ? 2 : 1)) * sizeof(double);
    const int cta = (VAR == 7) ? 32 : 128;
    cudaFuncSetAttribute(probe<VAR, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int ctas = 1; ctas <= 4; ctas *= 2) {   // 4, 8, 16 warps per SM
        probe<VAR, T><<<sms * ctas * (128 / cta), cta, smem>>>(gtab, out, 20, entries, 0.3, 0.9);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        probe<VAR, T><<<sms * ctas * (128 / cta), cta, smem>>>(gtab, out, passes, entries, 0.3, 0.9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t err = cudaGetLastError();
        const double warp_entries = (double)passes * entries * 4 * ctas * T;   // per SM
        std::printf("%-46s T=%d %2d warps/SM %8.3f ms  %6.2f clk per warp-entry%s\n", what, T, 4 * ctas, ms,
                    ms * 1e-3 * ghz * 1e9 / warp_entries, err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const double ghz = p.clockRate * 1e-6;
    std::printf("%s, %d SMs, %.3f GHz (nominal); FP64 floor: 14 instr x 0.5 = 7.0 clk per warp-entry in this probe\n", p.name,
                p.multiProcessorCount, ghz);
    const int sms = p.multiProcessorCount;
    double h[ENTRIES * 6 + 16];
    for (int i = 0; i < ENTRIES * 6 + 16; ++i) h[i] = 1e-3 + 1e-9 * (i % 977);
    cudaMemcpyToSymbol(ctab, h, sizeof(double) * (ENTRIES * 5 + 16));
    double *gtab, *out;
    cudaMalloc(&gtab, sizeof(h));
    cudaMemcpy(gtab, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMalloc(&out, sizeof(double) * sms * 4 * 128);
    run<0, 1>("smem, lane-varying l&7 (K2 today)", sms, ghz, gtab, out);
    run<1, 1>("smem, warp-uniform address", sms, ghz, gtab, out);
    run<2, 1>("__constant__, warp-uniform index", sms, ghz, gtab, out);
    run<3, 1>("global __ldg, warp-uniform address", sms, ghz, gtab, out);
    run<4, 2>("smem, lane-varying l&7, two trajectories", sms, ghz, gtab, out);
    run<5, 2>("smem, warp-uniform, two trajectories", sms, ghz, gtab, out);
    run<6, 1>("smem, warp-uniform, 56-byte record (11+1 FP64)", sms, ghz, gtab, out);
    run<7, 1>("__constant__, one-warp CTAs, staggered", sms, ghz, gtab, out);
    return 0;
}
