// smem_probe.cu — micro-benchmark for DESIGN.md §11 item 1(b): how many shared-memory wavefronts does sm_100 spend on a
// warp-wide 128-bit (or 64-bit) shared load when addresses repeat inside the warp?  The cooperative kernel K2 reads its
// coefficient records with `LDS.128` where the four 8-lane groups of a warp request the SAME 128 bytes; if repeated addresses
// are only merged inside a quarter-warp phase that costs 4 wavefronts per load, if they are merged across the warp it costs 1.
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/smem_probe scripts/smem_probe.cu && /tmp/smem_probe
// Output: SM clocks per warp-level load instruction (16 resident warps per SM issuing back to back), per address pattern.
#include <cstdio>
#include <cuda_runtime.h>

#define TAB_PIECES 1024   // 16 KB of 16-byte pieces

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

// PATTERN: which 16-byte piece lane l reads (before the per-iteration offset)
//  0  l & 7      K2 today: 4 groups x 8 lane positions (each quarter-warp reads the same 128 B)
//  1  l          32 distinct pieces (512 B): the no-sharing baseline
//  2  0          the whole warp reads one piece
//  3  l >> 2     4 consecutive lanes share a piece (8 distinct pieces per warp, 2 per quarter-warp)
//  4  l >> 3     one piece per quarter-warp (4 distinct pieces per warp)
//  5  l & 15     2 groups x 16 positions
template <int PATTERN, int WIDE>
__global__ void __launch_bounds__(512) probe(double* out, int iters) {
    __shared__ __align__(16) double2 tab[TAB_PIECES];
    for (int i = threadIdx.x; i < TAB_PIECES; i += blockDim.x) tab[i] = make_double2(1e-9 * i, 2e-9 * i);
    __syncthreads();
    const int l = threadIdx.x & 31;
    int piece;
    switch (PATTERN) {
    case 0: piece = l & 7; break;
    case 1: piece = l; break;
    case 2: piece = 0; break;
    case 3: piece = l >> 2; break;
    case 4: piece = l >> 3; break;
    default: piece = l & 15; break;
    }
    const unsigned base = (unsigned)__cvta_generic_to_shared(tab);
    double acc0 = 0.0, acc1 = 0.0;
    unsigned off = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned addr = base + (((unsigned)piece + off) & (TAB_PIECES - 1)) * 16u;
            if (WIDE) {
                const double2 v = lds128(addr);
                acc0 += v.x; acc1 += v.y;
            } else {
                acc0 += lds64(addr); acc1 += lds64(addr + 8);
            }
            off += 32;
        }
    }
    if (acc0 + acc1 == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = acc0;   // keep the loads alive
}

template <int PATTERN, int WIDE>
static void run(const char* what, int sms, double ghz) {
    const int iters = 20000, warps = 16;
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 512);
    probe<PATTERN, WIDE><<<sms, 512>>>(out, 200);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<PATTERN, WIDE><<<sms, 512>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double loads = (double)iters * 8 * warps * (WIDE ? 1 : 2);   // warp-level load instructions per SM
    std::printf("%-58s %s  %7.3f ms  %6.2f clk per warp load\n", what, WIDE ? "LDS.128" : "2xLDS.64", ms, ms * 1e-3 * ghz * 1e9 / loads);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const double ghz = p.clockRate * 1e-6;
    std::printf("%s, %d SMs, %.3f GHz (nominal)\n", p.name, p.multiProcessorCount, ghz);
    const int sms = p.multiProcessorCount;
    run<0, 1>("l & 7   (K2 today: 4 groups read the same 128 B)", sms, ghz);
    run<1, 1>("l       (32 distinct pieces, 512 B)", sms, ghz);
    run<2, 1>("0       (one piece for the whole warp)", sms, ghz);
    run<3, 1>("l >> 2  (4 neighbouring lanes share a piece)", sms, ghz);
    run<4, 1>("l >> 3  (one piece per quarter-warp)", sms, ghz);
    run<5, 1>("l & 15  (2 groups x 16 positions)", sms, ghz);
    run<0, 0>("l & 7", sms, ghz);
    run<1, 0>("l", sms, ghz);
    run<3, 0>("l >> 2", sms, ghz);
    return 0;
}
