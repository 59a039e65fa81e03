// nyxb_traj.cu — batched Hermite resampling of recorded trajectories on a common epoch grid.
//
// What `Traj::every` / `every_between` (md/trajectory/traj.rs:148-162, traj_it.rs:32-63) do one state at a time on the host —
// `Traj::at` (traj.rs:83-126) = binary search + 13-sample window + Hermite interpolation of (r, v) pairs
// (interpolatable.rs:53-108) — as ONE launch over (query epoch j, trajectory i) reading the step-major SoA sink the
// propagation kernels append to (include/nyxb.h: nyxb_traj_sink).  i is the fastest index: a warp reads 32 neighbouring
// trajectories of one record, 256-byte coalesced rows; the window is 13 records x 7 rows, re-read from L2 by neighbouring
// queries.  HBM bound: algorithmic bytes per (query, trajectory) = 13 x 56 (window, when not shared) + 52 (out).
// Built with -fmad=false: bit-identical to the numpy restatement (nyx_b200/trajectory.py).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nyxb.h"
#include "nyxb_hermite.h"

__global__ void __launch_bounds__(128)
nyxb_k_traj_resample(const NyxbTrajView tv, size_t n, size_t m, const long long* __restrict__ query,
                     double* __restrict__ out_state, int* __restrict__ out_status) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * m) return;
    const size_t i = gid % n, j = gid / n;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double rv[6] = {nan, nan, nan, nan, nan, nan};
    const int status = nyxb_traj_at(tv, n, i, query[j], rv);
#pragma unroll
    for (int c = 0; c < 6; ++c) out_state[((size_t)c * m + j) * n + i] = rv[c];
    out_status[j * n + i] = status;
}

extern "C" cudaError_t nyxb_launch_traj_resample(long long cap, const long long* epoch, const double* state, const long long* count,
                                                 size_t n, size_t m, const long long* query, double* out_state, int* out_status,
                                                 cudaStream_t stream) {
    if (n == 0 || m == 0) return cudaSuccess;
    NyxbTrajView tv;
    tv.cap = cap; tv.epoch = epoch; tv.state = state; tv.count = count;
    const size_t total = n * m;
    const unsigned grid = (unsigned)((total + 127) / 128);
    nyxb_k_traj_resample<<<grid, 128, 0, stream>>>(tv, n, m, query, out_state, out_status);
    return cudaGetLastError();
}

// ---- event location: one thread per trajectory searches the last recorded step (nyxb_hermite.h: nyxb_event_locate_one)
__global__ void __launch_bounds__(128)
nyxb_k_event_locate(const NyxbTrajView tv, size_t n, int kind, double value, long long precision_ns, const int* __restrict__ run_status,
                    long long* __restrict__ out_epoch, double* __restrict__ out_state, int* __restrict__ out_status) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double rv[6] = {nan, nan, nan, nan, nan, nan};
    long long ev = 0;
    int status = NYXB_TRAJ_NO_DATA;
    if (!run_status || (run_status[i] & 0xff) == 0) {   // failed runs have no event to locate
        status = nyxb_event_locate_one(tv, n, i, kind, value, precision_ns, &ev, rv);
        if (status) {
#pragma unroll
            for (int c = 0; c < 6; ++c) rv[c] = nan;
            ev = 0;
        }
    }
    out_epoch[i] = ev;
#pragma unroll
    for (int c = 0; c < 6; ++c) out_state[(size_t)c * n + i] = rv[c];
    out_status[i] = status;
}

extern "C" cudaError_t nyxb_launch_event_locate(long long cap, const long long* epoch, const double* state, const long long* count,
                                                size_t n, int kind, double value, long long precision_ns, const int* run_status,
                                                long long* out_epoch, double* out_state, int* out_status, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    NyxbTrajView tv;
    tv.cap = cap; tv.epoch = epoch; tv.state = state; tv.count = count;
    nyxb_k_event_locate<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(tv, n, kind, value, precision_ns, run_status, out_epoch, out_state,
                                                                          out_status);
    return cudaGetLastError();
}
