// nyxb_coop.h — lane-cooperative harmonic tables shared between nyxb_api.cu (owner of the
// device buffers) and nyxb_coop.cu (builder + kernel).
#pragma once
#include <vector>

#include "nyxb_device.cuh"

// One record per entry (n, m) of the derived-Legendre triangle, m >= 1, m <= n <= N+1,
// laid out [entry e][lane] in the order each lane walks its columns.
struct __align__(16) DevCoopRec {
    double p1, p2;  // sqrt2 * m * (Cbar, Sbar)[n][m]                       -> X, Y sums
    double p3, p4;  // sqrt2 * vr01[n][m-1] * (Cbar, Sbar)[n][m-1]          -> Z sum
    double p5, p6;  // sqrt2 * vr11[n-1][m-1] * (Cbar, Sbar)[n-1][m-1]      -> W sum
    double bq, cq;  // recursion factors producing A[n+1][m] from A[n][m], A[n-1][m]
};

struct DevCoop {
    int G, L, kmax;
    const DevCoopRec* recs;  // [L][G]
    const int* col_start;    // [G][kmax] entry index at which the k-th column of the lane starts (L+1: none)
    const int* col_m;        // [G][kmax] order m of that column
};

struct CoopHost {
    int G = 0, L = 0, kmax = 0;
    std::vector<DevCoopRec> recs;
    std::vector<int> col_start, col_m;
};

// ---- STRICT cooperative kernel (nyxb_coop_strict.cu): column lists for phase 1, degree lists for phase 2
struct DevCoopStrict {
    int G, kc, kr;
    const int* cols;  // [G][kc+1], -1 terminated, ascending m
    const int* rows;  // [G][kr+1], -1 terminated, ascending n
};
struct CoopStrictHost {
    int G = 0, kc = 0, kr = 0;
    std::vector<int> cols, rows;
};
void nyxb_coop_strict_build_host(int N, int M, int G, CoopStrictHost& out);
extern "C" cudaError_t nyxb_launch_coop_strict(const DevSetup* S, const DevCoopStrict* Cs, size_t n, const double* state,
                                               const double* consts, const long long* epoch0, long long end_epoch,
                                               long long* step_io, double* out_state, long long* out_epoch,
                                               nyxb_details* out_details, int* out_status, cudaStream_t stream);

void nyxb_coop_build_host(int N, int M, const double* c_nm, const double* s_nm, int G, CoopHost& out);

// T = trajectories integrated together by one lane group (1 or 2: register blocking over the coefficient records)
extern "C" cudaError_t nyxb_launch_coop(const DevSetup* S, const DevCoop* Cp, int T, size_t n, const double* state,
                                        const double* consts, const long long* epoch0, long long end_epoch,
                                        long long* step_io, double* out_state, long long* out_epoch,
                                        nyxb_details* out_details, int* out_status, cudaStream_t stream);
