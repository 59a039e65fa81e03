// nyxb_tx.h — tables of the TRANSPOSED cooperative kernel (nyxb_tx.cu): lane = trajectory, warp = column position.
//
// 32 trajectories (one "set") are integrated by one CTA of P warps.  Every lane of a warp walks the SAME entries of the
// derived-Legendre triangle for its own trajectory, so the coefficient records are warp-uniform shared-memory loads (LDS.128 with
// one address per warp: 2 clk instead of the 4 clk a lane-varying LDS.128 costs, profiles/r02a_smem_probe.txt) and every column
// boundary is a uniform branch.  The columns m = 1..min(M, N)+1 are dealt to the P positions in a zigzag over the exponent
// e = m - 1 (position of e: r = e mod 2P, r < P ? r : 2P-1-r), which (a) balances the entry counts as well as bin packing does
// (21x21 / 8 positions: 32 of ideal 29, the same maximum LPT reaches) and (b) makes the per-column powers
// (cos, sin)(e lambda) cos^e(phi) and rho^(e+1) two geometric sequences of ratio z^(2P): one complex multiplication per column.
//
// Record of entry (n, m) — same algebra as nyxb_coop.h (un-normalised Q recursion, normalisation folded into the record):
//   recA[e] = {p1, p2, p3, p4}   32 B   p1,p2 = sqrt2 m (C,S)[n][m] scale[n][m];  p3,p4 = sqrt2 vr01[n][m-1] (C,S)[n][m-1] scale[n][m]
//   recK[e] = kappa(n)            8 B   W term of degree n+1 = kappa(n) Q[n+1] (p3, p4)(n)
// laid out position-major, each position's columns in ascending m, one null record behind the last entry (prefetch target).
#pragma once
#include <vector>

#include "nyxb_device.cuh"

#define NYXB_TX_MAXP 16
#define NYXB_TX_KMAX 16   /* columns per position: ceil((N + 1) / P) <= 16 */

struct DevTx {
    int P, n_rec, kmax;
    // ONE device allocation [recA (n_rec+1)x4 | recK n_rec+2 | colseed (N+2)x4 | sched P x (2+2 kmax) ints]: the kernel stages it
    // into shared memory with a single TMA bulk copy.  colseed[m] = {(2m-1)!!, pd1, pd2 (W term of the column's first entry), 2m+1};
    // sched[w] = {rec_off, ncol, (m, len) per column}
    const double* recA;
};

struct TxHost {
    int P = 0, n_rec = 0, kmax = 0;
    std::vector<double> recA, recK, colseed;
    std::vector<int> sched;
};

void nyxb_tx_build_host(int N, int M, const double* c_nm, const double* s_nm, int P, TxHost& out);
// packs the tables into the blob the kernel expects; returns its size (dst == NULL: size only)
size_t nyxb_tx_pack_blob(const TxHost* h, int N, unsigned char* dst);

// work queue + parking area of the persistent kernel (device pointers, owned by the engine; `ctl` and `ring` zeroed before a launch).
// Fresh sets are handed out by a counter; a set parked at the end of a time slice is pushed on a ring of resumable sets and popped
// by whichever context asks next (spin lock around the two ring indices: one push + one pop per context and slice).  A context
// that finds neither a fresh nor a parked set exits: every unfinished set is then in progress in some other context, which will
// pop it again itself after parking it — nothing ever waits for another context's slice to end.
struct DevTxQueue {
    int* ctl;                     // [4]: next fresh set, lock, ring head, ring tail
    int* ring;                    // [n_sets] parked sets
    long long* ws_step;           // [n]   adapted step of a parked trajectory
    double* ws_f64;               // [2][n] raw retry step h, previous event value
    int* ws_flags;                // [n]   fixed | retry << 1 | done << 2 | warn << 3 | rc << 8
    nyxb_details* details;        // [n]   never NULL inside the kernel (user buffer or engine scratch)
    int n_sets, slice;            // slice: step attempts per slice (0: run every set to completion)
    unsigned long long* trace;    // diagnostic builds (-DNYXB_TX_TRACE) only: [warps of CTA 0][NYXB_TX_TRACE_CAP] timeline records
};
#define NYXB_TX_TRACE_CAP 8192
enum { TXQ_FRESH = 0, TXQ_LOCK = 1, TXQ_HEAD = 2, TXQ_TAIL = 3 };

extern "C" cudaError_t nyxb_launch_tx(const DevSetup* S, const DevTx* Tx, const DevTxQueue* q, size_t n, const double* state,
                                      const double* consts, const long long* epoch0, long long end_epoch, long long* step_io,
                                      double* out_state, long long* out_epoch, int* out_status, const DevSink* sink,
                                      int grid, cudaStream_t stream);
// set contexts per SM (one persistent CTA per SM holding 1 or 2 sets; 0: the tables do not fit) and the CTA's dynamic shared memory
extern "C" int nyxb_tx_occupancy(const DevSetup* S, const DevTx* Tx, size_t* smem_bytes);
