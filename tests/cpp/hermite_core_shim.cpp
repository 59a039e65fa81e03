// Host build of the resampling kernel's per-(query, trajectory) function (nyx_b200/csrc/nyxb_hermite.h), for the CPU test
// tests/test_trajectory.py::test_resample_core_matches_traj_at: same source as the CUDA kernel compiles, looped on the host.
// Test infrastructure only — not linked into libnyxb.so.
#include "../../nyx_b200/csrc/nyxb_hermite.h"

extern "C" void shim_traj_resample(long long cap, const long long* epoch, const double* state, const long long* count, size_t n,
                                   size_t m, const long long* query, double* out_state, int* out_status) {
    NyxbTrajView tv;
    tv.cap = cap; tv.epoch = epoch; tv.state = state; tv.count = count;
    const double nan = __builtin_nan("");
    for (size_t j = 0; j < m; ++j)
        for (size_t i = 0; i < n; ++i) {
            double rv[6] = {nan, nan, nan, nan, nan, nan};
            out_status[j * n + i] = nyxb_traj_at(tv, n, i, query[j], rv);
            for (int c = 0; c < 6; ++c) out_state[((size_t)c * m + j) * n + i] = rv[c];
        }
}

extern "C" void shim_event_locate(long long cap, const long long* epoch, const double* state, const long long* count, size_t n, int kind,
                                  double value, long long precision_ns, const int* run_status, long long* out_epoch, double* out_state,
                                  int* out_status) {
    NyxbTrajView tv;
    tv.cap = cap; tv.epoch = epoch; tv.state = state; tv.count = count;
    const double nan = __builtin_nan("");
    for (size_t i = 0; i < n; ++i) {
        double rv[6] = {nan, nan, nan, nan, nan, nan};
        long long ev = 0;
        int status = 1;
        if (!run_status || (run_status[i] & 0xff) == 0) {
            status = nyxb_event_locate_one(tv, n, i, kind, value, precision_ns, &ev, rv);
            if (status) { for (int c = 0; c < 6; ++c) rv[c] = nan; ev = 0; }
        }
        out_epoch[i] = ev;
        for (int c = 0; c < 6; ++c) out_state[(size_t)c * n + i] = rv[c];
        out_status[i] = status;
    }
}
