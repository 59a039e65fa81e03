// nyxb_od.cuh — device-side data model of the STM / sequential-filter path, shared by
// nyxb_od.cu (kernels, built STRICT and FAST) and nyxb_api.cu (host packing).
#pragma once
#include "nyxb_device.cuh"

struct DevStation {
    double pos[3], up[3];
    double mask_deg;
    DevRotation rot;
    int body, n_types;
    int types[2];
    double noise_var[2], bias[2];
    double body_radius;
};

struct DevOd {
    int variant, msr_size;
    double reject;                 // < 0: no sigma rejection
    long long max_step_ns, eps_ns;
    int snc_enabled, snc_frame;
    double snc_diag[3];
    long long snc_disable_ns;
    int n_stations;
    const DevStation* stations;
    long long n_msr;
    const long long* msr_epoch;    // [m]
    const int* msr_tracker;        // [m]
    const double* obs;             // [m][2][n]
    const double* covar0;          // [81][n]
    // outputs (any of the per-measurement ones may be null)
    double* covar;                 // [81][n]
    double* state_dev;             // [9][n] or null
    double* ratio;                 // [m][2][n]
    double* prefit;                // [m][2][n]
    double* postfit;               // [m][2][n]
    int* flags;                    // [m][n]
    double* est_state;             // [m][9][n]
    double* est_cov;               // [m][9][n]
};

extern "C" cudaError_t nyxb_launch_stm_strict(const DevSetup*, size_t, const double*, const double*, const long long*, long long,
                                              long long*, const double*, double*, long long*, double*, nyxb_details*, int*, cudaStream_t);
extern "C" cudaError_t nyxb_launch_stm_fast(const DevSetup*, size_t, const double*, const double*, const long long*, long long,
                                            long long*, const double*, double*, long long*, double*, nyxb_details*, int*, cudaStream_t);
extern "C" cudaError_t nyxb_launch_od_strict(const DevSetup*, const DevOd*, size_t, const double*, const double*, const long long*,
                                             double*, long long*, nyxb_details*, int*, cudaStream_t);
extern "C" cudaError_t nyxb_launch_od_fast(const DevSetup*, const DevOd*, size_t, const double*, const double*, const long long*,
                                           double*, long long*, nyxb_details*, int*, cudaStream_t);
