// nyxb_frame.cu — IntegratorOptions.integration_frame (options.rs:60): translation of the caller's states into the integration
// frame before the propagation loop and back after it (instance.rs:117-142, 167-176, 211-220: anise `transform_to` between two
// frames of the same inertial axes = subtracting / adding the position and velocity of one centre relative to the other at the
// state's epoch).  One thread per trajectory, coalesced SoA rows; built without FMA contraction (same bits as the oracle).
#include "nyxb_device.cuh"

__global__ void nyxb_k_frame_shift(const DevBody b, double sign, size_t n, double* __restrict__ state,
                                   const long long* __restrict__ epoch, int* __restrict__ status) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[3], v[3];
    const long long t = epoch[i];
    if (!body_position(b, t, p) || !body_velocity(b, t, v)) {
        if (status && (status[i] & 0xFF) == 0) status[i] |= NYXB_ERR_EPHEMERIS;   // DynamicsAlmanacError: epoch outside the ephemeris
        return;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        state[(size_t)c * n + i] = __dadd_rn(state[(size_t)c * n + i], __dmul_rn(sign, p[c]));
        state[(size_t)(3 + c) * n + i] = __dadd_rn(state[(size_t)(3 + c) * n + i], __dmul_rn(sign, v[c]));
    }
}

extern "C" cudaError_t nyxb_launch_frame_shift(const DevBody* b, double sign, size_t n, double* state, const long long* epoch,
                                               int* status, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    nyxb_k_frame_shift<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(*b, sign, n, state, epoch, status);
    return cudaGetLastError();
}
