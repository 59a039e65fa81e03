// placeholder until the lane-cooperative kernel lands
#include "nyxb_device.cuh"
extern "C" int nyxb_coop_supported(const DevSetup*, int) { return 0; }
extern "C" cudaError_t nyxb_launch_coop(const DevSetup*, int, size_t, const double*, const double*, const long long*,
                                        long long, long long*, double*, long long*, nyxb_details*, int*, cudaStream_t) {
    return cudaErrorNotSupported;
}
