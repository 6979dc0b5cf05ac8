// Error state, version, launch counter, small reductions.
#include "common.cuh"
#include <string.h>

namespace ocb {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// one CTA, fixed order: deterministic sum of squares
__global__ void __launch_bounds__(1024) sqnorm_kernel(const double* __restrict__ X, int64_t ldx,
                                                     int64_t n, int64_t k, double* out) {
    __shared__ double red[32];
    double acc = 0.0;
    int64_t total = n * k;
    for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
        int64_t i = e / k, c = e - i * k;
        double v = X[i * ldx + c];
        acc = fma(v, v, acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) *out = v;
    }
}
}  // namespace ocb

extern "C" {
const char* ocb_last_error(void) { return ocb::g_err; }
int ocb_version(void) { return 100; }
int64_t ocb_launch_count(void) { return (int64_t)ocb::g_launches.load(); }

int ocb_set_sync_mode(int blocking) {
    OCB_CUDA(cudaSetDeviceFlags(blocking ? cudaDeviceScheduleBlockingSync : cudaDeviceScheduleAuto));
    return OCB_OK;
}

int ocb_sqnorm(const double* d_X, int64_t ldx, int64_t n, int64_t k, double* d_out, void* stream) {
    OCB_ARG(d_X && d_out && n >= 0 && k >= 0 && ldx >= k, "sqnorm");
    ocb::sqnorm_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_X, ldx, n, k, d_out);
    OCB_LAUNCH_CHECK();
    return OCB_OK;
}
}
