// Shared helpers for the optconpy_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/optconpy_b200.h"

namespace ocb {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define OCB_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            ocb::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call,               \
                           cudaGetErrorString(e__));                                \
            return OCB_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define OCB_LAUNCH_CHECK()                                                          \
    do {                                                                            \
        ocb::count_launch();                                                        \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            ocb::set_error("%s:%d launch: %s", __FILE__, __LINE__,                  \
                           cudaGetErrorString(e__));                                \
            return OCB_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define OCB_ARG(cond, msg)                                                          \
    do {                                                                            \
        if (!(cond)) {                                                              \
            ocb::set_error("%s:%d bad argument: %s", __FILE__, __LINE__, msg);      \
            return OCB_ERR_ARG;                                                     \
        }                                                                           \
    } while (0)

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// carve aligned sub-buffers out of one workspace
struct WsCarver {
    char* base;
    int64_t off, cap;
    WsCarver(void* p, int64_t bytes) : base((char*)p), off(0), cap(bytes) {}
    template <typename T>
    T* take(int64_t count) {
        off = align_up(off, 256);
        T* r = (T*)(base + off);
        off += count * (int64_t)sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap; }
};

int sm_count();

// FP64 tensor-core MMA (DMMA.8x8x4): D(8x8) += A(8x4, row) * B(4x8, col)
//   a = A[lane/4][lane%4], b = B[lane%4][lane/4], c0,c1 = C[lane/4][2*(lane%4)+{0,1}]
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

}  // namespace ocb
