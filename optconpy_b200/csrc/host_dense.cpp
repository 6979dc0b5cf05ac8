// Dense host kernels of the program builder, compiled by g++ (not nvcc) so that the hot
// loops can be cloned per instruction set and dispatched at load time: the same .so runs on
// any x86-64 host and uses AVX2/AVX-512 + FMA where the CPU has them.
#include "lu_program.h"

namespace ocb {

// X = inverse of the w x w triangular matrix D (both row-major, X zero on entry), row-oriented
// so that the inner loops are unit-stride:
//   lower:  X[i,:] = (e_i - sum_{k<i} D[i,k] X[k,:]) / D[i,i]
//   upper:  X[i,:] = (e_i - sum_{k>i} D[i,k] X[k,:]) / D[i,i]          (D[i,i] = 1 if unit)
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target_clones("avx512f", "avx2,fma", "default")))
#endif
void tri_inverse(const double* __restrict__ D, double* __restrict__ X, int w, bool upper, bool unit) {
    if (!upper) {
        for (int i = 0; i < w; ++i) {
            double* __restrict__ xi = X + (size_t)i * w;
            xi[i] = 1.0;
            for (int k = 0; k < i; ++k) {
                const double d = D[(size_t)i * w + k];
                if (d == 0.0) continue;
                const double* __restrict__ xk = X + (size_t)k * w;
                for (int j = 0; j <= k; ++j) xi[j] -= d * xk[j];
            }
            if (!unit) {
                const double inv = 1.0 / D[(size_t)i * w + i];
                for (int j = 0; j <= i; ++j) xi[j] *= inv;
            }
        }
    } else {
        for (int i = w - 1; i >= 0; --i) {
            double* __restrict__ xi = X + (size_t)i * w;
            xi[i] = 1.0;
            for (int k = i + 1; k < w; ++k) {
                const double d = D[(size_t)i * w + k];
                if (d == 0.0) continue;
                const double* __restrict__ xk = X + (size_t)k * w;
                for (int j = k; j < w; ++j) xi[j] -= d * xk[j];
            }
            if (!unit) {
                const double inv = 1.0 / D[(size_t)i * w + i];
                for (int j = i; j < w; ++j) xi[j] *= inv;
            }
        }
    }
}

// P = X * T, X w x w triangular, T and P w x m row-major (P zero on entry)
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target_clones("avx512f", "avx2,fma", "default")))
#endif
void tri_times_dense(const double* __restrict__ X, const double* __restrict__ T, double* __restrict__ P,
                     int w, int m, bool upper) {
    for (int k = 0; k < w; ++k) {
        double* __restrict__ pk = P + (size_t)k * m;
        const int s0 = upper ? k : 0, s1 = upper ? w : k + 1;
        for (int s = s0; s < s1; ++s) {
            const double x = X[(size_t)k * w + s];
            if (x == 0.0) continue;
            const double* __restrict__ ts = T + (size_t)s * m;
            for (int j = 0; j < m; ++j) pk[j] += x * ts[j];
        }
    }
}

}  // namespace ocb
