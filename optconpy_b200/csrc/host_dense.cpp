// Dense host kernels of the program builder and of the numeric refactorisation, compiled by
// g++ (not nvcc) so that the hot loops can be cloned per instruction set and dispatched at load
// time: the same .so runs on any x86-64 host and uses AVX2/AVX-512 + FMA where the CPU has them.
#include <string.h>

#include <algorithm>
#include <vector>

#include "lu_program.h"

namespace ocb {

typedef double v8d __attribute__((vector_size(64), aligned(8)));

// C[M x N] += alpha * A[M x K] * B[K x N], all row-major with leading dimensions lda, ldb, ldc.
// 4 x 16 register tile: 8 vector accumulators, per k two loads of B and four broadcasts
// (~50 GFLOP/s per core with AVX-512 for K >= 32).  Every column chunk of a 4-row tile is loaded,
// updated in registers and stored once, so C may alias rows of B that the tile's coefficients
// do not touch (the in-place triangular products below rely on it); single leftover rows are
// updated in place, row by row.
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target_clones("avx512f", "avx2,fma", "default")))
#endif
void gemm_acc(int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* Cm, int ldc,
              double alpha) {
    int i = 0;
    for (; i + 4 <= M; i += 4) {
        const double* a0 = A + (size_t)i * lda;
        const double* a1 = a0 + lda;
        const double* a2 = a1 + lda;
        const double* a3 = a2 + lda;
        double* c0 = Cm + (size_t)i * ldc;
        double* c1 = c0 + ldc;
        double* c2 = c1 + ldc;
        double* c3 = c2 + ldc;
        int j = 0;
        for (; j + 16 <= N; j += 16) {
            v8d s00 = *(const v8d*)(c0 + j), s01 = *(const v8d*)(c0 + j + 8);
            v8d s10 = *(const v8d*)(c1 + j), s11 = *(const v8d*)(c1 + j + 8);
            v8d s20 = *(const v8d*)(c2 + j), s21 = *(const v8d*)(c2 + j + 8);
            v8d s30 = *(const v8d*)(c3 + j), s31 = *(const v8d*)(c3 + j + 8);
            for (int k = 0; k < K; ++k) {
                const double* b = B + (size_t)k * ldb + j;
                const v8d b0 = *(const v8d*)b, b1 = *(const v8d*)(b + 8);
                const double x0 = alpha * a0[k], x1 = alpha * a1[k], x2 = alpha * a2[k], x3 = alpha * a3[k];
                s00 += x0 * b0; s01 += x0 * b1;
                s10 += x1 * b0; s11 += x1 * b1;
                s20 += x2 * b0; s21 += x2 * b1;
                s30 += x3 * b0; s31 += x3 * b1;
            }
            *(v8d*)(c0 + j) = s00; *(v8d*)(c0 + j + 8) = s01;
            *(v8d*)(c1 + j) = s10; *(v8d*)(c1 + j + 8) = s11;
            *(v8d*)(c2 + j) = s20; *(v8d*)(c2 + j + 8) = s21;
            *(v8d*)(c3 + j) = s30; *(v8d*)(c3 + j + 8) = s31;
        }
        for (; j + 8 <= N; j += 8) {
            v8d s0 = *(const v8d*)(c0 + j), s1 = *(const v8d*)(c1 + j);
            v8d s2 = *(const v8d*)(c2 + j), s3 = *(const v8d*)(c3 + j);
            for (int k = 0; k < K; ++k) {
                const v8d b0 = *(const v8d*)(B + (size_t)k * ldb + j);
                s0 += (alpha * a0[k]) * b0; s1 += (alpha * a1[k]) * b0;
                s2 += (alpha * a2[k]) * b0; s3 += (alpha * a3[k]) * b0;
            }
            *(v8d*)(c0 + j) = s0; *(v8d*)(c1 + j) = s1; *(v8d*)(c2 + j) = s2; *(v8d*)(c3 + j) = s3;
        }
        if (j < N) {
            const int r = N - j;       // 1..7 columns left
            double t0[8] = {0}, t1[8] = {0}, t2[8] = {0}, t3[8] = {0};
            for (int k = 0; k < K; ++k) {
                const double* b = B + (size_t)k * ldb + j;
                const double x0 = alpha * a0[k], x1 = alpha * a1[k], x2 = alpha * a2[k], x3 = alpha * a3[k];
                for (int u = 0; u < r; ++u) {
                    t0[u] += x0 * b[u]; t1[u] += x1 * b[u]; t2[u] += x2 * b[u]; t3[u] += x3 * b[u];
                }
            }
            for (int u = 0; u < r; ++u) {
                c0[j + u] += t0[u]; c1[j + u] += t1[u]; c2[j + u] += t2[u]; c3[j + u] += t3[u];
            }
        }
    }
    for (; i < M; ++i) {
        const double* a = A + (size_t)i * lda;
        double* c = Cm + (size_t)i * ldc;
        for (int k = 0; k < K; ++k) {
            const double x = alpha * a[k];
            if (x == 0.0) continue;
            const double* b = B + (size_t)k * ldb;
            for (int j = 0; j < N; ++j) c[j] += x * b[j];
        }
    }
}

namespace {

// Row-oriented substitution for small blocks (unit-stride inner loops):
//   lower:  X[i,:] = (e_i - sum_{k<i} D[i,k] X[k,:]) / D[i,i]
//   upper:  X[i,:] = (e_i - sum_{k>i} D[i,k] X[k,:]) / D[i,i]          (D[i,i] = 1 if unit)
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target_clones("avx512f", "avx2,fma", "default")))
#endif
void tri_inverse_small(const double* __restrict__ D, int ldd, double* __restrict__ X, int ldx, int w,
                       bool upper, bool unit) {
    if (!upper) {
        for (int i = 0; i < w; ++i) {
            double* __restrict__ xi = X + (size_t)i * ldx;
            xi[i] = 1.0;
            for (int k = 0; k < i; ++k) {
                const double d = D[(size_t)i * ldd + k];
                if (d == 0.0) continue;
                const double* __restrict__ xk = X + (size_t)k * ldx;
                for (int j = 0; j <= k; ++j) xi[j] -= d * xk[j];
            }
            if (!unit) {
                const double inv = 1.0 / D[(size_t)i * ldd + i];
                for (int j = 0; j <= i; ++j) xi[j] *= inv;
            }
        }
    } else {
        for (int i = w - 1; i >= 0; --i) {
            double* __restrict__ xi = X + (size_t)i * ldx;
            xi[i] = 1.0;
            for (int k = i + 1; k < w; ++k) {
                const double d = D[(size_t)i * ldd + k];
                if (d == 0.0) continue;
                const double* __restrict__ xk = X + (size_t)k * ldx;
                for (int j = k; j < w; ++j) xi[j] -= d * xk[j];
            }
            if (!unit) {
                const double inv = 1.0 / D[(size_t)i * ldd + i];
                for (int j = i; j < w; ++j) xi[j] *= inv;
            }
        }
    }
}

// Recursive halving above 32 rows: with D = [D11 0; D21 D22],
//   inv(D) = [X11 0; -X22 D21 X11, X22]      (upper: [X11, -X11 D12 X22; 0, X22]),
// the two products through the GEMM kernel (twice the flops of substitution at ten times its
// speed).  X zero on entry; work: (w/2 + 2)^2 doubles.
void tri_inverse_rec(const double* D, int ldd, double* X, int ldx, int w, bool upper, bool unit, double* work) {
    if (w <= 32) {
        tri_inverse_small(D, ldd, X, ldx, w, upper, unit);
        return;
    }
    const int h = ((w / 2) + 3) & ~3, r = w - h;
    tri_inverse_rec(D, ldd, X, ldx, h, upper, unit, work);
    tri_inverse_rec(D + (size_t)h * ldd + h, ldd, X + (size_t)h * ldx + h, ldx, r, upper, unit, work);
    if (!upper) {
        memset(work, 0, (size_t)r * h * sizeof(double));                                       // T = -D21 X11
        gemm_acc(r, h, h, D + (size_t)h * ldd, ldd, X, ldx, work, h, -1.0);
        gemm_acc(r, h, r, X + (size_t)h * ldx + h, ldx, work, h, X + (size_t)h * ldx, ldx, 1.0);   // X21 = X22 T
    } else {
        memset(work, 0, (size_t)h * r * sizeof(double));                                       // T = -D12 X22
        gemm_acc(h, r, r, D + h, ldd, X + (size_t)h * ldx + h, ldx, work, r, -1.0);
        gemm_acc(h, r, h, X, ldx, work, r, X + h, ldx, 1.0);                                  // X12 = X11 T
    }
}

}  // namespace

// X = inverse of the w x w triangular matrix D (both row-major, X zero on entry)
void tri_inverse(const double* D, double* X, int w, bool upper, bool unit) {
    if (w <= 32) {
        tri_inverse_small(D, w, X, w, w, upper, unit);
        return;
    }
    static thread_local std::vector<double> work;
    const size_t need = (size_t)(w / 2 + 4) * (w / 2 + 4);
    if (work.size() < need) work.resize(need);
    tri_inverse_rec(D, w, X, w, w, upper, unit, work.data());
}

// P = X * T, X w x w triangular, T and P w x m row-major (P zero on entry): 4-row tiles of the
// GEMM kernel over the non-zero part of X's rows
void tri_times_dense(const double* X, const double* T, double* P, int w, int m, bool upper) {
    for (int i = 0; i < w; i += 4) {
        const int rows = std::min(4, w - i);
        if (!upper)
            gemm_acc(rows, m, i + rows, X + (size_t)i * w, w, T, m, P + (size_t)i * m, m, 1.0);
        else
            gemm_acc(rows, m, w - i, X + (size_t)i * w + i, w, T + (size_t)i * m, m, P + (size_t)i * m, m, 1.0);
    }
}

}  // namespace ocb
