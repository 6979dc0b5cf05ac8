// K2: CSR x dense block, Y = alpha*S*X + beta*Y (FP64, row-major blocks).
//
// HBM-bound: algorithmic bytes = 12*nnz + 4*(n+1) + 8*m*k + 8*n*k.
// One warp per row; the row's (col, val) pairs are read coalesced by the lanes and
// broadcast by shuffle; lanes own columns c = c0 + lane + 32*u, so every X row
// gather and every Y store is a contiguous 256-byte segment.
#include "common.cuh"

namespace ocb {

template <int CPT>
__global__ void __launch_bounds__(256) spmm_kernel(int64_t nrows, const int32_t* __restrict__ rowptr,
                                                   const int32_t* __restrict__ colidx,
                                                   const double* __restrict__ vals,
                                                   const double* __restrict__ X, int64_t ldx,
                                                   double* __restrict__ Y, int64_t ldy, int64_t k,
                                                   double alpha, double beta,
                                                   const int* __restrict__ skip) {
    if (skip && *skip) return;   // device-side stop flag of the ADI loop (lowrank.cu)
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int64_t c0 = (int64_t)blockIdx.y * (32 * CPT);
    const int beg = rowptr[row], end = rowptr[row + 1];
    double acc[CPT];
#pragma unroll
    for (int u = 0; u < CPT; ++u) acc[u] = 0.0;
    for (int base = beg; base < end; base += 32) {
        const int p = base + lane;
        int cj = 0;
        double v = 0.0;
        if (p < end) {
            cj = __ldg(colidx + p);
            v = __ldg(vals + p);
        }
        const int cnt = min(32, end - base);
        for (int t = 0; t < cnt; ++t) {
            const int j = __shfl_sync(0xffffffffu, cj, t);
            const double vv = __shfl_sync(0xffffffffu, v, t);
            const double* xr = X + (int64_t)j * ldx + c0 + lane;
#pragma unroll
            for (int u = 0; u < CPT; ++u)
                if (c0 + lane + 32 * u < k) acc[u] = fma(vv, __ldg(xr + 32 * u), acc[u]);
        }
    }
    double* yr = Y + row * ldy + c0 + lane;
#pragma unroll
    for (int u = 0; u < CPT; ++u) {
        if (c0 + lane + 32 * u < k) {
            double r = alpha * acc[u];
            if (beta != 0.0) r = fma(beta, yr[32 * u], r);
            yr[32 * u] = r;
        }
    }
}

int spmm_launch(int64_t nrows, const int32_t* rp, const int32_t* ci, const double* va,
                const double* X, int64_t ldx, double* Y, int64_t ldy, int64_t k, double alpha,
                double beta, cudaStream_t st, const int* skip) {
    if (nrows == 0 || k == 0) return OCB_OK;
    const int wpb = 8;
    dim3 block(32 * wpb);
    if (k <= 32) {
        dim3 grid((unsigned)((nrows + wpb - 1) / wpb), 1);
        spmm_kernel<1><<<grid, block, 0, st>>>(nrows, rp, ci, va, X, ldx, Y, ldy, k, alpha, beta, skip);
    } else if (k <= 64) {
        dim3 grid((unsigned)((nrows + wpb - 1) / wpb), 1);
        spmm_kernel<2><<<grid, block, 0, st>>>(nrows, rp, ci, va, X, ldx, Y, ldy, k, alpha, beta, skip);
    } else {
        dim3 grid((unsigned)((nrows + wpb - 1) / wpb), (unsigned)((k + 127) / 128));
        spmm_kernel<4><<<grid, block, 0, st>>>(nrows, rp, ci, va, X, ldx, Y, ldy, k, alpha, beta, skip);
    }
    OCB_LAUNCH_CHECK();
    return OCB_OK;
}
}  // namespace ocb

extern "C" int ocb_spmm(int64_t nrows, int64_t ncols, const int32_t* d_rowptr,
                        const int32_t* d_colidx, const double* d_vals, const double* d_X,
                        int64_t ldx, double* d_Y, int64_t ldy, int64_t k, double alpha, double beta,
                        void* stream) {
    OCB_ARG(nrows >= 0 && ncols >= 0 && k >= 0, "negative size");
    OCB_ARG(ldx >= k && ldy >= k, "leading dimension < k");
    OCB_ARG(nrows == 0 || k == 0 || (d_rowptr && d_X && d_Y), "null pointer");
    OCB_ARG(d_X != d_Y, "X and Y must not alias");
    return ocb::spmm_launch(nrows, d_rowptr, d_colidx, d_vals, d_X, ldx, d_Y, ldy, k, alpha, beta,
                            (cudaStream_t)stream, nullptr);
}
