// Low-rank factor machinery of the hot path:
//   * compress_Zsvd  (rank-revealing pivoted Cholesky of Z^T Z  + Jacobi core + DMMA products)
//   * the LR-ADI loop with Sherman-Morrison-Woodbury corrections (K6/K7 fused update + norm)
//   * single SMW saddle-point solves, the feedback product  Mt (Z (Z^T tB)).
#include "common.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include <algorithm>

struct ocb_lu;

namespace ocb {
// from the other translation units
int spmm_launch(int64_t nrows, const int32_t* rp, const int32_t* ci, const double* va,
                const double* X, int64_t ldx, double* Y, int64_t ldy, int64_t k, double alpha,
                double beta, cudaStream_t st, const int* skip = nullptr);
int lu_solve_impl(const ocb_lu* lu, const double* B, int64_t ldb, int64_t nrows_b, double* X,
                  int64_t ldx, int64_t nrows_x, int64_t k, void* ws, int64_t ws_bytes,
                  cudaStream_t st, const int* skip = nullptr);
void lu_solve_prof_discard_last(int64_t n);
int gram_impl(const double* Z, int64_t ldz, int64_t ka, const double* W, int64_t ldw, int64_t kb,
              int64_t n, double* G, int64_t ldg, void* ws, int64_t ws_bytes, cudaStream_t st);
int tall_gemm_impl(const double* Z, int64_t ldz, int64_t n, int64_t k, const double* T, int64_t ldt,
                   int64_t kc, double* C, int64_t ldc, double alpha, double beta, cudaStream_t st);
int sym_eig_impl(double* G, int64_t ldg, int64_t k, double* lam, double* V, int64_t ldv,
                 int32_t* h_sweeps, cudaStream_t st);
int sym_eig_async(double* G, int64_t ldg, int64_t k, double* lam, double* V, int64_t ldv,
                  int* d_sweeps, cudaStream_t st);
int smw_core_inv(const double* C, int64_t ldc, int m, double* Sinv, int* flag, cudaStream_t st);

// optional hook that turns the local ||V_i||_F^2 of an ADI step into the global one when the
// right-hand-side columns are sharded over ranks (the caller all-reduces the scalar)
typedef void (*ocb_norm_hook_t)(double* v_nsq, void* ctx);
static thread_local ocb_norm_hook_t g_norm_hook = nullptr;
static thread_local void* g_norm_hook_ctx = nullptr;

// one side stream + a few events per process (SMW preparation overlaps the ADI chain)
struct SideStream {
    static constexpr int MAXEV = 34;
    cudaStream_t s = nullptr;
    cudaEvent_t ev[MAXEV];
};
static SideStream* side_stream() {
    static SideStream* p = nullptr;
    static bool failed = false;
    if (!p && !failed) {
        SideStream* q = new SideStream();
        bool ok = cudaStreamCreateWithFlags(&q->s, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; ok && i < SideStream::MAXEV; ++i)
            ok = cudaEventCreateWithFlags(&q->ev[i], cudaEventDisableTiming) == cudaSuccess;
        if (ok) p = q; else failed = true;
    }
    return p;
}

// pinned host scratch for the few scalars that steer host-side loops; one per host thread (the
// stepper's tail thread runs ocb_smw_solve while the main thread sits in ocb_adi_run)
static double* pinned_scratch() {
    static thread_local double* p = nullptr;
    if (!p) {
        if (cudaHostAlloc((void**)&p, 16384, cudaHostAllocDefault) != cudaSuccess) p = nullptr;
    }
    return p;
}

// =================================================================================
// compress_Zsvd
// =================================================================================
struct CholState {
    double d0max;  // largest initial column norm^2
    double dp;     // pivot value of the current step
    int piv;       // pivot column of the current step
    int done;      // 1 once the stop rule fired
    int rank;      // number of Cholesky columns produced
    int pad;
};

constexpr int CH_THREADS = 512, CH_WARPS = 16;

// sum over rows of Z[i][j] * Z[i][p] for 32 columns j per CTA (p < 0: Z[i][j]^2).
// 16 warps stride the rows; fixed-order shared-memory reduction => deterministic.
__device__ __forceinline__ double col_dot_block(const double* __restrict__ Z, int64_t ldz, int64_t i_begin,
                                                int64_t n, int64_t K, int64_t j, int p, double (*red)[33]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0.0;
    if (j < K) {
        int64_t i = i_begin + warp;
        // 4 independent row streams per warp for memory-level parallelism
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (; i + 3 * CH_WARPS < n; i += 4 * CH_WARPS) {
            const double* r0 = Z + i * ldz;
            const double* r1 = r0 + (int64_t)CH_WARPS * ldz;
            const double* r2 = r1 + (int64_t)CH_WARPS * ldz;
            const double* r3 = r2 + (int64_t)CH_WARPS * ldz;
            const double z0 = __ldg(r0 + j), z1 = __ldg(r1 + j), z2 = __ldg(r2 + j), z3 = __ldg(r3 + j);
            const double p0 = p < 0 ? z0 : __ldg(r0 + p), p1 = p < 0 ? z1 : __ldg(r1 + p);
            const double p2 = p < 0 ? z2 : __ldg(r2 + p), p3 = p < 0 ? z3 : __ldg(r3 + p);
            a0 = fma(z0, p0, a0); a1 = fma(z1, p1, a1); a2 = fma(z2, p2, a2); a3 = fma(z3, p3, a3);
        }
        for (; i < n; i += CH_WARPS) {
            const double z0 = __ldg(Z + i * ldz + j);
            const double p0 = p < 0 ? z0 : __ldg(Z + i * ldz + p);
            a0 = fma(z0, p0, a0);
        }
        acc = (a0 + a1) + (a2 + a3);
    }
    red[warp][lane] = acc;
    __syncthreads();
    double s = 0.0;
    if (warp == 0)
        for (int w = 0; w < CH_WARPS; ++w) s += red[w][lane];
    return s;  // valid in warp 0
}

__global__ void __launch_bounds__(CH_THREADS) chol_colsq_kernel(const double* __restrict__ Z, int64_t ldz,
                                                               int64_t n, int64_t K, double* __restrict__ d) {
    __shared__ double red[CH_WARPS][33];
    const int64_t j = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
    const double s = col_dot_block(Z, ldz, 0, n, K, j, -1, red);
    if (threadIdx.x < 32 && j < K) d[j] = s;
}

__global__ void __launch_bounds__(1024) chol_pick_kernel(const double* __restrict__ d, int64_t K, int t,
                                                        int rmax, double eta, CholState* st) {
    __shared__ double bv[32];
    __shared__ int bi[32];
    if (st->done) return;
    double best = -1.0;
    int arg = -1;
    for (int64_t j = threadIdx.x; j < K; j += blockDim.x) {
        const double v = d[j];
        if (v > best) { best = v; arg = (int)j; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ov > best || (ov == best && oi >= 0 && (arg < 0 || oi < arg))) { best = ov; arg = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = arg; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w)
            if (bv[w] > best || (bv[w] == best && bi[w] >= 0 && (arg < 0 || bi[w] < arg))) { best = bv[w]; arg = bi[w]; }
        if (t == 0) st->d0max = best;
        if (t >= rmax || arg < 0 || !(best > eta * st->d0max) || !(best > 0.0)) {
            st->done = 1;
            st->rank = t;
        } else {
            st->piv = arg;
            st->dp = best;
        }
    }
}

// column t of the Cholesky factor: Rt[j][t] = (z_j . z_p - sum_{s<t} Rt[p][s] Rt[j][s]) / sqrt(dp)
// Pass 1 (grid: column blocks x CH_SPLIT row chunks, so that all SMs stream Z): partial dot
// products z_j . z_p of one row chunk.  Pass 2: ordered sum of the chunks (deterministic),
// subtraction of the previous factor columns, diagonal downdate.
constexpr int CH_SPLIT = 4;

__global__ void __launch_bounds__(CH_THREADS) chol_col_kernel(const double* __restrict__ Z, int64_t ldz,
                                                             int64_t n, int64_t K,
                                                             double* __restrict__ gpart,
                                                             const CholState* __restrict__ st) {
    __shared__ double red[CH_WARPS][33];
    if (st->done) return;
    const int p = st->piv;
    const int64_t chunk = (n + CH_SPLIT - 1) / CH_SPLIT;
    const int64_t i0 = (int64_t)blockIdx.y * chunk, i1 = i0 + chunk < n ? i0 + chunk : n;
    const int64_t j = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
    const double g = col_dot_block(Z, ldz, i0, i1, K, j, p, red);  // contains a __syncthreads
    if (threadIdx.x < 32 && j < K) gpart[(int64_t)blockIdx.y * K + j] = g;
}

__global__ void __launch_bounds__(256) chol_col_finish_kernel(const double* __restrict__ gpart, int64_t K,
                                                             int t, double* __restrict__ Rt, int64_t ldr,
                                                             double* __restrict__ d,
                                                             const CholState* __restrict__ st) {
    __shared__ double rp[1024];
    if (st->done) return;
    const int p = st->piv;
    const double dp = st->dp;
    for (int s = threadIdx.x; s < t; s += blockDim.x) rp[s] = Rt[(int64_t)p * ldr + s];
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= K) return;
    double g = 0.0;
    for (int c = 0; c < CH_SPLIT; ++c) g += gpart[(int64_t)c * K + j];
    const double* rj = Rt + j * ldr;
    double s0 = 0.0, s1 = 0.0;
    int s = 0;
    for (; s + 1 < t; s += 2) { s0 = fma(rp[s], rj[s], s0); s1 = fma(rp[s + 1], rj[s + 1], s1); }
    if (s < t) s0 = fma(rp[s], rj[s], s0);
    const double row = (g - (s0 + s1)) / sqrt(dp);
    Rt[j * ldr + t] = row;
    d[j] = (j == p) ? -1.0 : d[j] - row * row;
}

__global__ void scale_cols_kernel(const double* __restrict__ U, int64_t ldu, int r, int kk,
                                  const double* __restrict__ lam, double* __restrict__ Us, int64_t lds) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r * kk) return;
    const int i = e / kk, j = e % kk;
    Us[(int64_t)i * lds + j] = U[(int64_t)i * ldu + j] / sqrt(lam[j]);
}

// ---------------------------------------------------------------------------------
// Gram route (K <= GRAM_K_MAX): G = Z^T Z once on the FP64 tensor pipe (DMMA, symmetric half),
// then the SAME pivoted Cholesky on the explicit K x K matrix in ONE cooperative kernel: one
// grid-wide barrier per pivot, no pass over Z per pivot and no host round trip in the loop.
//   step t:  p = argmax d (ties: lowest index);  stop if t == rmax or d_p <= eta * d_max(0)
//            Rt[j][t] = (G[j][p] - sum_{s<t} Rt[p][s] Rt[j][s]) / sqrt(d_p);  d[j] -= Rt[j][t]^2
// ---------------------------------------------------------------------------------
constexpr int64_t GRAM_K_MAX = 4096;
constexpr int CG_THREADS = 256, CG_MAXBLOCKS = 64;

__global__ void __launch_bounds__(CG_THREADS) chol_gram_kernel(const double* __restrict__ G, int64_t ldg,
                                                              int64_t K, int rmax, double eta,
                                                              double* __restrict__ Rt, int64_t ldr,
                                                              double* __restrict__ d,
                                                              double* __restrict__ pval, int* __restrict__ pidx,
                                                              CholState* __restrict__ st) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double rp[1024];
    __shared__ double bv[CG_THREADS / 32];
    __shared__ int bi[CG_THREADS / 32];
    __shared__ double sp_val;
    __shared__ int sp_idx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + tid, gsize = (int64_t)gridDim.x * blockDim.x;
    auto better = [](double v, int i, double bvv, int bii) {
        return v > bvv || (v == bvv && i >= 0 && (bii < 0 || i < bii));
    };
    // local arg-max of d over the columns of this CTA -> pval/pidx[blockIdx.x]
    auto publish = [&](double best, int arg) {
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
            if (better(ov, oi, best, arg)) { best = ov; arg = oi; }
        }
        if (lane == 0) { bv[warp] = best; bi[warp] = arg; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < CG_THREADS / 32; ++w)
                if (better(bv[w], bi[w], best, arg)) { best = bv[w]; arg = bi[w]; }
            pval[blockIdx.x] = best;
            pidx[blockIdx.x] = arg;
        }
    };
    {
        double best = -1.0;
        int arg = -1;
        for (int64_t j = gtid; j < K; j += gsize) {
            const double v = G[j * ldg + j];
            d[j] = v;
            if (better(v, (int)j, best, arg)) { best = v; arg = (int)j; }
        }
        publish(best, arg);
    }
    grid.sync();
    double d0max = 0.0;
    for (int t = 0;; ++t) {
        // every CTA reduces the partial arg-maxima (identical result everywhere)
        if (warp == 0) {
            double best = -1.0;
            int arg = -1;
            for (int b = lane; b < (int)gridDim.x; b += 32)
                if (better(pval[b], pidx[b], best, arg)) { best = pval[b]; arg = pidx[b]; }
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
                if (better(ov, oi, best, arg)) { best = ov; arg = oi; }
            }
            if (lane == 0) { sp_val = best; sp_idx = arg; }
        }
        __syncthreads();
        const double dp = sp_val;
        const int p = sp_idx;
        if (t == 0) d0max = dp;
        if (t >= rmax || p < 0 || !(dp > eta * d0max) || !(dp > 0.0)) {
            if (gtid == 0) { st->done = 1; st->rank = t; st->d0max = d0max; }
            return;   // uniform over the grid
        }
        for (int s2 = tid; s2 < t; s2 += blockDim.x) rp[s2] = Rt[(int64_t)p * ldr + s2];
        __syncthreads();
        const double inv = 1.0 / sqrt(dp);
        double best = -1.0;
        int arg = -1;
        for (int64_t j = gtid; j < K; j += gsize) {
            const double* rj = Rt + j * ldr;
            double s0 = 0.0, s1 = 0.0;
            int s2 = 0;
            for (; s2 + 1 < t; s2 += 2) { s0 = fma(rp[s2], rj[s2], s0); s1 = fma(rp[s2 + 1], rj[s2 + 1], s1); }
            if (s2 < t) s0 = fma(rp[s2], rj[s2], s0);
            const double row = (G[j * ldg + p] - (s0 + s1)) * inv;
            Rt[j * ldr + t] = row;
            const double dj = ((int)j == p) ? -1.0 : d[j] - row * row;
            d[j] = dj;
            if (better(dj, (int)j, best, arg)) { best = dj; arg = (int)j; }
        }
        __syncthreads();   // bv/bi/rp are reused
        publish(best, arg);
        grid.sync();
    }
}

struct CompressWs {
    double *d, *Rt, *S, *U, *lam, *Us, *T, *gws, *gpart, *G, *pval;
    int *pidx, *sweeps;
    CholState* st;
    int64_t gws_bytes;
};

static bool gram_route(int64_t K) {
    static const bool off = getenv("OCB_COMPRESS_IMPLICIT") != nullptr;
    return !off && K <= GRAM_K_MAX;
}

static int64_t compress_carve(void* ws, int64_t bytes, int64_t n, int64_t K, int64_t rmax, CompressWs* o) {
    WsCarver c(ws, bytes);
    CompressWs w;
    w.st = c.take<CholState>(1);
    w.sweeps = c.take<int>(4);
    w.pval = c.take<double>(CG_MAXBLOCKS);
    w.pidx = c.take<int>(CG_MAXBLOCKS);
    w.d = c.take<double>(K);
    w.gpart = c.take<double>(K * CH_SPLIT);
    w.Rt = c.take<double>(K * rmax);
    w.S = c.take<double>(rmax * rmax);
    w.U = c.take<double>(rmax * rmax);
    w.lam = c.take<double>(rmax);
    w.Us = c.take<double>(rmax * rmax);
    w.T = c.take<double>(K * rmax);
    w.gws_bytes = ocb_gram_ws_bytes(K, rmax, rmax);
    w.G = nullptr;
    if (gram_route(K)) {
        w.G = c.take<double>(K * K);
        w.gws_bytes = std::max(w.gws_bytes, ocb_gram_ws_bytes(n, K, K));
    }
    w.gws = (double*)c.take<char>(w.gws_bytes);
    if (o) *o = w;
    return c.off + 256;
}

}  // namespace ocb

extern "C" {

int64_t ocb_compress_ws_bytes(int64_t n, int64_t K, int64_t rmax) {
    return ocb::compress_carve(nullptr, 0, n, K, rmax, nullptr);
}

}  // extern "C"

namespace ocb {

// pivoted Cholesky of the explicit Gram matrix in one cooperative launch -> Rt, *rank (host)
static int chol_from_gram(const double* G, int64_t ldg, int64_t K, int64_t rmax, double eta, const CompressWs& w,
                          CholState* hst, cudaStream_t st) {
    int per_sm = 0;
    OCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chol_gram_kernel, CG_THREADS, 0));
    int64_t blocks = std::min<int64_t>((K + CG_THREADS - 1) / CG_THREADS * 4, CG_MAXBLOCKS);
    blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)std::max(per_sm, 1) * sm_count()));
    int64_t Kk = K, ldr = rmax, ldg2 = ldg;
    int rmx = (int)rmax;
    double* Rt = w.Rt; double* d = w.d; double* pval = w.pval; int* pidx = w.pidx; CholState* dst = w.st;
    void* args[] = {(void*)&G, &ldg2, &Kk, &rmx, &eta, &Rt, &ldr, &d, &pval, &pidx, &dst};
    OCB_CUDA(cudaLaunchCooperativeKernel((void*)chol_gram_kernel, dim3((unsigned)blocks), dim3(CG_THREADS),
                                         args, 0, st));
    count_launch();
    OCB_CUDA(cudaMemcpyAsync(hst, w.st, sizeof(CholState), cudaMemcpyDeviceToHost, st));
    OCB_CUDA(cudaStreamSynchronize(st));
    return OCB_OK;
}

// From the Cholesky factor Rt (K x r): core S = Rt^T Rt, its eigen-decomposition, the kept count
// and T (K x keep) = Rt U_keep Sigma_keep^-1  (the right singular vectors: Zc = Z T).
static int finish_from_rt(int64_t K, int r, int64_t rmax, double thresh, int64_t kmax, const CompressWs& w,
                          double* d_T, int64_t ldt, int64_t t_capacity_cols, double* d_sigma, int64_t* h_info3,
                          double* hp, cudaStream_t st) {
    int rc = gram_impl(w.Rt, rmax, r, w.Rt, rmax, r, K, w.S, rmax, w.gws, w.gws_bytes, st);
    if (rc) return rc;
    int32_t sweeps = 0;
    double* hlam = hp + 64;
    OCB_ARG(r <= 1900, "compress: rank above pinned scratch");
    rc = sym_eig_async(w.S, rmax, r, w.lam, w.U, rmax, w.sweeps, st);
    const bool small_eig = rc == OCB_OK;
    if (rc == -100) {
        rc = sym_eig_impl(w.S, rmax, r, w.lam, w.U, rmax, &sweeps, st);
        if (rc) return rc;
    } else if (rc) {
        return rc;
    } else {
        OCB_CUDA(cudaMemcpyAsync(hp + 32, w.sweeps, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    OCB_CUDA(cudaMemcpyAsync(hlam, w.lam, r * sizeof(double), cudaMemcpyDeviceToHost, st));
    OCB_CUDA(cudaStreamSynchronize(st));
    if (small_eig) sweeps = *(int*)(hp + 32);
    if (sweeps >= 40) { set_error("compress: Jacobi did not converge in 40 sweeps"); return OCB_ERR_NOCONV; }
    h_info3[2] = sweeps;
    int keep = 0;
    for (int i = 0; i < r; ++i) {
        const double sg = hlam[i] > 0.0 ? sqrt(hlam[i]) : 0.0;
        if (sg > 0.0 && (thresh < 0.0 || sg > thresh)) ++keep; else break;
    }
    if (kmax > 0) keep = (int)std::min<int64_t>(keep, kmax);
    h_info3[0] = keep;
    if (d_sigma) {
        for (int i = 0; i < r; ++i) hlam[i] = hlam[i] > 0.0 ? sqrt(hlam[i]) : 0.0;
        OCB_CUDA(cudaMemcpyAsync(d_sigma, hlam, r * sizeof(double), cudaMemcpyHostToDevice, st));
        OCB_CUDA(cudaStreamSynchronize(st));
    }
    if (keep == 0) return OCB_OK;
    if (keep > t_capacity_cols || ldt < keep) {
        set_error("compress: %d columns kept but capacity is %lld", keep, (long long)t_capacity_cols);
        return OCB_ERR_CAPACITY;
    }
    scale_cols_kernel<<<(r * keep + 255) / 256, 256, 0, st>>>(w.U, rmax, r, keep, w.lam, w.Us, rmax);
    OCB_LAUNCH_CHECK();
    return tall_gemm_impl(w.Rt, rmax, K, r, w.Us, rmax, keep, d_T, ldt, 1.0, 0.0, st);
}

}  // namespace ocb

extern "C" {

int64_t ocb_compress_gram_ws_bytes(int64_t K, int64_t rmax) {
    // the Gram matrix itself is the caller's: only the small pieces
    return ocb::compress_carve(nullptr, 0, 0, K + ocb::GRAM_K_MAX + 1, rmax, nullptr);
}

int ocb_compress_from_gram(const double* d_G, int64_t ldg, int64_t K, double thresh, int64_t kmax, double eta,
                           int64_t rmax, double* d_T, int64_t ldt, int64_t t_capacity_cols, double* d_sigma,
                           int64_t* h_info3, void* d_ws, int64_t ws_bytes, void* stream) {
    using namespace ocb;
    OCB_ARG(d_G && K >= 1 && ldg >= K && rmax >= 1 && rmax <= 1024 && d_T && h_info3 && d_ws, "compress_from_gram");
    OCB_ARG(eta > 0.0 && eta < 1e-6, "compress eta");
    cudaStream_t st = (cudaStream_t)stream;
    h_info3[0] = h_info3[1] = h_info3[2] = 0;
    rmax = std::min<int64_t>(rmax, K);
    CompressWs w;
    const int64_t need = compress_carve(d_ws, ws_bytes, 0, K + GRAM_K_MAX + 1, rmax, &w);   // no G inside
    if (need > ws_bytes) {
        set_error("compress_from_gram: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)need);
        return OCB_ERR_CAPACITY;
    }
    double* hp = pinned_scratch();
    OCB_ARG(hp != nullptr, "pinned scratch allocation failed");
    OCB_CUDA(cudaMemsetAsync(w.st, 0, sizeof(CholState), st));
    CholState* hst = (CholState*)hp;
    int rc = chol_from_gram(d_G, ldg, K, rmax, eta, w, hst, st);
    if (rc) return rc;
    const int r = hst->rank;
    h_info3[1] = r;
    if (r == 0) return OCB_OK;
    return finish_from_rt(K, r, rmax, thresh, kmax, w, d_T, ldt, t_capacity_cols, d_sigma, h_info3, hp, st);
}

int ocb_compress(const double* d_Z, int64_t ldz, int64_t n, int64_t K, double thresh, int64_t kmax,
                 double eta, int64_t rmax, double* d_Zc, int64_t ldzc, int64_t zc_capacity_cols,
                 double* d_sigma, int64_t* h_info3, void* d_ws, int64_t ws_bytes, void* stream) {
    using namespace ocb;
    OCB_ARG(n >= 0 && K >= 0 && ldz >= K && rmax >= 1 && rmax <= 1024, "compress sizes");
    OCB_ARG(d_Z && d_Zc && h_info3 && d_ws, "compress null");
    OCB_ARG(eta > 0.0 && eta < 1e-6, "compress eta");
    cudaStream_t st = (cudaStream_t)stream;
    h_info3[0] = h_info3[1] = h_info3[2] = 0;
    if (K == 0 || n == 0) return OCB_OK;
    rmax = std::min<int64_t>(rmax, std::min(K, n));
    CompressWs w;
    const int64_t need = compress_carve(d_ws, ws_bytes, n, K, rmax, &w);
    if (need > ws_bytes) {
        set_error("compress: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)need);
        return OCB_ERR_CAPACITY;
    }
    double* hp = pinned_scratch();
    OCB_ARG(hp != nullptr, "pinned scratch allocation failed");
    OCB_CUDA(cudaMemsetAsync(w.st, 0, sizeof(CholState), st));
    CholState* hst = (CholState*)hp;
    int rc;
    if (w.G) {
        // Gram route: one DMMA product, then the whole pivoted Cholesky in one cooperative launch
        rc = gram_impl(d_Z, ldz, K, d_Z, ldz, K, n, w.G, K, w.gws, w.gws_bytes, st);
        if (rc) return rc;
        rc = chol_from_gram(w.G, K, K, rmax, eta, w, hst, st);
        if (rc) return rc;
    } else {
        const unsigned cblocks = (unsigned)((K + 31) / 32);
        chol_colsq_kernel<<<cblocks, CH_THREADS, 0, st>>>(d_Z, ldz, n, K, w.d);
        OCB_LAUNCH_CHECK();
        int t = 0;
        bool done = false;
        while (!done) {
            const int batch_end = (int)std::min<int64_t>(rmax, t + 32);
            for (; t < batch_end; ++t) {
                chol_pick_kernel<<<1, 1024, 0, st>>>(w.d, K, t, (int)rmax, eta, w.st);
                OCB_LAUNCH_CHECK();
                chol_col_kernel<<<dim3(cblocks, CH_SPLIT), CH_THREADS, 0, st>>>(d_Z, ldz, n, K, w.gpart, w.st);
                OCB_LAUNCH_CHECK();
                chol_col_finish_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(w.gpart, K, t, w.Rt, rmax,
                                                                                    w.d, w.st);
                OCB_LAUNCH_CHECK();
            }
            if (t >= rmax) {  // closes the factorisation at rank rmax if the rule never fired
                chol_pick_kernel<<<1, 1024, 0, st>>>(w.d, K, t, (int)rmax, eta, w.st);
                OCB_LAUNCH_CHECK();
            }
            OCB_CUDA(cudaMemcpyAsync(hst, w.st, sizeof(CholState), cudaMemcpyDeviceToHost, st));
            OCB_CUDA(cudaStreamSynchronize(st));
            done = hst->done != 0;
        }
    }
    const int r = hst->rank;
    h_info3[1] = r;
    if (r == 0) return OCB_OK;
    // T (K x keep) = right singular vectors;  Zc = Z T
    rc = finish_from_rt(K, r, rmax, thresh, kmax, w, w.T, rmax, std::min<int64_t>(zc_capacity_cols, rmax),
                        d_sigma, h_info3, hp, st);
    if (rc) return rc;
    const int keep = (int)h_info3[0];
    if (keep == 0) return OCB_OK;
    if (ldzc < keep) {
        set_error("compress: %d columns kept but ldzc is %lld", keep, (long long)ldzc);
        return OCB_ERR_CAPACITY;
    }
    return tall_gemm_impl(d_Z, ldz, n, K, w.T, rmax, keep, d_Zc, ldzc, 1.0, 0.0, st);
}
}

// =================================================================================
// LR-ADI
// =================================================================================
namespace ocb {

// Device-resident state of one LR-ADI run (or one SMW solve): the stopping test
//     ||V_i||_F / ||[V_1..V_i]||_F <= reltol
// is evaluated ON THE DEVICE by the last CTA of the update kernel, in the same arithmetic and
// order the host loop used (z += v; rel = sqrt(v / z)), so that the iteration count is the
// oracle's.  Once `done` is set every later kernel of the loop returns immediately: the host
// enqueues iterations ahead without a round trip per step and reads the state only now and then.
struct AdiState {
    double z_nsq;           // running ||Z||_F^2
    double v_nsq;           // ||V_i||_F^2 of the last executed step
    int done;               // 1: converged, later launches are no-ops
    int steps;              // executed steps
    int singular;           // SMW core was singular (smw_core_inv_kernel)
    unsigned int counter;   // CTA arrival counter of the update kernel (per call: re-entrant)
};

// S2 (m x k) = Sinv (m x m) * small (m x k)
__global__ void smw_s2_kernel(const double* __restrict__ Sinv, int m, const double* __restrict__ small,
                              int64_t k, double* __restrict__ S2, const int* __restrict__ skip) {
    if (skip && *skip) return;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)m * k) return;
    const int r = (int)(e / k);
    const int64_t c = e % k;
    double s = 0.0;
    for (int l = 0; l < m; ++l) s = fma(Sinv[r * m + l], small[(int64_t)l * k + c], s);
    S2[e] = s;
}

// Vnew = a*Vold + b*(Y + AiU*S2);  ||Vnew||_F^2 (per-CTA partials, last CTA sums in order) and,
// with decide_step >= 0, the stopping test of ADI step decide_step
template <int MMAX>
__global__ void __launch_bounds__(256) adi_update_kernel(const double* __restrict__ Vold, int64_t ldv,
                                                        const double* __restrict__ Y, int64_t ldy,
                                                        const double* __restrict__ AiU, int64_t lda, int m,
                                                        const double* __restrict__ S2,
                                                        double* __restrict__ Vnew, int64_t ldn,
                                                        int64_t nrows, int64_t k, double a, double b,
                                                        double* __restrict__ partials,
                                                        AdiState* __restrict__ state,
                                                        const double* __restrict__ Sinv,
                                                        const double* __restrict__ small,
                                                        int decide_step, double reltol,
                                                        double* __restrict__ relnorms) {
    extern __shared__ double s2s[];   // fused variant: S2 = Sinv * small, recomputed per CTA
    __shared__ double red[8];
    __shared__ bool last;
    if (state->done) return;          // uniform over the grid (set by an earlier launch)
    if (MMAX > 0 && Sinv != nullptr) {
        // same formula and summation order as smw_s2_kernel: bit-identical, one launch less
        for (int64_t e = threadIdx.x; e < (int64_t)m * k; e += blockDim.x) {
            const int r = (int)(e / k);
            const int64_t c = e % k;
            double sv = 0.0;
            for (int l = 0; l < m; ++l) sv = fma(Sinv[r * m + l], small[(int64_t)l * k + c], sv);
            s2s[e] = sv;
        }
        __syncthreads();
        S2 = s2s;
    }
    const int64_t total = nrows * k;
    double acc = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / k, c = e - i * k;
        double x = Y[i * ldy + c];
        if (MMAX > 0) {
            const double* ai = AiU + i * lda;
            for (int l = 0; l < m; ++l) x = fma(ai[l], S2[(int64_t)l * k + c], x);
        }
        double v = b * x;
        if (Vold) v = fma(a, Vold[i * ldv + c], v);
        if (Vnew) Vnew[i * ldn + c] = v;
        acc = fma(v, v, acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        partials[blockIdx.x] = s;
        __threadfence();
        const unsigned int done = atomicAdd(&state->counter, 1u);
        last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double s = 0.0;
        for (unsigned int bI = 0; bI < gridDim.x; ++bI) s += ((volatile double*)partials)[bI];
        state->v_nsq = s;
        state->counter = 0;
        if (decide_step >= 0) {
            const double z = state->z_nsq + s;
            state->z_nsq = z;
            const double rel = z > 0.0 ? sqrt(s / z) : 0.0;
            relnorms[decide_step] = rel;
            state->steps = decide_step + 1;
            if (!(rel > reltol)) state->done = 1;
        }
    }
}

constexpr int UPD_MAXBLOCKS = 1024;

constexpr int64_t UPD_FUSE_MAX = 4096;   // m*k doubles of shared memory for the fused S2

// Sinv/small != null: S2 = Sinv*small is computed inside the update kernel (requires
// m*k <= UPD_FUSE_MAX); otherwise S2 must hold it already.
static int adi_update(const double* Vold, int64_t ldv, const double* Y, int64_t ldy, const double* AiU,
                      int64_t lda, int m, const double* S2, double* Vnew, int64_t ldn, int64_t nrows,
                      int64_t k, double a, double b, double* partials, AdiState* state, cudaStream_t st,
                      const double* Sinv = nullptr, const double* small = nullptr, int decide_step = -1,
                      double reltol = 0.0, double* relnorms = nullptr) {
    const int64_t total = nrows * k;
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(UPD_MAXBLOCKS, (total + 1023) / 1024));
    const size_t smem = (m > 0 && Sinv) ? (size_t)m * k * sizeof(double) : 0;
    if (m > 0)
        adi_update_kernel<1><<<blocks, 256, smem, st>>>(Vold, ldv, Y, ldy, AiU, lda, m, S2, Vnew, ldn, nrows, k, a, b, partials, state, Sinv, small, decide_step, reltol, relnorms);
    else
        adi_update_kernel<0><<<blocks, 256, 0, st>>>(Vold, ldv, Y, ldy, AiU, lda, 0, S2, Vnew, ldn, nrows, k, a, b, partials, state, nullptr, nullptr, decide_step, reltol, relnorms);
    OCB_LAUNCH_CHECK();
    return OCB_OK;
}

// prepare the SMW pieces of one factorisation: AiU = A^-1 [U;0] (first nrows rows), Sinv
static int smw_prepare(const ocb_lu* lu, int64_t NV, const double* Ufb, int64_t ldu, int m,
                       const int32_t* vt_rp, const int32_t* vt_ci, const double* vt_va,
                       double* AiU, int64_t nrows_aiu, double* core, double* Sinv, int* flag,
                       void* lws, int64_t lws_bytes, cudaStream_t st) {
    int rc = lu_solve_impl(lu, Ufb, ldu, NV, AiU, m, nrows_aiu, m, lws, lws_bytes, st);
    if (rc) return rc;
    // core = Vt (m x NV) * AiU[:NV]  (m x m)
    rc = spmm_launch(m, vt_rp, vt_ci, vt_va, AiU, m, core, m, m, 1.0, 0.0, st);
    if (rc) return rc;
    return smw_core_inv(core, m, m, Sinv, flag, st);
}

}  // namespace ocb

extern "C" {

int ocb_adi_set_norm_hook(void (*hook)(double*, void*), void* ctx) {
    ocb::g_norm_hook = hook;
    ocb::g_norm_hook_ctx = ctx;
    return OCB_OK;
}

constexpr int64_t ADI_MAXSTEPS_CAP = 8192;   // device array of relative norms

int64_t ocb_adi_ws_bytes(int64_t n_sad, int64_t k, int64_t m, int64_t nshifts, ocb_lu* const* lus) {
    using namespace ocb;
    int64_t lws = 0;
    for (int64_t i = 0; i < nshifts; ++i)
        lws = std::max(lws, ocb_lu_solve_ws_bytes(lus[i], std::max(k, m)));
    WsCarver c(nullptr, 0);
    c.take<double>(n_sad * k);                   // T = Mt V
    c.take<double>(n_sad * k);                   // Y
    c.take<double>(UPD_MAXBLOCKS);               // partials
    c.take<AdiState>(1);                         // device-resident loop state
    c.take<double>(ADI_MAXSTEPS_CAP);            // relative norms
    c.take<double>(nshifts * n_sad * std::max<int64_t>(m, 1));
    c.take<double>(nshifts * m * m + 1);
    c.take<double>(nshifts * m * m + 1);
    c.take<double>(m * k + 1);
    c.take<double>(m * k + 1);
    c.take<char>(lws);
    return c.off + 256;
}

// number of steps the previous run of this thread took: the first batch of iterations that is
// enqueued without looking at the result (consecutive Newton steps / time steps of the DRE take
// nearly the same number of ADI steps)
static thread_local int64_t g_adi_steps_hint = 0;

int ocb_adi_run(ocb_lu* const* lus, const double* h_shifts, int64_t nshifts, int64_t NV, int64_t NP,
                const int32_t* d_Mt_rowptr, const int32_t* d_Mt_colidx, const double* d_Mt_vals,
                const double* d_W, int64_t ldw, int64_t k, const double* d_Ufb, int64_t ldu, int64_t m,
                const int32_t* d_Vt_rowptr, const int32_t* d_Vt_colidx, const double* d_Vt_vals,
                int64_t maxsteps, double reltol, double* d_Z, int64_t ldz, int64_t z_capacity_cols,
                double* h_relnorms, int64_t* h_steps, void* d_ws, int64_t ws_bytes, void* stream) {
    using namespace ocb;
    OCB_ARG(lus && h_shifts && nshifts >= 1 && NV >= 1 && NP >= 0 && k >= 1, "adi sizes");
    OCB_ARG(d_Mt_rowptr && d_W && d_Z && h_relnorms && h_steps && d_ws, "adi null");
    OCB_ARG(m == 0 || (d_Ufb && d_Vt_rowptr && m <= 32 && ldu >= m), "adi low-rank part");
    OCB_ARG(maxsteps >= 1 && maxsteps <= ADI_MAXSTEPS_CAP && ldw >= k, "adi steps/ld");
    for (int64_t i = 0; i < nshifts; ++i) OCB_ARG(h_shifts[i] < 0.0, "adi shifts must be negative reals");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_sad = NV + NP;
    const int64_t need = ocb_adi_ws_bytes(n_sad, k, m, nshifts, lus);
    if (ws_bytes < need) {
        set_error("adi: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)need);
        return OCB_ERR_CAPACITY;
    }
    int64_t lws_bytes = 0;
    for (int64_t i = 0; i < nshifts; ++i)
        lws_bytes = std::max(lws_bytes, ocb_lu_solve_ws_bytes(lus[i], std::max(k, m)));
    WsCarver c(d_ws, ws_bytes);
    double* T = c.take<double>(n_sad * k);
    double* Y = c.take<double>(n_sad * k);
    double* partials = c.take<double>(UPD_MAXBLOCKS);
    AdiState* state = c.take<AdiState>(1);
    double* drel = c.take<double>(ADI_MAXSTEPS_CAP);
    double* AiU = c.take<double>(nshifts * n_sad * std::max<int64_t>(m, 1));
    double* Sinv = c.take<double>(nshifts * m * m + 1);
    double* core = c.take<double>(nshifts * m * m + 1);
    double* small = c.take<double>(m * k + 1);
    double* S2 = c.take<double>(m * k + 1);
    void* lws = c.take<char>(lws_bytes);
    double* hp = pinned_scratch();
    OCB_ARG(hp != nullptr, "pinned scratch allocation failed");
    AdiState* hstate = (AdiState*)hp;
    OCB_CUDA(cudaMemsetAsync(state, 0, sizeof(AdiState), st));
    int* flag = &state->singular;
    const int* skip = &state->done;
    std::vector<char> prepared(nshifts, 0);
    // The Sherman-Morrison-Woodbury pieces of the shifts (A_i^-1 U, 8 columns each) do not
    // depend on the iteration: they are all launched up front on a side stream, so that they
    // fill the SMs the narrow ADI solves leave idle; the main stream waits for shift i's
    // event right before its first use.
    SideStream* side = nullptr;
    static const bool no_side = getenv("OCB_NO_SIDE_STREAM") != nullptr;
    if (m > 0 && lws_bytes == 0 && nshifts <= SideStream::MAXEV - 1 && !no_side) {
        side = side_stream();
        if (side) {
            OCB_CUDA(cudaEventRecord(side->ev[SideStream::MAXEV - 1], st));
            OCB_CUDA(cudaStreamWaitEvent(side->s, side->ev[SideStream::MAXEV - 1], 0));
            for (int64_t i = 0; i < nshifts; ++i) {
                int rc = smw_prepare(lus[i], NV, d_Ufb, ldu, (int)m, d_Vt_rowptr, d_Vt_colidx, d_Vt_vals,
                                     AiU + i * NV * m, NV, core + i * m * m, Sinv + i * m * m, flag, lws,
                                     lws_bytes, side->s);
                if (rc) return rc;
                OCB_CUDA(cudaEventRecord(side->ev[i], side->s));
            }
        }
    }
    // error paths below must not leave side-stream work running on the caller's workspace
    auto fail = [&](int rc) {
        if (side) cudaStreamSynchronize(side->s);
        cudaStreamSynchronize(st);
        return rc;
    };

    // One ADI iteration = SpMM, multi-RHS solve, (SMW: small SpMM), fused update + norm + the
    // stopping test.  The test runs on the device; the host only enqueues.  It looks at the
    // state after a first batch sized by the previous run and then every few steps; the
    // iterations enqueued past the converged one are no-ops (their kernels see done = 1).
    const bool hooked = g_norm_hook != nullptr;   // column-sharded run: the host decides, step by step
    static const int chunk_env = getenv("OCB_ADI_CHUNK") ? atoi(getenv("OCB_ADI_CHUNK")) : 0;
    const int64_t chunk_next = chunk_env > 0 ? chunk_env : 4;
    int64_t chunk = hooked ? 1 : (chunk_env > 0 ? chunk_env : (g_adi_steps_hint > 1 ? g_adi_steps_hint : 8));
    double z_nsq = 0.0;
    int64_t step = 0;          // iterations enqueued so far
    int64_t steps_done = 0;    // iterations that really ran
    bool done = false;
    *h_steps = 0;
    while (!done && step < maxsteps) {
        const int64_t batch_end = std::min<int64_t>(maxsteps, step + chunk);
        for (; step < batch_end; ++step) {
            if ((step + 1) * k > z_capacity_cols || (step + 1) * k > ldz) {
                set_error("adi: Z capacity (%lld cols) exhausted at step %lld", (long long)z_capacity_cols,
                          (long long)step);
                return fail(OCB_ERR_CAPACITY);
            }
            const int64_t i = step % nshifts, ip = (step + nshifts - 1) % nshifts;
            const ocb_lu* lu = lus[i];
            int rc;
            if (m > 0 && !prepared[i]) {
                if (side) {
                    OCB_CUDA(cudaStreamWaitEvent(st, side->ev[i], 0));
                } else {
                    rc = smw_prepare(lu, NV, d_Ufb, ldu, (int)m, d_Vt_rowptr, d_Vt_colidx, d_Vt_vals,
                                     AiU + i * NV * m, NV, core + i * m * m, Sinv + i * m * m, flag, lws,
                                     lws_bytes, st);
                    if (rc) return fail(rc);
                }
                prepared[i] = 1;
            }
            const double* Vprev = step > 0 ? d_Z + (step - 1) * k : nullptr;
            double* Vnew = d_Z + step * k;
            if (step == 0) {
                rc = lu_solve_impl(lu, d_W, ldw, NV, Y, k, NV, k, lws, lws_bytes, st, skip);
            } else {
                rc = spmm_launch(NV, d_Mt_rowptr, d_Mt_colidx, d_Mt_vals, Vprev, ldz, T, k, k, 1.0, 0.0, st, skip);
                if (rc) return fail(rc);
                rc = lu_solve_impl(lu, T, k, NV, Y, k, NV, k, lws, lws_bytes, st, skip);
            }
            if (rc) return fail(rc);
            const bool fuse_s2 = m > 0 && m * k <= UPD_FUSE_MAX;
            if (m > 0) {
                rc = spmm_launch(m, d_Vt_rowptr, d_Vt_colidx, d_Vt_vals, Y, k, small, k, k, 1.0, 0.0, st, skip);
                if (rc) return fail(rc);
                if (!fuse_s2) {
                    smw_s2_kernel<<<(unsigned)((m * k + 255) / 256), 256, 0, st>>>(Sinv + i * m * m, (int)m, small, k, S2, skip);
                    OCB_LAUNCH_CHECK();
                }
            }
            double a, b;
            if (step == 0) { a = 0.0; b = sqrt(-2.0 * h_shifts[0]); }
            else {
                const double cs = sqrt(h_shifts[i] / h_shifts[ip]);
                a = cs;
                b = -cs * (h_shifts[i] + h_shifts[ip]);
            }
            rc = adi_update(Vprev, ldz, Y, k, AiU + i * NV * m, m, (int)m, S2, Vnew, ldz, NV, k, a, b,
                            partials, state, st, fuse_s2 ? Sinv + i * m * m : nullptr, fuse_s2 ? small : nullptr,
                            hooked ? -1 : (int)step, reltol, drel);
            if (rc) return fail(rc);
        }
        OCB_CUDA(cudaMemcpyAsync(hstate, state, sizeof(AdiState), cudaMemcpyDeviceToHost, st));
        OCB_CUDA(cudaStreamSynchronize(st));
        if (hstate->singular != 0) {
            set_error("adi: singular Sherman-Morrison-Woodbury core");
            return fail(OCB_ERR_SINGULAR);
        }
        if (hooked) {
            double v_nsq = hstate->v_nsq;
            g_norm_hook(&v_nsq, g_norm_hook_ctx);   // global ||V_i||_F^2 (the caller all-reduces)
            z_nsq += v_nsq;
            const double rel = z_nsq > 0.0 ? sqrt(v_nsq / z_nsq) : 0.0;
            h_relnorms[step - 1] = rel;
            steps_done = step;
            done = !(rel > reltol);
        } else {
            steps_done = hstate->steps;
            done = hstate->done != 0;
        }
        chunk = hooked ? 1 : chunk_next;
    }
    if (!hooked) {
        OCB_CUDA(cudaMemcpyAsync(h_relnorms, drel, (size_t)steps_done * sizeof(double), cudaMemcpyDeviceToHost, st));
        OCB_CUDA(cudaStreamSynchronize(st));
        lu_solve_prof_discard_last(step - steps_done);   // no-op launches leave the roofline statistics
        g_adi_steps_hint = steps_done;
    }
    *h_steps = steps_done;
    if (side)   // the workspace may be reused by the caller: order the main stream after all side work
        OCB_CUDA(cudaStreamWaitEvent(st, side->ev[nshifts - 1], 0));
    return OCB_OK;
}

int64_t ocb_smw_solve_ws_bytes(const ocb_lu* lu, int64_t k, int64_t m) {
    using namespace ocb;
    int64_t n = 0;
    if (lu) { int64_t info[8]; ocb_lu_info(lu, info); n = info[0]; }
    WsCarver c(nullptr, 0);
    c.take<double>(n * k);
    c.take<double>(n * std::max<int64_t>(m, 1));
    c.take<double>(m * m + 1);
    c.take<double>(m * m + 1);
    c.take<double>(m * k + 1);
    c.take<double>(m * k + 1);
    c.take<double>(UPD_MAXBLOCKS);
    c.take<AdiState>(1);
    c.take<char>(ocb_lu_solve_ws_bytes(lu, std::max(k, m)));
    return c.off + 256;
}

int ocb_smw_solve(const ocb_lu* lu, int64_t NV, const double* d_B, int64_t ldb, int64_t nrows_b, int64_t k,
                  const double* d_Ufb, int64_t ldu, int64_t m, const int32_t* d_Vt_rowptr,
                  const int32_t* d_Vt_colidx, const double* d_Vt_vals, double* d_X, int64_t ldx,
                  int64_t nrows_x, void* d_ws, int64_t ws_bytes, void* stream) {
    using namespace ocb;
    OCB_ARG(lu && d_B && d_X && k >= 1 && ldb >= k && ldx >= k, "smw_solve args");
    OCB_ARG(m == 0 || (d_Ufb && d_Vt_rowptr && m <= 32 && ldu >= m), "smw_solve low-rank part");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t info[8];
    ocb_lu_info(lu, info);
    const int64_t n = info[0];
    OCB_ARG(NV <= n && nrows_b <= n && nrows_x <= n, "smw_solve sizes");
    const int64_t need = ocb_smw_solve_ws_bytes(lu, k, m);
    if (!d_ws || ws_bytes < need) {
        set_error("smw_solve: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)need);
        return OCB_ERR_CAPACITY;
    }
    const int64_t lws_bytes = ocb_lu_solve_ws_bytes(lu, std::max(k, m));
    WsCarver c(d_ws, ws_bytes);
    double* Y = c.take<double>(n * k);
    double* AiU = c.take<double>(n * std::max<int64_t>(m, 1));
    double* core = c.take<double>(m * m + 1);
    double* Sinv = c.take<double>(m * m + 1);
    double* small = c.take<double>(m * k + 1);
    double* S2 = c.take<double>(m * k + 1);
    double* partials = c.take<double>(UPD_MAXBLOCKS);
    AdiState* state = c.take<AdiState>(1);
    void* lws = c.take<char>(lws_bytes);
    if (m == 0) return lu_solve_impl(lu, d_B, ldb, nrows_b, d_X, ldx, nrows_x, k, lws, lws_bytes, st);
    double* hp = pinned_scratch();
    OCB_ARG(hp != nullptr, "pinned scratch allocation failed");
    OCB_CUDA(cudaMemsetAsync(state, 0, sizeof(AdiState), st));
    int* flag = &state->singular;
    int rc = smw_prepare(lu, NV, d_Ufb, ldu, (int)m, d_Vt_rowptr, d_Vt_colidx, d_Vt_vals, AiU, n, core,
                         Sinv, flag, lws, lws_bytes, st);
    if (rc) return rc;
    rc = lu_solve_impl(lu, d_B, ldb, nrows_b, Y, k, n, k, lws, lws_bytes, st);
    if (rc) return rc;
    rc = spmm_launch(m, d_Vt_rowptr, d_Vt_colidx, d_Vt_vals, Y, k, small, k, k, 1.0, 0.0, st);
    if (rc) return rc;
    smw_s2_kernel<<<(unsigned)((m * k + 255) / 256), 256, 0, st>>>(Sinv, (int)m, small, k, S2, nullptr);
    OCB_LAUNCH_CHECK();
    rc = adi_update(nullptr, 0, Y, k, AiU, m, (int)m, S2, d_X, ldx, nrows_x, k, 0.0, 1.0, partials, state, st);
    if (rc) return rc;
    OCB_CUDA(cudaMemcpyAsync(hp, state, sizeof(AdiState), cudaMemcpyDeviceToHost, st));
    OCB_CUDA(cudaStreamSynchronize(st));
    if (((AdiState*)hp)->singular != 0) {
        set_error("smw_solve: singular Sherman-Morrison-Woodbury core");
        return OCB_ERR_SINGULAR;
    }
    return OCB_OK;
}

int64_t ocb_feedback_ws_bytes(int64_t NV, int64_t kz, int64_t m) {
    using namespace ocb;
    WsCarver c(nullptr, 0);
    c.take<double>(kz * m + 1);
    c.take<double>(NV * m + 1);
    c.take<char>(ocb_gram_ws_bytes(NV, kz, m));
    return c.off + 256;
}

int ocb_feedback(const int32_t* d_Mt_rowptr, const int32_t* d_Mt_colidx, const double* d_Mt_vals, int64_t NV,
                 const double* d_Z, int64_t ldz, int64_t kz, const double* d_tB, int64_t ldb, int64_t m,
                 double* d_Out, int64_t ldo, double alpha, void* d_ws, int64_t ws_bytes, void* stream) {
    using namespace ocb;
    OCB_ARG(d_Mt_rowptr && d_Z && d_tB && d_Out && NV >= 1 && kz >= 1 && m >= 1, "feedback args");
    OCB_ARG(ldz >= kz && ldb >= m && ldo >= m, "feedback ld");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t need = ocb_feedback_ws_bytes(NV, kz, m);
    if (!d_ws || ws_bytes < need) {
        set_error("feedback: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)need);
        return OCB_ERR_CAPACITY;
    }
    WsCarver c(d_ws, ws_bytes);
    double* ztb = c.take<double>(kz * m + 1);
    double* tmp = c.take<double>(NV * m + 1);
    const int64_t gws_bytes = ocb_gram_ws_bytes(NV, kz, m);
    void* gws = c.take<char>(gws_bytes);
    int rc = gram_impl(d_Z, ldz, kz, d_tB, ldb, m, NV, ztb, m, gws, gws_bytes, st);
    if (rc) return rc;
    rc = tall_gemm_impl(d_Z, ldz, NV, kz, ztb, m, m, tmp, m, 1.0, 0.0, st);
    if (rc) return rc;
    return spmm_launch(NV, d_Mt_rowptr, d_Mt_colidx, d_Mt_vals, tmp, m, d_Out, ldo, m, alpha, 0.0, st);
}
}
