// K1: multi-RHS sparse triangular solves  x = Pc U^-1 L^-1 Pr b  on an LU factorisation.
//
// Design (see DESIGN.md "K1"): the right-hand-side COLUMNS are independent, so the
// block is cut into column panels of KP columns and ONE CTA owns one panel for the whole
// solve (row permutation, L levels, U levels, column permutation) -- no inter-CTA
// synchronisation at all, only __syncthreads() between dependency levels.  The panel
// lives in shared memory when n*KP*8 bytes fit (the reference's cavity configs do),
// otherwise in a per-CTA global slab that stays L1/L2 resident.  Rows of L and U are
// stored sorted by dependency level so each level streams contiguous CSR (coalesced
// int32 column / FP64 value loads); a row is reduced by a group of G = 2^g lanes
// (g chosen per level on the host from the mean row length), every lane holding KP
// partial sums, followed by a shuffle reduction.
//
// Algorithmic bytes per solve: 12*(nnzL+nnzU) + 16*(n+1) + 32*n*k.
#include "common.cuh"
#include <vector>
#include <algorithm>

namespace ocb {

struct TriDev {
    const int32_t* lvl_ptr;  // nlev+1, positions in level-sorted row order
    const int32_t* rowid;    // n: original row of sorted position q
    const int32_t* rowptr;   // n+1 (sorted order)
    const int32_t* colidx;   // off-diagonal entries only
    const double* vals;
    const double* dinv;      // U only: 1/diag by sorted position
    const uint8_t* glog;     // nlev: log2 of lanes per row
    int nlev;
};

struct TriHost {
    int32_t *lvl_ptr = nullptr, *rowid = nullptr, *rowptr = nullptr, *colidx = nullptr;
    double *vals = nullptr, *dinv = nullptr;
    uint8_t* glog = nullptr;
    int nlev = 0;
    int64_t nnz = 0, maxwidth = 0;
    TriDev dev() const { return TriDev{lvl_ptr, rowid, rowptr, colidx, vals, dinv, glog, nlev}; }
};

}  // namespace ocb

struct ocb_lu {
    int64_t n = 0;
    ocb::TriHost L, U;
    int32_t *perm_r = nullptr, *perm_c = nullptr;
    int64_t bytes = 0;
    int max_smem_optin = 0;
};

namespace ocb {

constexpr int TRSM_THREADS = 512;

template <int KP, bool UPPER>
__device__ __forceinline__ void tri_levels(const TriDev F, double* x) {
    const int tid = threadIdx.x;
    const int l0 = UPPER ? 0 : 1;  // level 0 of unit-lower L has nothing to subtract
    if (F.nlev <= l0) return;
    int q0 = __ldg(F.lvl_ptr + l0), q1 = __ldg(F.lvl_ptr + l0 + 1);
    int gl = __ldg(F.glog + l0);
    for (int l = l0; l < F.nlev; ++l) {
        // prefetch the next level's descriptor; consumed after the barrier
        int nq1 = 0, ngl = 0;
        if (l + 1 < F.nlev) {
            nq1 = __ldg(F.lvl_ptr + l + 2);
            ngl = __ldg(F.glog + l + 1);
        }
        const int G = 1 << gl;
        const int gid = tid >> gl, glane = tid & (G - 1);
        const int ngroups = TRSM_THREADS >> gl;
        for (int qq = q0; qq < q1; qq += ngroups) {  // warp-uniform trip count
            const int q = qq + gid;
            const bool valid = q < q1;
            double acc[KP];
#pragma unroll
            for (int c = 0; c < KP; ++c) acc[c] = 0.0;
            int beg = 0, end = 0;
            if (valid) {
                beg = __ldg(F.rowptr + q);
                end = __ldg(F.rowptr + q + 1);
            }
            for (int p = beg + glane; p < end; p += G) {
                const int j = __ldg(F.colidx + p);
                const double v = __ldg(F.vals + p);
                const double* xj = x + (int64_t)j * KP;
#pragma unroll
                for (int c = 0; c < KP; ++c) acc[c] = fma(v, xj[c], acc[c]);
            }
            for (int o = G >> 1; o > 0; o >>= 1) {
#pragma unroll
                for (int c = 0; c < KP; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
            }
            if (valid && glane == 0) {
                double* xi = x + (int64_t)__ldg(F.rowid + q) * KP;
                if (UPPER) {
                    const double d = __ldg(F.dinv + q);
#pragma unroll
                    for (int c = 0; c < KP; ++c) xi[c] = (xi[c] - acc[c]) * d;
                } else {
#pragma unroll
                    for (int c = 0; c < KP; ++c) xi[c] -= acc[c];
                }
            }
        }
        __syncthreads();
        q0 = q1;
        q1 = nq1;
        gl = ngl;
    }
}

struct SolveArgs {
    TriDev L, U;
    const int32_t *perm_r, *perm_c;
    int64_t n;
    const double* B;
    int64_t ldb, nrows_b;
    double* X;
    int64_t ldx, nrows_x, k;
    double* ws;
};

template <int KP, bool SMEMX>
__global__ void __launch_bounds__(TRSM_THREADS, 1) sptrsm_panel_kernel(const SolveArgs a) {
    extern __shared__ double xs[];
    double* x = SMEMX ? xs : a.ws + (int64_t)blockIdx.x * a.n * KP;
    const int64_t c0 = (int64_t)blockIdx.x * KP;
    const int tid = threadIdx.x;
    // x[perm_r[i]] = b[i]  (Pr b), rows beyond nrows_b are zero
    for (int64_t e = tid; e < a.n * KP; e += TRSM_THREADS) {
        const int64_t i = e / KP;
        const int c = (int)(e - i * KP);
        double v = 0.0;
        if (i < a.nrows_b && c0 + c < a.k) v = a.B[i * a.ldb + c0 + c];
        x[(int64_t)__ldg(a.perm_r + i) * KP + c] = v;
    }
    __syncthreads();
    tri_levels<KP, false>(a.L, x);
    tri_levels<KP, true>(a.U, x);
    // out[j] = z[perm_c[j]]
    for (int64_t e = tid; e < a.nrows_x * KP; e += TRSM_THREADS) {
        const int64_t j = e / KP;
        const int c = (int)(e - j * KP);
        if (c0 + c < a.k) a.X[j * a.ldx + c0 + c] = x[(int64_t)__ldg(a.perm_c + j) * KP + c];
    }
}

// ---- host-side analysis ---------------------------------------------------------

static int build_factor(int64_t n, const int32_t* rp, const int32_t* ci, const double* va,
                        bool upper, TriHost* out, int64_t* bytes, cudaStream_t st) {
    std::vector<int32_t> level(n, 0);
    int nlev = 0;
    int64_t nnz_off = 0;
    std::vector<double> diag(upper ? n : 0, 0.0);
    if (!upper) {
        for (int64_t i = 0; i < n; ++i) {
            int lv = 0;
            for (int32_t p = rp[i]; p < rp[i + 1]; ++p) {
                const int32_t j = ci[p];
                if (j < i) {
                    lv = std::max(lv, level[j] + 1);
                    ++nnz_off;
                } else if (j > i) {
                    set_error("L has an entry above the diagonal (row %lld col %d)", (long long)i, j);
                    return OCB_ERR_ARG;
                }
            }
            level[i] = lv;
            nlev = std::max(nlev, lv + 1);
        }
    } else {
        for (int64_t i = n - 1; i >= 0; --i) {
            int lv = 0;
            for (int32_t p = rp[i]; p < rp[i + 1]; ++p) {
                const int32_t j = ci[p];
                if (j > i) {
                    lv = std::max(lv, level[j] + 1);
                    ++nnz_off;
                } else if (j == i) {
                    diag[i] = va[p];
                } else {
                    set_error("U has an entry below the diagonal (row %lld col %d)", (long long)i, j);
                    return OCB_ERR_ARG;
                }
            }
            if (diag[i] == 0.0) {
                set_error("U has a zero pivot in row %lld", (long long)i);
                return OCB_ERR_SINGULAR;
            }
            level[i] = lv;
            nlev = std::max(nlev, lv + 1);
        }
    }
    if (n == 0) nlev = 0;
    // counting sort of rows by level
    std::vector<int32_t> lvl_ptr(nlev + 1, 0);
    for (int64_t i = 0; i < n; ++i) lvl_ptr[level[i] + 1]++;
    for (int l = 0; l < nlev; ++l) lvl_ptr[l + 1] += lvl_ptr[l];
    std::vector<int32_t> rowid(n), pos(lvl_ptr.begin(), lvl_ptr.end() - (nlev > 0 ? 1 : 0));
    if (nlev == 0) pos.clear();
    for (int64_t i = 0; i < n; ++i) rowid[pos[level[i]]++] = (int32_t)i;
    std::vector<int32_t> srp(n + 1, 0), sci(nnz_off);
    std::vector<double> sva(nnz_off), dinv(upper ? n : 0);
    int64_t w = 0;
    for (int64_t q = 0; q < n; ++q) {
        const int64_t i = rowid[q];
        srp[q] = (int32_t)w;
        for (int32_t p = rp[i]; p < rp[i + 1]; ++p) {
            if (ci[p] != i) {
                sci[w] = ci[p];
                sva[w] = va[p];
                ++w;
            }
        }
        if (upper) dinv[q] = 1.0 / diag[i];
    }
    srp[n] = (int32_t)w;
    std::vector<uint8_t> glog(std::max(nlev, 1), 0);
    int64_t maxwidth = 0;
    for (int l = 0; l < nlev; ++l) {
        const int64_t width = lvl_ptr[l + 1] - lvl_ptr[l];
        maxwidth = std::max(maxwidth, width);
        const int64_t ent = srp[lvl_ptr[l + 1]] - srp[lvl_ptr[l]];
        const double avg = width > 0 ? (double)ent / (double)width : 0.0;
        int g = 0;
        while (g < 5 && (double)(8 << g) < avg) ++g;  // about 8 entries per lane
        // narrow level: spare lanes are free, use them
        while (g < 5 && width * (int64_t)(2 << g) <= TRSM_THREADS && (double)(2 << g) <= avg) ++g;
        glog[l] = (uint8_t)g;
    }
    out->nlev = nlev;
    out->nnz = nnz_off;
    out->maxwidth = maxwidth;
#define OCB_UP(dst, vec, T)                                                               \
    do {                                                                                  \
        size_t b__ = std::max<size_t>((vec).size(), 1) * sizeof(T);                       \
        OCB_CUDA(cudaMalloc((void**)&(dst), b__));                                        \
        *bytes += (int64_t)b__;                                                           \
        if (!(vec).empty())                                                               \
            OCB_CUDA(cudaMemcpyAsync((dst), (vec).data(), (vec).size() * sizeof(T),       \
                                     cudaMemcpyHostToDevice, st));                        \
    } while (0)
    OCB_UP(out->lvl_ptr, lvl_ptr, int32_t);
    OCB_UP(out->rowid, rowid, int32_t);
    OCB_UP(out->rowptr, srp, int32_t);
    OCB_UP(out->colidx, sci, int32_t);
    OCB_UP(out->vals, sva, double);
    OCB_UP(out->glog, glog, uint8_t);
    if (upper) OCB_UP(out->dinv, dinv, double);
    // the vectors die at scope exit: the copies must have been consumed
    OCB_CUDA(cudaStreamSynchronize(st));
    return OCB_OK;
}

static void free_factor(TriHost* f) {
    cudaFree(f->lvl_ptr);
    cudaFree(f->rowid);
    cudaFree(f->rowptr);
    cudaFree(f->colidx);
    cudaFree(f->vals);
    cudaFree(f->dinv);
    cudaFree(f->glog);
    *f = TriHost();
}

// panel width / placement policy
static void solve_policy(const ocb_lu* lu, int* kp, bool* smem) {
    const int64_t cap = lu->max_smem_optin - 1024;
    if (lu->n * 4 * 8 <= cap) {
        *kp = 4;
        *smem = true;
    } else if (lu->n * 2 * 8 <= cap) {
        *kp = 2;
        *smem = true;
    } else {
        *kp = 8;
        *smem = false;
    }
}

template <int KP, bool SMEMX>
static int launch_solve(const ocb_lu* lu, const SolveArgs& a, cudaStream_t st) {
    const size_t smem = SMEMX ? (size_t)lu->n * KP * sizeof(double) : 0;
    if (SMEMX)
        OCB_CUDA(cudaFuncSetAttribute(sptrsm_panel_kernel<KP, SMEMX>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((a.k + KP - 1) / KP);
    sptrsm_panel_kernel<KP, SMEMX><<<grid, TRSM_THREADS, smem, st>>>(a);
    OCB_LAUNCH_CHECK();
    return OCB_OK;
}

int lu_solve_impl(const ocb_lu* lu, const double* B, int64_t ldb, int64_t nrows_b, double* X,
                  int64_t ldx, int64_t nrows_x, int64_t k, void* ws, int64_t ws_bytes,
                  cudaStream_t st) {
    if (k == 0 || lu->n == 0) return OCB_OK;
    int kp;
    bool smem;
    solve_policy(lu, &kp, &smem);
    SolveArgs a;
    a.L = lu->L.dev();
    a.U = lu->U.dev();
    a.perm_r = lu->perm_r;
    a.perm_c = lu->perm_c;
    a.n = lu->n;
    a.B = B;
    a.ldb = ldb;
    a.nrows_b = nrows_b;
    a.X = X;
    a.ldx = ldx;
    a.nrows_x = nrows_x;
    a.k = k;
    a.ws = (double*)ws;
    if (!smem) {
        const int64_t need = ocb_lu_solve_ws_bytes(lu, k);
        if (ws == nullptr || ws_bytes < need) {
            set_error("lu_solve: workspace too small (%lld < %lld)", (long long)ws_bytes,
                      (long long)need);
            return OCB_ERR_CAPACITY;
        }
        return launch_solve<8, false>(lu, a, st);
    }
    if (kp == 4) return launch_solve<4, true>(lu, a, st);
    return launch_solve<2, true>(lu, a, st);
}

}  // namespace ocb

extern "C" {

int ocb_lu_create(ocb_lu** out, int64_t n, const int32_t* h_L_rowptr, const int32_t* h_L_colidx,
                  const double* h_L_vals, const int32_t* h_U_rowptr, const int32_t* h_U_colidx,
                  const double* h_U_vals, const int32_t* h_perm_r, const int32_t* h_perm_c,
                  void* stream) {
    OCB_ARG(out && n >= 0, "lu_create");
    OCB_ARG(h_L_rowptr && h_U_rowptr && h_perm_r && h_perm_c, "lu_create: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    ocb_lu* lu = new ocb_lu();
    lu->n = n;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&lu->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    int rc = ocb::build_factor(n, h_L_rowptr, h_L_colidx, h_L_vals, false, &lu->L, &lu->bytes, st);
    if (rc == OCB_OK)
        rc = ocb::build_factor(n, h_U_rowptr, h_U_colidx, h_U_vals, true, &lu->U, &lu->bytes, st);
    if (rc == OCB_OK) {
        const size_t pb = std::max<int64_t>(n, 1) * sizeof(int32_t);
        cudaError_t e = cudaMalloc((void**)&lu->perm_r, pb);
        if (e == cudaSuccess) e = cudaMalloc((void**)&lu->perm_c, pb);
        if (e == cudaSuccess && n > 0)
            e = cudaMemcpyAsync(lu->perm_r, h_perm_r, pb, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess && n > 0)
            e = cudaMemcpyAsync(lu->perm_c, h_perm_c, pb, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            ocb::set_error("lu_create: %s", cudaGetErrorString(e));
            rc = OCB_ERR_CUDA;
        }
        lu->bytes += 2 * (int64_t)pb;
    }
    if (rc != OCB_OK) {
        ocb_lu_destroy(lu);
        return rc;
    }
    *out = lu;
    return OCB_OK;
}

int ocb_lu_destroy(ocb_lu* lu) {
    if (!lu) return OCB_OK;
    ocb::free_factor(&lu->L);
    ocb::free_factor(&lu->U);
    cudaFree(lu->perm_r);
    cudaFree(lu->perm_c);
    delete lu;
    return OCB_OK;
}

int ocb_lu_info(const ocb_lu* lu, int64_t* info8) {
    OCB_ARG(lu && info8, "lu_info");
    info8[0] = lu->n;
    info8[1] = lu->L.nnz;
    info8[2] = lu->U.nnz + lu->n;
    info8[3] = lu->L.nlev;
    info8[4] = lu->U.nlev;
    info8[5] = lu->bytes;
    info8[6] = lu->L.maxwidth;
    info8[7] = lu->U.maxwidth;
    return OCB_OK;
}

int64_t ocb_lu_solve_ws_bytes(const ocb_lu* lu, int64_t k) {
    if (!lu) return 0;
    int kp;
    bool smem;
    ocb::solve_policy(lu, &kp, &smem);
    if (smem) return 0;
    return ((k + kp - 1) / kp) * lu->n * kp * (int64_t)sizeof(double);
}

int ocb_lu_solve(const ocb_lu* lu, const double* d_B, int64_t ldb, int64_t nrows_b, double* d_X,
                 int64_t ldx, int64_t nrows_x, int64_t k, void* d_ws, int64_t ws_bytes,
                 void* stream) {
    OCB_ARG(lu, "lu_solve: null handle");
    OCB_ARG(k >= 0 && nrows_b >= 0 && nrows_b <= lu->n && nrows_x >= 0 && nrows_x <= lu->n,
            "lu_solve: sizes");
    OCB_ARG(ldb >= k && ldx >= k, "lu_solve: leading dimension < k");
    OCB_ARG(k == 0 || (d_B && d_X), "lu_solve: null pointer");
    return ocb::lu_solve_impl(lu, d_B, ldb, nrows_b, d_X, ldx, nrows_x, k, d_ws, ws_bytes,
                              (cudaStream_t)stream);
}
}
