// K1: multi-RHS sparse triangular solves  x = Pc U^-1 L^-1 Pr b  on an LU factorisation.
//
// Design (DESIGN.md "K1").  The right-hand-side COLUMNS are independent, so the block is
// cut into column panels of KP columns and ONE CTA owns one panel for the whole solve
// (row permutation, L levels, U levels, column permutation): no inter-CTA synchronisation,
// only __syncthreads() between dependency levels.  Rows of L and U are sorted by
// dependency level; a row is reduced by a group of G = 2^g lanes (g per level, chosen on
// the host from the mean row length), every lane holding KP partial sums, then a shuffle
// reduction.
//
// Two kernels:
//  * sptrsm_stream_kernel (the hot one for the reference's configs): the panel x lives in
//    shared memory and the factor is consumed from a packed, 16-byte aligned BATCH STREAM
//    that the TMA engine (cp.async.bulk + mbarrier complete_tx) copies into a shared-memory
//    ring a few batches ahead of the consumers.  The dependency chain of a level then only
//    sees shared-memory latency; the factor bytes arrive as large coalesced bulk copies.
//  * sptrsm_panel_kernel: generic fallback (panel in shared memory or in a per-CTA global
//    slab for large n), factor read with ordinary coalesced loads.
//
// Algorithmic bytes per solve: 12*(nnzL+nnzU) + 16*(n+1) + 32*n*k.
#include "common.cuh"
#include <vector>
#include <algorithm>
#include <string.h>
#include <stdlib.h>

namespace ocb {

struct TriDev {
    const int32_t* lvl_ptr;  // nlev+1, positions in level-sorted row order
    const int32_t* rowid;    // n: original row of sorted position q
    const int32_t* rowptr;   // n+1 (sorted order)
    const int32_t* colidx;   // off-diagonal entries only
    const double* vals;
    const double* dinv;      // 1/diag by sorted position (1.0 for the unit-lower factor)
    const uint8_t* glog;     // nlev: log2 of lanes per row
    int nlev;
};

struct TriSorted {  // host image of one level-sorted factor
    std::vector<int32_t> lvl_ptr, rowid, rowptr, colidx;
    std::vector<double> vals, dinv;
    std::vector<uint8_t> glog;       // lanes per row (<= 32) for the generic kernel
    std::vector<uint8_t> glog_wide;  // lanes per row (<= TRSM_THREADS) for the stream kernel
    int nlev = 0;
    int64_t maxwidth = 0;
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
    // packed batch stream for the TMA-fed kernel (L levels 1.., then U levels 0..)
    unsigned char* stream = nullptr;
    int64_t* batch_off = nullptr;  // nbatch+1 byte offsets into stream
    int nbatch = 0;
    int stage_bytes = 0;  // largest batch (multiple of 16)
    int kp_stream = 0;    // panel width of the stream kernel (0: stream kernel unavailable)
    int nstages = 0;
};

namespace ocb {

constexpr int TRSM_THREADS = 512;

// ---------------------------------------------------------------------------------
// generic kernel: factor from global memory
// ---------------------------------------------------------------------------------
template <int KP>
__device__ __forceinline__ void tri_levels(const TriDev F, double* x, int l0) {
    const int tid = threadIdx.x;
    if (F.nlev <= l0) return;
    int q0 = __ldg(F.lvl_ptr + l0), q1 = __ldg(F.lvl_ptr + l0 + 1);
    int gl = __ldg(F.glog + l0);
    for (int l = l0; l < F.nlev; ++l) {
        int nq1 = 0, ngl = 0;  // next level's descriptor, consumed after the barrier
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
            int beg = 0, end = 0, row = 0;
            double d = 1.0;
            if (valid) {
                beg = __ldg(F.rowptr + q);
                end = __ldg(F.rowptr + q + 1);
                row = __ldg(F.rowid + q);
                d = __ldg(F.dinv + q);
            }
#pragma unroll 4
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
                double* xi = x + (int64_t)row * KP;
#pragma unroll
                for (int c = 0; c < KP; ++c) xi[c] = (xi[c] - acc[c]) * d;
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
    // stream kernel
    const unsigned char* stream;
    const int64_t* batch_off;
    int nbatch, stage_bytes, nstages;
};

template <int KP>
__device__ __forceinline__ void load_panel(const SolveArgs& a, double* x, int64_t c0) {
    // x[perm_r[i]] = b[i]  (Pr b); rows beyond nrows_b are zero
    for (int64_t e = threadIdx.x; e < a.n * KP; e += blockDim.x) {
        const int64_t i = e / KP;
        const int c = (int)(e - i * KP);
        double v = 0.0;
        if (i < a.nrows_b && c0 + c < a.k) v = a.B[i * a.ldb + c0 + c];
        x[(int64_t)__ldg(a.perm_r + i) * KP + c] = v;
    }
}

template <int KP>
__device__ __forceinline__ void store_panel(const SolveArgs& a, const double* x, int64_t c0) {
    // out[j] = z[perm_c[j]]
    for (int64_t e = threadIdx.x; e < a.nrows_x * KP; e += blockDim.x) {
        const int64_t j = e / KP;
        const int c = (int)(e - j * KP);
        if (c0 + c < a.k) a.X[j * a.ldx + c0 + c] = x[(int64_t)__ldg(a.perm_c + j) * KP + c];
    }
}

template <int KP, bool SMEMX>
__global__ void __launch_bounds__(TRSM_THREADS, 1) sptrsm_panel_kernel(const SolveArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* x = SMEMX ? (double*)smem_raw : a.ws + (int64_t)blockIdx.x * a.n * KP;
    const int64_t c0 = (int64_t)blockIdx.x * KP;
    load_panel<KP>(a, x, c0);
    __syncthreads();
    tri_levels<KP>(a.L, x, 1);  // level 0 of the unit-lower factor has nothing to subtract
    tri_levels<KP>(a.U, x, 0);
    store_panel<KP>(a, x, c0);
}

// ---------------------------------------------------------------------------------
// stream kernel: factor arrives through a TMA-fed shared-memory ring
// ---------------------------------------------------------------------------------
// Batch record (every section 16-byte aligned, offsets in bytes from the record start):
//   int32 hdr[8] = {nlev, nrows, nent, off_lvl, off_rowbeg, off_rowid, off_dinv, off_col}
//   int32 off_val at hdr-extension [8], pad to 48 bytes
//   int32 lvl[nlev+1] (local row offsets), int32 glog[nlev]
//   int32 rowbeg[nrows+1] (local entry offsets), int32 rowid[nrows], f64 dinv[nrows]
//   int32 col[nent], f64 val[nent]
constexpr int HDR_INTS = 12;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

template <int KP>
__global__ void __launch_bounds__(TRSM_THREADS, 1) sptrsm_stream_kernel(const SolveArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [ring: nstages * stage_bytes][mbarriers: 8 * 8 bytes][x panel: n*KP doubles]
    unsigned char* ring = smem_raw;
    uint64_t* full = (uint64_t*)(smem_raw + (size_t)a.nstages * a.stage_bytes);
    double* red = (double*)(smem_raw + (size_t)a.nstages * a.stage_bytes + 64);  // 16 warps x KP
    double* x = red + (TRSM_THREADS / 32) * KP;
    const int tid = threadIdx.x;
    const int S = a.nstages;
    const int64_t c0 = (int64_t)blockIdx.x * KP;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {  // prologue: the first S-1 batches are in flight while the panel loads
        for (int b = 0; b < S - 1 && b < a.nbatch; ++b) {
            const int64_t o = __ldg(a.batch_off + b);
            const uint32_t bytes = (uint32_t)(__ldg(a.batch_off + b + 1) - o);
            mbar_expect_tx(full + b, bytes);
            bulk_g2s(ring + (size_t)b * a.stage_bytes, a.stream + o, bytes, full + b);
        }
    }
    load_panel<KP>(a, x, c0);
    __syncthreads();
    for (int b = 0; b < a.nbatch; ++b) {
        const int s = b % S;
        if (tid == 0) {
            const int nb = b + S - 1;  // its stage was released by the barrier ending batch b-1
            if (nb < a.nbatch) {
                const int ns = nb % S;
                const int64_t o = __ldg(a.batch_off + nb);
                const uint32_t bytes = (uint32_t)(__ldg(a.batch_off + nb + 1) - o);
                mbar_expect_tx(full + ns, bytes);
                bulk_g2s(ring + (size_t)ns * a.stage_bytes, a.stream + o, bytes, full + ns);
            }
        }
        mbar_wait(full + s, (uint32_t)((b / S) & 1));
        const unsigned char* rec = ring + (size_t)s * a.stage_bytes;
        const int32_t* hdr = (const int32_t*)rec;
        const int nlev = hdr[0];
        const int32_t* lvl = (const int32_t*)(rec + hdr[3]);
        const int32_t* glg = lvl + nlev + 1;
        const int32_t* rowbeg = (const int32_t*)(rec + hdr[4]);
        const int32_t* rowid = (const int32_t*)(rec + hdr[5]);
        const double* dinv = (const double*)(rec + hdr[6]);
        const int32_t* col = (const int32_t*)(rec + hdr[7]);
        const double* val = (const double*)(rec + hdr[8]);
        for (int l = 0; l < nlev; ++l) {
            const int q0 = lvl[l], q1 = lvl[l + 1];
            const int gl = glg[l];
            if (gl < 5) {
                // ---- G < 32 lanes per row: several rows per warp, plain butterflies.
                // Warps without a row in this level fall through to the barrier.
                const int G = 1 << gl;
                const int glane = tid & (G - 1);
                const int ngroups = TRSM_THREADS >> gl;
                for (int qb = q0 + ((tid & ~31) >> gl); qb < q1; qb += ngroups) {  // warp-uniform
                    const int q = qb + ((tid & 31) >> gl);
                    const bool valid = q < q1;
                    double acc[KP];
#pragma unroll
                    for (int c = 0; c < KP; ++c) acc[c] = 0.0;
                    int beg = 0, end = 0;
                    if (valid) {
                        beg = rowbeg[q];
                        end = rowbeg[q + 1];
                    }
#pragma unroll 4
                    for (int p = beg + glane; p < end; p += G) {
                        const int j = col[p];
                        const double v = val[p];
                        const double* xj = x + (size_t)j * KP;
#pragma unroll
                        for (int c = 0; c < KP; ++c) acc[c] = fma(v, xj[c], acc[c]);
                    }
                    for (int o = G >> 1; o > 0; o >>= 1) {
#pragma unroll
                        for (int c = 0; c < KP; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
                    }
                    if (valid && glane == 0) {
                        double* xi = x + (size_t)rowid[q] * KP;
                        const double d = dinv[q];
#pragma unroll
                        for (int c = 0; c < KP; ++c) xi[c] = (xi[c] - acc[c]) * d;
                    }
                }
                __syncthreads();
            } else {
                // ---- G >= 32: W = G/32 warps per row, (threads/G) rows per pass.
                const int W = 1 << (gl - 5);
                const int warp = tid >> 5, lane = tid & 31;
                const int rpp = (TRSM_THREADS / 32) >> (gl - 5);
                for (int qb = q0; qb < q1; qb += rpp) {  // block-uniform trip count
                const int q = qb + (warp >> (gl - 5));
                const bool valid = q < q1;  // warp-uniform
                double tot = 0.0;           // lane c*(32/KP) of the warp ends with column c
                if (valid) {
                    double acc[KP];
#pragma unroll
                    for (int c = 0; c < KP; ++c) acc[c] = 0.0;
                    const int beg = rowbeg[q], end = rowbeg[q + 1];
                    const int G = 32 * W;
#pragma unroll 4
                    for (int p = beg + (tid & (G - 1)); p < end; p += G) {
                        const int j = col[p];
                        const double v = val[p];
                        const double* xj = x + (size_t)j * KP;
#pragma unroll
                        for (int c = 0; c < KP; ++c) acc[c] = fma(v, xj[c], acc[c]);
                    }
                    // transposing butterfly: 6 (KP=4) instead of 20 double shuffles
                    if (KP == 4) {
                        const bool hi = lane & 16;
                        const double s0 = hi ? acc[0] : acc[2], s1 = hi ? acc[1] : acc[3 % KP];
                        const double k0 = hi ? acc[2] : acc[0], k1 = hi ? acc[3 % KP] : acc[1];
                        const double v0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
                        const double v1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
                        const bool hi2 = lane & 8;
                        tot = (hi2 ? v1 : v0) + __shfl_xor_sync(0xffffffffu, hi2 ? v0 : v1, 8);
                        tot += __shfl_xor_sync(0xffffffffu, tot, 4);
                        tot += __shfl_xor_sync(0xffffffffu, tot, 2);
                        tot += __shfl_xor_sync(0xffffffffu, tot, 1);
                    } else if (KP == 2) {
                        const bool hi = lane & 16;
                        tot = (hi ? acc[KP - 1] : acc[0]) +
                              __shfl_xor_sync(0xffffffffu, hi ? acc[0] : acc[KP - 1], 16);
                        tot += __shfl_xor_sync(0xffffffffu, tot, 8);
                        tot += __shfl_xor_sync(0xffffffffu, tot, 4);
                        tot += __shfl_xor_sync(0xffffffffu, tot, 2);
                        tot += __shfl_xor_sync(0xffffffffu, tot, 1);
                    } else {
                        tot = acc[0];
                        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                    }
                }
                const int cstep = 32 / KP;  // lane c*cstep holds column c
                if (W == 1) {
                    if (valid && (lane % cstep) == 0) {
                        double* xi = x + (size_t)rowid[q] * KP + lane / cstep;
                        *xi = (*xi - tot) * dinv[q];
                    }
                } else {
                    if (valid && (lane % cstep) == 0) red[warp * KP + lane / cstep] = tot;
                    __syncthreads();
                    if (valid && (warp & (W - 1)) == 0 && lane < KP) {
                        double sum = 0.0;
                        for (int w = 0; w < W; ++w) sum += red[(warp + w) * KP + lane];
                        double* xi = x + (size_t)rowid[q] * KP + lane;
                        *xi = (*xi - sum) * dinv[q];
                    }
                    if (qb + rpp < q1) __syncthreads();  // red[] is reused by the next pass
                }
                }
                __syncthreads();
            }
        }
    }
    store_panel<KP>(a, x, c0);
}

// ---------------------------------------------------------------------------------
// host-side analysis
// ---------------------------------------------------------------------------------
static int analyse_factor(int64_t n, const int32_t* rp, const int32_t* ci, const double* va,
                          bool upper, TriSorted* out) {
    std::vector<int32_t> level(n, 0);
    int nlev = 0;
    int64_t nnz_off = 0;
    std::vector<double> diag(n, 1.0);
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
            double dg = 0.0;
            for (int32_t p = rp[i]; p < rp[i + 1]; ++p) {
                const int32_t j = ci[p];
                if (j > i) {
                    lv = std::max(lv, level[j] + 1);
                    ++nnz_off;
                } else if (j == i) {
                    dg = va[p];
                } else {
                    set_error("U has an entry below the diagonal (row %lld col %d)", (long long)i, j);
                    return OCB_ERR_ARG;
                }
            }
            if (dg == 0.0) {
                set_error("U has a zero pivot in row %lld", (long long)i);
                return OCB_ERR_SINGULAR;
            }
            diag[i] = dg;
            level[i] = lv;
            nlev = std::max(nlev, lv + 1);
        }
    }
    if (n == 0) nlev = 0;
    TriSorted& t = *out;
    t.nlev = nlev;
    t.lvl_ptr.assign(nlev + 1, 0);
    for (int64_t i = 0; i < n; ++i) t.lvl_ptr[level[i] + 1]++;
    for (int l = 0; l < nlev; ++l) t.lvl_ptr[l + 1] += t.lvl_ptr[l];
    t.rowid.resize(n);
    {
        std::vector<int32_t> pos(t.lvl_ptr.begin(), t.lvl_ptr.begin() + nlev);
        for (int64_t i = 0; i < n; ++i) t.rowid[pos[level[i]]++] = (int32_t)i;
    }
    t.rowptr.assign(n + 1, 0);
    t.colidx.resize(nnz_off);
    t.vals.resize(nnz_off);
    t.dinv.resize(n);
    int64_t w = 0;
    for (int64_t q = 0; q < n; ++q) {
        const int64_t i = t.rowid[q];
        t.rowptr[q] = (int32_t)w;
        for (int32_t p = rp[i]; p < rp[i + 1]; ++p) {
            if (ci[p] != i) {
                t.colidx[w] = ci[p];
                t.vals[w] = va[p];
                ++w;
            }
        }
        t.dinv[q] = upper ? 1.0 / diag[i] : 1.0;
    }
    t.rowptr[n] = (int32_t)w;
    t.glog.assign(std::max(nlev, 1), 0);
    t.glog_wide.assign(std::max(nlev, 1), 0);
    t.maxwidth = 0;
    for (int l = 0; l < nlev; ++l) {
        const int64_t width = t.lvl_ptr[l + 1] - t.lvl_ptr[l];
        t.maxwidth = std::max(t.maxwidth, width);
        const int64_t ent = t.rowptr[t.lvl_ptr[l + 1]] - t.rowptr[t.lvl_ptr[l]];
        const double avg = width > 0 ? (double)ent / (double)width : 0.0;
        int g = 0;
        while (g < 5 && (double)(8 << g) < avg) ++g;  // about 8 entries per lane
        // narrow level: spare lanes are free, use them
        while (g < 5 && width * (int64_t)(2 << g) <= TRSM_THREADS && (double)(2 << g) <= avg) ++g;
        t.glog[l] = (uint8_t)g;
        // stream kernel: several warps may share one long row (cross-warp reduction in smem)
        int gw = g;
        while (gw < 9 && width * (int64_t)(2 << gw) <= TRSM_THREADS && (double)(4 << gw) <= avg) ++gw;
        t.glog_wide[l] = (uint8_t)gw;
    }
    return OCB_OK;
}

template <typename T>
static int upload(T** dst, const std::vector<T>& v, int64_t* bytes, cudaStream_t st) {
    const size_t b = std::max<size_t>(v.size(), 1) * sizeof(T);
    OCB_CUDA(cudaMalloc((void**)dst, b));
    *bytes += (int64_t)b;
    if (!v.empty()) OCB_CUDA(cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return OCB_OK;
}

static int upload_factor(const TriSorted& t, TriHost* out, int64_t* bytes, cudaStream_t st) {
    out->nlev = t.nlev;
    out->nnz = (int64_t)t.colidx.size();
    out->maxwidth = t.maxwidth;
    int rc;
    if ((rc = upload(&out->lvl_ptr, t.lvl_ptr, bytes, st))) return rc;
    if ((rc = upload(&out->rowid, t.rowid, bytes, st))) return rc;
    if ((rc = upload(&out->rowptr, t.rowptr, bytes, st))) return rc;
    if ((rc = upload(&out->colidx, t.colidx, bytes, st))) return rc;
    if ((rc = upload(&out->vals, t.vals, bytes, st))) return rc;
    if ((rc = upload(&out->dinv, t.dinv, bytes, st))) return rc;
    if ((rc = upload(&out->glog, t.glog, bytes, st))) return rc;
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

static inline int64_t a16(int64_t x) { return (x + 15) & ~(int64_t)15; }

static int64_t record_bytes(int64_t nlev, int64_t nrows, int64_t nent) {
    int64_t o = a16(HDR_INTS * 4);
    o = a16(o + (2 * nlev + 1) * 4);
    o = a16(o + (nrows + 1) * 4);
    o = a16(o + nrows * 4);
    o = a16(o + nrows * 8);
    o = a16(o + nent * 4);
    o = a16(o + nent * 8);
    return o;
}

struct BatchPlan {
    int factor;      // 0 = L, 1 = U
    int lev0, lev1;  // level range (single level if split by rows)
    int q0, q1;      // sorted-row range
};

// Greedy packing of whole levels into batches of at most cap bytes; a level that does not
// fit alone is split by rows.  Returns false if a single row exceeds the capacity.
static bool plan_batches(const TriSorted* f[2], int64_t cap, std::vector<BatchPlan>* plan) {
    plan->clear();
    for (int fi = 0; fi < 2; ++fi) {
        const TriSorted& t = *f[fi];
        int l = (fi == 0) ? 1 : 0;
        while (l < t.nlev) {
            // try to extend [l, l2)
            int l2 = l;
            while (l2 < t.nlev) {
                const int64_t rows = t.lvl_ptr[l2 + 1] - t.lvl_ptr[l];
                const int64_t ent = t.rowptr[t.lvl_ptr[l2 + 1]] - t.rowptr[t.lvl_ptr[l]];
                if (record_bytes(l2 + 1 - l, rows, ent) > cap) break;
                ++l2;
            }
            if (l2 > l) {
                plan->push_back(BatchPlan{fi, l, l2, t.lvl_ptr[l], t.lvl_ptr[l2]});
                l = l2;
                continue;
            }
            // level l alone is too large: split by rows
            int q = t.lvl_ptr[l];
            const int qend = t.lvl_ptr[l + 1];
            while (q < qend) {
                int q2 = q;
                while (q2 < qend && record_bytes(1, q2 + 1 - q, t.rowptr[q2 + 1] - t.rowptr[q]) <= cap) ++q2;
                if (q2 == q) return false;  // one row does not fit
                plan->push_back(BatchPlan{fi, l, l + 1, q, q2});
                q = q2;
            }
            ++l;
        }
    }
    return true;
}

static void write_record(const TriSorted& t, const BatchPlan& b, unsigned char* rec) {
    const int nlev = b.lev1 - b.lev0, nrows = b.q1 - b.q0;
    const int e0 = t.rowptr[b.q0], nent = t.rowptr[b.q1] - e0;
    int32_t* hdr = (int32_t*)rec;
    int64_t o = a16(HDR_INTS * 4);
    const int64_t off_lvl = o;
    o = a16(o + (2 * nlev + 1) * 4);
    const int64_t off_rowbeg = o;
    o = a16(o + (nrows + 1) * 4);
    const int64_t off_rowid = o;
    o = a16(o + nrows * 4);
    const int64_t off_dinv = o;
    o = a16(o + (int64_t)nrows * 8);
    const int64_t off_col = o;
    o = a16(o + (int64_t)nent * 4);
    const int64_t off_val = o;
    hdr[0] = nlev; hdr[1] = nrows; hdr[2] = nent;
    hdr[3] = (int32_t)off_lvl; hdr[4] = (int32_t)off_rowbeg; hdr[5] = (int32_t)off_rowid;
    hdr[6] = (int32_t)off_dinv; hdr[7] = (int32_t)off_col; hdr[8] = (int32_t)off_val;
    int32_t* lvl = (int32_t*)(rec + off_lvl);
    int32_t* glg = lvl + nlev + 1;
    if (nlev == 1) {  // possibly a row slice of one level
        lvl[0] = 0; lvl[1] = nrows; glg[0] = t.glog_wide[b.lev0];
    } else {
        for (int l = 0; l <= nlev; ++l) lvl[l] = t.lvl_ptr[b.lev0 + l] - b.q0;
        for (int l = 0; l < nlev; ++l) glg[l] = t.glog_wide[b.lev0 + l];
    }
    int32_t* rowbeg = (int32_t*)(rec + off_rowbeg);
    for (int q = 0; q <= nrows; ++q) rowbeg[q] = t.rowptr[b.q0 + q] - e0;
    memcpy(rec + off_rowid, t.rowid.data() + b.q0, (size_t)nrows * 4);
    memcpy(rec + off_dinv, t.dinv.data() + b.q0, (size_t)nrows * 8);
    memcpy(rec + off_col, t.colidx.data() + e0, (size_t)nent * 4);
    memcpy(rec + off_val, t.vals.data() + e0, (size_t)nent * 8);
}

static int build_stream(ocb_lu* lu, const TriSorted& L, const TriSorted& U, cudaStream_t st) {
    lu->kp_stream = 0;
    const char* env = getenv("OCB_SPTRSM_NO_STREAM");
    if (env && env[0] == '1') return OCB_OK;
    const int64_t smem_cap = (int64_t)lu->max_smem_optin - 1024;
    const TriSorted* f[2] = {&L, &U};
    std::vector<BatchPlan> plan;
    // prefer KP=4 with >= 2 stages of >= 16 KB; fall back to narrower panels
    const int kps[3] = {4, 2, 1};
    for (int ki = 0; ki < 3; ++ki) {
        const int kp = kps[ki];
        const int64_t left = smem_cap - 64 - (TRSM_THREADS / 32) * kp * 8 - lu->n * kp * 8;
        if (left < 2 * 8192) continue;
        int64_t cap = std::min<int64_t>(left / 3, 24 * 1024);
        int nst = 3;
        if (cap < 12 * 1024) { cap = std::min<int64_t>(left / 2, 24 * 1024); nst = 2; }
        cap &= ~(int64_t)15;
        if (!plan_batches(f, cap, &plan)) {
            // a long row: try two bigger stages
            cap = (left / 2) & ~(int64_t)15;
            nst = 2;
            if (!plan_batches(f, cap, &plan)) continue;
        }
        std::vector<int64_t> off(plan.size() + 1, 0);
        int64_t maxb = 0;
        for (size_t b = 0; b < plan.size(); ++b) {
            const TriSorted& t = *f[plan[b].factor];
            const int64_t rb = record_bytes(plan[b].lev1 - plan[b].lev0, plan[b].q1 - plan[b].q0,
                                            t.rowptr[plan[b].q1] - t.rowptr[plan[b].q0]);
            off[b + 1] = off[b] + rb;
            maxb = std::max(maxb, rb);
        }
        std::vector<unsigned char> img((size_t)std::max<int64_t>(off.back(), 16), 0);
        for (size_t b = 0; b < plan.size(); ++b) write_record(*f[plan[b].factor], plan[b], img.data() + off[b]);
        OCB_CUDA(cudaMalloc((void**)&lu->stream, img.size()));
        OCB_CUDA(cudaMalloc((void**)&lu->batch_off, off.size() * sizeof(int64_t)));
        OCB_CUDA(cudaMemcpyAsync(lu->stream, img.data(), img.size(), cudaMemcpyHostToDevice, st));
        OCB_CUDA(cudaMemcpyAsync(lu->batch_off, off.data(), off.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        OCB_CUDA(cudaStreamSynchronize(st));
        lu->bytes += (int64_t)img.size() + (int64_t)off.size() * 8;
        lu->nbatch = (int)plan.size();
        lu->stage_bytes = (int)std::max<int64_t>(maxb, 16);
        lu->nstages = nst;
        lu->kp_stream = kp;
        return OCB_OK;
    }
    return OCB_OK;  // stream kernel unavailable: generic kernel is used
}

// panel width / placement policy of the generic kernel
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
static int launch_panel(const ocb_lu* lu, const SolveArgs& a, cudaStream_t st) {
    const size_t smem = SMEMX ? (size_t)lu->n * KP * sizeof(double) : 0;
    if (SMEMX)
        OCB_CUDA(cudaFuncSetAttribute(sptrsm_panel_kernel<KP, SMEMX>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((a.k + KP - 1) / KP);
    sptrsm_panel_kernel<KP, SMEMX><<<grid, TRSM_THREADS, smem, st>>>(a);
    OCB_LAUNCH_CHECK();
    return OCB_OK;
}

template <int KP>
static int launch_stream(const ocb_lu* lu, const SolveArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)lu->nstages * lu->stage_bytes + 64 + (TRSM_THREADS / 32) * KP * sizeof(double) +
                        (size_t)lu->n * KP * sizeof(double);
    OCB_CUDA(cudaFuncSetAttribute(sptrsm_stream_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    const unsigned grid = (unsigned)((a.k + KP - 1) / KP);
    sptrsm_stream_kernel<KP><<<grid, TRSM_THREADS, smem, st>>>(a);
    OCB_LAUNCH_CHECK();
    return OCB_OK;
}

// optional per-launch timing of the solve kernel (bench.py roofline): CUDA events on the
// launching stream around every solve launch, summed on collect
struct SolveProf {
    bool on = false;
    std::vector<cudaEvent_t> ev;
    size_t used = 0;
    double alg_bytes = 0.0;
    long long launches = 0;
};
static SolveProf g_prof;

static int lu_solve_dispatch(const ocb_lu* lu, const double* B, int64_t ldb, int64_t nrows_b, double* X,
                             int64_t ldx, int64_t nrows_x, int64_t k, void* ws, int64_t ws_bytes,
                             cudaStream_t st);

int lu_solve_impl(const ocb_lu* lu, const double* B, int64_t ldb, int64_t nrows_b, double* X,
                  int64_t ldx, int64_t nrows_x, int64_t k, void* ws, int64_t ws_bytes,
                  cudaStream_t st) {
    if (k == 0 || lu->n == 0) return OCB_OK;
    if (!g_prof.on) return lu_solve_dispatch(lu, B, ldb, nrows_b, X, ldx, nrows_x, k, ws, ws_bytes, st);
    if (g_prof.used + 2 > g_prof.ev.size()) {
        const size_t old = g_prof.ev.size();
        g_prof.ev.resize(old + 1024);
        for (size_t i = old; i < g_prof.ev.size(); ++i) OCB_CUDA(cudaEventCreate(&g_prof.ev[i]));
    }
    OCB_CUDA(cudaEventRecord(g_prof.ev[g_prof.used], st));
    const int rc = lu_solve_dispatch(lu, B, ldb, nrows_b, X, ldx, nrows_x, k, ws, ws_bytes, st);
    OCB_CUDA(cudaEventRecord(g_prof.ev[g_prof.used + 1], st));
    g_prof.used += 2;
    g_prof.launches += 1;
    g_prof.alg_bytes += 12.0 * (double)(lu->L.nnz + lu->U.nnz + lu->n) + 16.0 * (double)(lu->n + 1) +
                        32.0 * (double)lu->n * (double)k;
    return rc;
}

static int lu_solve_dispatch(const ocb_lu* lu, const double* B, int64_t ldb, int64_t nrows_b, double* X,
                             int64_t ldx, int64_t nrows_x, int64_t k, void* ws, int64_t ws_bytes,
                             cudaStream_t st) {
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
    a.stream = lu->stream;
    a.batch_off = lu->batch_off;
    a.nbatch = lu->nbatch;
    a.stage_bytes = lu->stage_bytes;
    a.nstages = lu->nstages;
    if (lu->kp_stream == 4) return launch_stream<4>(lu, a, st);
    if (lu->kp_stream == 2) return launch_stream<2>(lu, a, st);
    if (lu->kp_stream == 1) return launch_stream<1>(lu, a, st);
    int kp;
    bool smem;
    solve_policy(lu, &kp, &smem);
    if (!smem) {
        const int64_t need = ocb_lu_solve_ws_bytes(lu, k);
        if (ws == nullptr || ws_bytes < need) {
            set_error("lu_solve: workspace too small (%lld < %lld)", (long long)ws_bytes,
                      (long long)need);
            return OCB_ERR_CAPACITY;
        }
        return launch_panel<8, false>(lu, a, st);
    }
    if (kp == 4) return launch_panel<4, true>(lu, a, st);
    return launch_panel<2, true>(lu, a, st);
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
    ocb::TriSorted Ls, Us;
    int rc = ocb::analyse_factor(n, h_L_rowptr, h_L_colidx, h_L_vals, false, &Ls);
    if (rc == OCB_OK) rc = ocb::analyse_factor(n, h_U_rowptr, h_U_colidx, h_U_vals, true, &Us);
    if (rc == OCB_OK) rc = ocb::upload_factor(Ls, &lu->L, &lu->bytes, st);
    if (rc == OCB_OK) rc = ocb::upload_factor(Us, &lu->U, &lu->bytes, st);
    if (rc == OCB_OK) {
        const size_t pb = std::max<int64_t>(n, 1) * sizeof(int32_t);
        cudaError_t e = cudaMalloc((void**)&lu->perm_r, pb);
        if (e == cudaSuccess) e = cudaMalloc((void**)&lu->perm_c, pb);
        if (e == cudaSuccess && n > 0)
            e = cudaMemcpyAsync(lu->perm_r, h_perm_r, pb, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess && n > 0)
            e = cudaMemcpyAsync(lu->perm_c, h_perm_c, pb, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // host vectors die below
        if (e != cudaSuccess) {
            ocb::set_error("lu_create: %s", cudaGetErrorString(e));
            rc = OCB_ERR_CUDA;
        }
        lu->bytes += 2 * (int64_t)pb;
    }
    if (rc == OCB_OK) rc = ocb::build_stream(lu, Ls, Us, st);
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
    cudaFree(lu->stream);
    cudaFree(lu->batch_off);
    delete lu;
    return OCB_OK;
}

int ocb_prof_enable(int on) {
    ocb::g_prof.on = on != 0;
    ocb::g_prof.used = 0;
    ocb::g_prof.alg_bytes = 0.0;
    ocb::g_prof.launches = 0;
    return OCB_OK;
}

int ocb_prof_collect(double* total_ms, int64_t* launches, double* alg_bytes) {
    OCB_ARG(total_ms && launches && alg_bytes, "prof_collect");
    double tot = 0.0;
    for (size_t i = 0; i + 1 < ocb::g_prof.used; i += 2) {
        OCB_CUDA(cudaEventSynchronize(ocb::g_prof.ev[i + 1]));
        float ms = 0.f;
        OCB_CUDA(cudaEventElapsedTime(&ms, ocb::g_prof.ev[i], ocb::g_prof.ev[i + 1]));
        tot += ms;
    }
    *total_ms = tot;
    *launches = ocb::g_prof.launches;
    *alg_bytes = ocb::g_prof.alg_bytes;
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
    info8[6] = lu->kp_stream;  // 0: generic kernel
    info8[7] = lu->nbatch;
    return OCB_OK;
}

int64_t ocb_lu_solve_ws_bytes(const ocb_lu* lu, int64_t k) {
    if (!lu || lu->kp_stream > 0) return 0;
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
