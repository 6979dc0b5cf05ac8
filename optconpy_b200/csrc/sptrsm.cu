// K1: multi-RHS sparse triangular solves  x = Pc U^-1 L^-1 Pr b  on an LU factorisation.
//
// Design (DESIGN.md "K1").  The host turns the factors into a GATHER PROGRAM
// (lu_program.h): supernodes with inverted diagonal blocks, so that the whole solve is a
// short sequence of sub-levels (139 for the N=25 cavity instead of 1784 scalar dependency
// levels), every sub-level a set of independent rows
//     xe[dst] = ((init >= 0 ? xe[init] : 0) - sum_p val[p] * xe[col[p]]) * scale,
// cut into SLICES (the unit of work of one warp, sliced-ELLPACK entry layout).
//
// sptrsm_stream_kernel<KP, CL> (the hot one at the reference's sizes).  The right-hand-side
// COLUMNS are independent: the block is cut into panels of KP columns and one thread-block
// CLUSTER of CL CTAs owns one panel for the whole solve (row permutation, all sub-levels,
// column permutation) - no grid-wide synchronisation.  The slices of every sub-level are dealt
// to the CL ranks; every CTA keeps a full copy of the panel xe in shared memory, streams its
// share of the program through a shared-memory ring fed by a producer warp with TMA bulk
// copies (cp.async.bulk + mbarrier complete_tx), and publishes every result row to all CL
// copies with st.async, counted on the level mbarrier of the destination CTA - data and
// barrier signal travel together.  A row is reduced by 2^g lanes (g per slice, chosen on the
// host from the row length) with shuffle butterflies.  The launch picks KP = 1 while one wave
// of clusters covers all columns (latency) and KP = 2 beyond.
//
// wide_level_kernel<T> (large n, very wide blocks): all columns at once, one launch per
// sub-level, lane = column; the program is read once per solve.
//
// Algorithmic bytes per solve: 12*(nnzL+nnzU) + 16*(n+1) + 32*n*k.
#include "common.cuh"
#include "lu_program.h"
#include <chrono>
#include <mutex>
#include <memory>
#include <vector>
#include <algorithm>
#include <string.h>
#include <stdlib.h>
#include <math.h>

struct ocb_lu {
    int64_t n = 0, n_ext = 0;
    int64_t nnzL = 0, nnzU = 0;
    int32_t nsub_L = 0, nsub_U = 0, nsuper = 0, max_w = 0;
    int64_t nseg = 0, nrows = 0, nent = 0;
    unsigned char* arena = nullptr;   // one device allocation: perms | batch offsets | stream
    bool arena_owned = false;         // false: the caller provided (and keeps) the buffer
    int32_t *perm_r = nullptr, *perm_c = nullptr;
    // the program is dealt to CL cluster ranks: rank r streams its own batches
    int cl = 1;
    int64_t* batch_off[8] = {nullptr};   // nbatch[r]+1 byte offsets into stream[r]
    unsigned char* stream[8] = {nullptr};
    int nbatch[8] = {0};
    int64_t bytes = 0;
    int stage_bytes = 0, nstages = 0;
    int kp_smem_max = 0;              // widest panel that fits shared memory next to the ring (0: none)
    int max_smem_optin = 0;
    // flat program for the wide (all columns at once, one launch per sub-level) executor
    bool has_flat = false;
    const int4* f_slices = nullptr;
    const int32_t *f_rowslice = nullptr, *f_dst = nullptr, *f_init = nullptr, *f_col = nullptr;
    const double *f_scale = nullptr, *f_val = nullptr;
    std::vector<int32_t> sub_row;     // host: first program row of every sub-level (nsub + 1)
    std::vector<int32_t> sub_maxlen;  // host: longest (padded) row of every sub-level
    // panel program (lu_program.h): register-blocked form for the all-columns-at-once executor
    bool has_panels = false;
    const ocb::Panel* p_panels = nullptr;
    const double *p_scale = nullptr, *p_val = nullptr;
    const int32_t* p_col = nullptr;
    int64_t npanels = 0, panel_entries = 0, panel_entries_actual = 0;
    const int32_t* p_sub_dev = nullptr;   // device copy: sub_pan (nsub + 1) followed by sub_maxcol (nsub)
    std::vector<int32_t> sub_pan;      // host: first panel of every sub-level (nsub + 1)
    std::vector<int32_t> sub_maxcol;   // host: longest column list of every sub-level
};

namespace ocb {

constexpr int TRSM_MAX_THREADS = 544;   // 16 consumer warps + the producer warp
static int trsm_threads() {
    static int t = 0;
    if (t == 0) {
        const char* env = getenv("OCB_TRSM_THREADS");
        t = env ? atoi(env) : 512;
        if (t != 128 && t != 256 && t != 384 && t != 512) t = 512;
    }
    return t;
}

struct SolveArgs {
    const int32_t *perm_r, *perm_c;
    int64_t n, n_ext;
    const double* B;
    int64_t ldb, nrows_b;
    double* X;
    int64_t ldx, nrows_x, k;
    double* ws;
    const unsigned char* stream[8];
    const int64_t* batch_off[8];
    int nbatch[8];
    int stage_bytes, nstages;
    int tma_chunk, debug_skip;
    const int* skip;    // device flag: non-zero = this launch is a no-op (ADI loop already converged)
    long long* trace;   // debug: clock64() of CTA 0 after every sub-level barrier (null: off)
};

// Batch record (every section 16-byte aligned, offsets in bytes from the record start):
//   int32 hdr[12] = {npiece, nslice, nrows, nent, off_piece, off_slice, off_dst, off_init,
//                    off_scale, off_col, off_val, -}
//   int32 piece[4*npiece] = {s0, s1, barrier, -}: slices [s0, s1) of one sub-level
//   int32 slice[4*nslice] = {ebase, trips, glog | nrows << 8, q0}  (local to the record)
//   int32 dst[nrows], int32 init[nrows], f64 scale[nrows], uint16 col[nent], f64 val[nent]
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() {
    asm volatile("fence.acq_rel.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_C:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_C;\n"
        "bra WAIT_LOOP_C;\n"
        "DONE_C:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// remote store whose completion is counted (8 bytes) on an mbarrier of the destination CTA:
// data and signal travel together, no fence on the sender
__device__ __forceinline__ void st_async_f64(uint32_t addr, double v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(addr),
                 "l"(__double_as_longlong(v)), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Sub-level barrier of a cluster: the consumers of this CTA meet at a named barrier, then ONE
// thread publishes the CTA's (local and remote) panel writes with a single cluster-scope
// fence and arrives once on the level barrier of every rank; everybody waits (acquire) on the
// barrier of the own CTA, which completes after CL arrivals.
template <int CL>
__device__ __forceinline__ void cluster_level_barrier(uint64_t* lvlbar, uint32_t& phase, int ncons, int tid) {
    asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
    if (tid == 0) {
        fence_acq_rel_cluster();
#pragma unroll
        for (int r = 0; r < CL; ++r) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(lvlbar), r));
    }
    mbar_wait_cluster(lvlbar, phase & 1);
    ++phase;
}

// one batch = several bulk copies completing on one mbarrier (more requests in flight)
__device__ __forceinline__ void issue_batch(unsigned char* dst, const unsigned char* src, uint32_t bytes,
                                            uint32_t chunk, uint64_t* bar) {
    mbar_expect_tx(bar, bytes);
    for (uint32_t o = 0; o < bytes; o += chunk)
        bulk_g2s(dst + o, src + o, min(chunk, bytes - o), bar);
}

// CL > 1: a thread-block CLUSTER of CL CTAs owns one column panel.  The slices of every
// sub-level are dealt to the CL ranks, so each CTA streams and gathers only 1/CL of the
// program; every CTA keeps a full copy of the panel xe in its shared memory, results are
// written to all CL copies through distributed shared memory (st.shared::cluster) and the
// sub-level barrier is an mbarrier in every CTA that all consumer warps of the cluster
// arrive on (release/acquire at cluster scope).
template <int KP, int CL>
__global__ void __launch_bounds__(TRSM_MAX_THREADS, 1) sptrsm_stream_kernel(const SolveArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.skip && *a.skip) return;   // uniform over the grid: written by an earlier kernel of the stream
    // layout: [ring: nstages * stage_bytes][mbarriers full/empty/level: 192 bytes][xe panel: n_ext*KP doubles]
    unsigned char* ring = smem_raw;
    uint64_t* full = (uint64_t*)(smem_raw + (size_t)a.nstages * a.stage_bytes);
    uint64_t* empty = full + 8;
    uint64_t* lvlbar = full + 16;
    const uint32_t rank = CL > 1 ? cluster_ctarank() : 0;
    const int64_t panel = (int64_t)blockIdx.x / CL;
    double* x = (double*)(smem_raw + (size_t)a.nstages * a.stage_bytes + 192);
    const int tid = threadIdx.x;
    // the last warp is the PRODUCER (feeds the ring with TMA bulk copies, runs ahead of the
    // consumers); all other warps are consumers and synchronise among themselves only
    const int warp = tid >> 5, lane = tid & 31, nwarps = (blockDim.x >> 5) - 1;
    const int ncons = nwarps * 32;
    const int S = a.nstages;
    const int64_t c0 = panel * KP;
    const int nbatch = a.nbatch[rank];
    const int64_t* batch_off = a.batch_off[rank];
    const unsigned char* stream = a.stream[rank];
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, nwarps);
        }
        mbar_init(lvlbar, 1);       // level barriers (parity of the sub-level): 1 arrival + tx bytes
        mbar_init(lvlbar + 1, 1);
        mbar_init(lvlbar + 2, CL);  // start barrier (panel copies loaded everywhere)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // every CTA's barriers exist before anyone arrives remotely
    if (warp == nwarps) {
        // batch offsets: one coalesced load per 31 batches, handed to lane 0 by shuffles
        // (a dependent global load per batch would serialise the ring at L2 latency)
        for (int b0 = 0; b0 < nbatch; b0 += 31) {
            const long long mine = __ldg((const long long*)batch_off + min(b0 + lane, nbatch));
            for (int i = 0; i < 31 && b0 + i < nbatch; ++i) {
                const long long o = __shfl_sync(0xffffffffu, mine, i);
                const long long o2 = __shfl_sync(0xffffffffu, mine, i + 1);
                if (lane == 0) {
                    const int b = b0 + i, s = b % S;
                    if (b >= S) mbar_wait(empty + s, (uint32_t)(((b / S) - 1) & 1));
                    issue_batch(ring + (size_t)s * a.stage_bytes, stream + o, (uint32_t)(o2 - o),
                                a.tma_chunk, full + s);
                }
            }
        }
        if (CL > 1) cluster_sync_all();   // stay until the cluster is done (peers write our smem)
        return;
    }
    // consumers
    uint32_t xbase[CL];   // the panel copy of every rank, as cluster shared addresses
#pragma unroll
    for (int r = 0; r < CL; ++r) xbase[r] = CL > 1 ? mapa_u32(smem_u32(x), r) : 0;
    uint32_t lvl_phase = 0;
    int trace_n = 0;
    for (int64_t e = tid; e < a.n * KP; e += ncons) {   // x[perm_r[i]] = b[i]; rows >= nrows_b are zero
        const int64_t i = e / KP;
        const int c = (int)(e - i * KP);
        double v = 0.0;
        if (i < a.nrows_b && c0 + c < a.k) v = a.B[i * a.ldb + c0 + c];
        x[(int64_t)__ldg(a.perm_r + i) * KP + c] = v;
    }
    if (CL == 1) {
        asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
    } else {   // nobody may write results into a panel copy that is still being loaded
        cluster_level_barrier<CL>(lvlbar + 2, lvl_phase, ncons, tid);
    }
    uint32_t lvl_idx = 0;          // sub-level counter: barrier lvlbar[lvl_idx & 1], phase (lvl_idx >> 1) & 1
    uint32_t barbase[CL];          // the two level barriers of every rank (cluster addresses)
#pragma unroll
    for (int r = 0; r < CL; ++r) barbase[r] = CL > 1 ? mapa_u32(smem_u32(lvlbar), r) : 0;
    if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[trace_n++] = clock64();
    for (int b = 0; b < nbatch; ++b) {
        const int s = b % S;
        mbar_wait(full + s, (uint32_t)((b / S) & 1));
        const unsigned char* rec = ring + (size_t)s * a.stage_bytes;
        const int32_t* hdr = (const int32_t*)rec;
        const int npiece = a.debug_skip ? 0 : hdr[0];
        const int4* piece = (const int4*)(rec + hdr[4]);
        const int4* slice = (const int4*)(rec + hdr[5]);
        const int32_t* dst = (const int32_t*)(rec + hdr[6]);
        const int32_t* init = (const int32_t*)(rec + hdr[7]);
        const double* scale = (const double*)(rec + hdr[8]);
        const uint16_t* col = (const uint16_t*)(rec + hdr[9]);
        const double* val = (const double*)(rec + hdr[10]);
        for (int pc = 0; pc < npiece; ++pc) {
            const int4 pd = piece[pc];
            // one slice per warp and pass: 32 >> gl rows, 2^gl lanes per row, entries trip-major
            for (int sl = pd.x + warp; sl < pd.y; sl += nwarps) {
                const int4 sd = slice[sl];
                const int gl = sd.z & 255, nr = sd.z >> 8;
                const int G = 1 << gl;
                const int r = lane >> gl;
                const bool owner = (r < nr) && ((lane & (G - 1)) == 0);
                int i0 = -1, di = 0;
                double sc = 0.0;
                if (owner) {   // row metadata: in flight while the entries stream
                    const int q = sd.w + r;
                    i0 = init[q];
                    di = dst[q];
                    sc = scale[q];
                }
                const uint16_t* cp = col + sd.x + lane;
                const double* vp = val + sd.x + lane;
                const int trips = sd.y;
                // UNR trips per step, all loads of a step issued before the first use;
                // NACC accumulators break the FP64 dependency chain
                constexpr int UNR = KP <= 2 ? 8 : 4, NACC = KP <= 2 ? 4 : 1;
                double acc[NACC][KP];
#pragma unroll
                for (int u = 0; u < NACC; ++u)
#pragma unroll
                    for (int c = 0; c < KP; ++c) acc[u][c] = 0.0;
                int u0 = 0;
                for (; u0 + UNR <= trips; u0 += UNR) {
                    int j[UNR];
                    double v[UNR];
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        j[u] = cp[(u0 + u) * 32];
                        v[u] = vp[(u0 + u) * 32];
                    }
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        if (KP == 2) {   // both columns of a row in ONE 16-byte shared-memory gather
                            const double2 xx = *reinterpret_cast<const double2*>(x + (size_t)j[u] * 2);
                            acc[u % NACC][0] = fma(v[u], xx.x, acc[u % NACC][0]);
                            acc[u % NACC][KP - 1] = fma(v[u], xx.y, acc[u % NACC][KP - 1]);
                        } else {
                            const double* xj = x + (size_t)j[u] * KP;
#pragma unroll
                            for (int c = 0; c < KP; ++c) acc[u % NACC][c] = fma(v[u], xj[c], acc[u % NACC][c]);
                        }
                    }
                }
                if (u0 < trips) {   // remainder (warp-uniform): up to UNR-1 trips, loads first
                    int j[UNR - 1];
                    double v[UNR - 1];
#pragma unroll
                    for (int u = 0; u < UNR - 1; ++u) {
                        if (u0 + u < trips) {
                            j[u] = cp[(u0 + u) * 32];
                            v[u] = vp[(u0 + u) * 32];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < UNR - 1; ++u) {
                        if (u0 + u < trips) {
                            if (KP == 2) {
                                const double2 xx = *reinterpret_cast<const double2*>(x + (size_t)j[u] * 2);
                                acc[u % NACC][0] = fma(v[u], xx.x, acc[u % NACC][0]);
                                acc[u % NACC][KP - 1] = fma(v[u], xx.y, acc[u % NACC][KP - 1]);
                            } else {
                                const double* xj = x + (size_t)j[u] * KP;
#pragma unroll
                                for (int c = 0; c < KP; ++c)
                                    acc[u % NACC][c] = fma(v[u], xj[c], acc[u % NACC][c]);
                            }
                        }
                    }
                }
                double tot[KP];
#pragma unroll
                for (int c = 0; c < KP; ++c)
                    tot[c] = NACC == 4 ? (acc[0][c] + acc[1 % NACC][c]) + (acc[2 % NACC][c] + acc[3 % NACC][c])
                                       : acc[0][c];
                for (int o = G >> 1; o > 0; o >>= 1) {
#pragma unroll
                    for (int c = 0; c < KP; ++c) tot[c] += __shfl_xor_sync(0xffffffffu, tot[c], o);
                }
                if (owner) {
                    double res[KP];
                    if (i0 >= 0) {
                        const double* xi = x + (size_t)i0 * KP;
#pragma unroll
                        for (int c = 0; c < KP; ++c) res[c] = (xi[c] - tot[c]) * sc;
                    } else {
#pragma unroll
                        for (int c = 0; c < KP; ++c) res[c] = -tot[c] * sc;
                    }
                    if (CL == 1) {
                        double* xd = x + (size_t)di * KP;
#pragma unroll
                        for (int c = 0; c < KP; ++c) xd[c] = res[c];
                    } else {
                        // every rank gets the result; each 8-byte store is counted on the level
                        // barrier of its destination, which expects rows * KP * 8 bytes
                        const uint32_t o = (uint32_t)di * (KP * 8);
                        const uint32_t bo = (lvl_idx & 1) * 8;
#pragma unroll
                        for (int r = 0; r < CL; ++r)
#pragma unroll
                            for (int c = 0; c < KP; ++c)
                                st_async_f64(xbase[r] + o + c * 8, res[c], barbase[r] + bo);
                    }
                }
            }
            if (pd.z) {
                if (CL == 1) {
                    asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
                } else {
                    // cluster-wide sub-level barrier: the barrier of this CTA completes when all
                    // pd.w rows of the sub-level (computed by whichever rank) have landed here
                    // (pd.w: rows of the sub-level | ranks without a row << 24).  A rank without
                    // rows sends an 8-byte TOKEN instead, so that every rank contributes to every
                    // barrier: nobody can run more than one sub-level ahead of anybody else, which
                    // is what makes two alternating barriers (and the y scratch parity) safe.
                    uint64_t* bar = lvlbar + (lvl_idx & 1);
                    if (tid == 0) {
                        if (pd.x == pd.y) {
                            const uint32_t bo = (lvl_idx & 1) * 8;
#pragma unroll
                            for (int r = 0; r < CL; ++r)
                                st_async_f64(xbase[r] + (uint32_t)a.n_ext * (KP * 8), 0.0, barbase[r] + bo);
                        }
                        mbar_expect_tx(bar, (uint32_t)(pd.w & 0xffffff) * (KP * 8) + (uint32_t)(pd.w >> 24) * 8);
                    }
                    mbar_wait_cluster(bar, (lvl_idx >> 1) & 1);
                    ++lvl_idx;
                }
                if (a.trace && blockIdx.x == 0 && tid == 0 && trace_n < 4000) a.trace[trace_n++] = clock64();
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);   // this warp is done with the stage
    }
    for (int64_t e = tid + (int64_t)rank * ncons; e < a.nrows_x * KP; e += (int64_t)ncons * CL) {   // out[j] = x[perm_c[j]]
        const int64_t j = e / KP;
        const int c = (int)(e - j * KP);
        if (c0 + c < a.k) a.X[j * a.ldx + c0 + c] = x[(int64_t)__ldg(a.perm_c + j) * KP + c];
    }
    if (CL > 1) cluster_sync_all();
}

// ---------------------------------------------------------------------------------
// host side: packing the program into the batch stream
// ---------------------------------------------------------------------------------
static inline int64_t a16(int64_t x) { return (x + 15) & ~(int64_t)15; }

static int64_t record_bytes(int64_t npiece, int64_t nslice, int64_t nrows, int64_t nent) {
    int64_t o = a16(HDR_INTS * 4);
    o = a16(o + npiece * 16);
    o = a16(o + nslice * 16);
    o = a16(o + nrows * 4);
    o = a16(o + nrows * 4);
    o = a16(o + nrows * 8);
    o = a16(o + nent * 2);   // 16-bit column indices: the panel has < 65536 rows whenever it fits shared memory
    o = a16(o + nent * 8);
    return o;
}

struct Piece {   // slices [s0, s1) of one sub-level placed in a batch
    int32_t s0, s1, barrier;
    int32_t rows_total;   // rows of the whole sub-level over all ranks (level barrier byte count)
};

static inline int slice_rows(const Slice& sl) { return sl.glog_nrows >> 8; }

// Greedy packing of slices into batches of at most cap bytes (a sub-level may span several
// batches).  Returns false if a single slice exceeds the capacity.
static bool plan_batches(const LuProgram& P, int64_t cap, int rank, int cl,
                         std::vector<std::vector<Piece>>* batches) {
    batches->clear();
    std::vector<Piece> cur;
    int64_t nsl = 0, rows = 0, ent = 0;
    auto flush = [&]() {
        if (cur.empty()) return;
        batches->push_back(cur);
        cur.clear();
        nsl = rows = ent = 0;
    };
    // rank r of cl owns the slices s0 + r, s0 + r + cl, ... of every sub-level (sorted by
    // decreasing work, so the deal is balanced); a Piece lists them as (first, end, stride cl)
    for (int64_t sb = 0; sb < P.nsub(); ++sb) {
        int32_t s = P.sub_ptr[sb] + rank;
        const int32_t send = P.sub_ptr[sb + 1];
        int32_t rows_total = 0;
        for (int32_t q = P.sub_ptr[sb]; q < send; ++q) rows_total += slice_rows(P.slices[q]);
        const int32_t nslices_sub = send - P.sub_ptr[sb];
        const int32_t n_empty = std::max(0, cl - nslices_sub);   // ranks r >= nslices_sub own no slice
        rows_total |= n_empty << 24;
        if (s >= send) {   // nothing for this rank: it still takes part in the barrier
            if (record_bytes((int64_t)cur.size() + 1, nsl, rows, ent) > cap) flush();
            cur.push_back(Piece{s, s, 1, rows_total});
            continue;
        }
        while (s < send) {
            int32_t s2 = s;
            int64_t r2 = rows, e2 = ent, n2 = nsl;
            while (s2 < send) {
                const int64_t rr = r2 + slice_rows(P.slices[s2]), ee = e2 + (int64_t)P.slices[s2].trips * 32;
                if (record_bytes((int64_t)cur.size() + 1, n2 + 1, rr, ee) > cap) break;
                r2 = rr;
                e2 = ee;
                ++n2;
                s2 += cl;
            }
            if (s2 == s) {
                if (cur.empty()) return false;   // one slice does not fit an empty batch
                flush();
                continue;
            }
            cur.push_back(Piece{s, s2, (s2 >= send) ? 1 : 0, rows_total});
            nsl = n2;
            rows = r2;
            ent = e2;
            s = s2;
            if (s < send) flush();
        }
    }
    flush();
    return true;
}

static void batch_counts(const LuProgram& P, const std::vector<Piece>& b, int cl, int64_t* nsl, int64_t* rows,
                         int64_t* ent) {
    *nsl = *rows = *ent = 0;
    for (const Piece& p : b)
        for (int32_t s = p.s0; s < p.s1; s += cl) {
            *nsl += 1;
            *rows += slice_rows(P.slices[s]);
            *ent += (int64_t)P.slices[s].trips * 32;
        }
}

static int64_t batch_bytes(const LuProgram& P, const std::vector<Piece>& b, int cl) {
    int64_t nsl, rows, ent;
    batch_counts(P, b, cl, &nsl, &rows, &ent);
    return record_bytes((int64_t)b.size(), nsl, rows, ent);
}

struct SliceDest {   // where the numbers of one program slice live in the image
    int32_t slice;
    int64_t val_off, scale_off;
};

static void write_record(const LuProgram& P, const std::vector<Piece>& b, int cl, unsigned char* rec,
                         int64_t rec_off = 0, std::vector<SliceDest>* dests = nullptr) {
    int64_t nsl, nrows, nent;
    batch_counts(P, b, cl, &nsl, &nrows, &nent);
    const int64_t npiece = (int64_t)b.size();
    int32_t* hdr = (int32_t*)rec;
    int64_t o = a16(HDR_INTS * 4);
    const int64_t off_piece = o;
    o = a16(o + npiece * 16);
    const int64_t off_slice = o;
    o = a16(o + nsl * 16);
    const int64_t off_dst = o;
    o = a16(o + nrows * 4);
    const int64_t off_init = o;
    o = a16(o + nrows * 4);
    const int64_t off_scale = o;
    o = a16(o + nrows * 8);
    const int64_t off_col = o;
    o = a16(o + nent * 2);
    const int64_t off_val = o;
    hdr[0] = (int32_t)npiece; hdr[1] = (int32_t)nsl; hdr[2] = (int32_t)nrows; hdr[3] = (int32_t)nent;
    hdr[4] = (int32_t)off_piece; hdr[5] = (int32_t)off_slice; hdr[6] = (int32_t)off_dst;
    hdr[7] = (int32_t)off_init; hdr[8] = (int32_t)off_scale; hdr[9] = (int32_t)off_col;
    hdr[10] = (int32_t)off_val;
    int32_t* piece = (int32_t*)(rec + off_piece);
    int32_t* slice = (int32_t*)(rec + off_slice);
    int32_t* dst = (int32_t*)(rec + off_dst);
    int32_t* init = (int32_t*)(rec + off_init);
    double* scale = (double*)(rec + off_scale);
    uint16_t* col = (uint16_t*)(rec + off_col);
    double* val = (double*)(rec + off_val);
    int32_t ls = 0, r = 0, e = 0;
    for (size_t pi = 0; pi < b.size(); ++pi) {
        const Piece& p = b[pi];
        int32_t cnt = 0;
        for (int32_t s = p.s0; s < p.s1; s += cl) ++cnt;
        piece[4 * pi + 0] = ls;
        piece[4 * pi + 1] = ls + cnt;
        piece[4 * pi + 2] = p.barrier;
        piece[4 * pi + 3] = p.barrier ? p.rows_total : 0;
        for (int32_t s = p.s0; s < p.s1; s += cl, ++ls) {
            const Slice& sl = P.slices[s];
            const int nr = slice_rows(sl), ne = sl.trips * 32;
            slice[4 * ls + 0] = e;
            slice[4 * ls + 1] = sl.trips;
            slice[4 * ls + 2] = sl.glog_nrows;
            slice[4 * ls + 3] = r;
            memcpy(dst + r, P.dst.data() + sl.q0, (size_t)nr * 4);
            memcpy(init + r, P.init.data() + sl.q0, (size_t)nr * 4);
            memcpy(scale + r, P.scale.data() + sl.q0, (size_t)nr * 8);
            for (int q = 0; q < ne; ++q) col[e + q] = (uint16_t)P.col[sl.ebase + q];
            memcpy(val + e, P.val.data() + sl.ebase, (size_t)ne * 8);
            if (dests) dests->push_back(SliceDest{s, rec_off + off_val + (int64_t)e * 8, rec_off + off_scale + (int64_t)r * 8});
            r += nr;
            e += ne;
        }
    }
}

// Self-describing host image of a factorisation: [meta int64[32]] perm_r | perm_c | batch
// offsets | batch stream.  Built without any CUDA call (worker processes build it next to the
// host LU); ocb_lu_create_from_image uploads it with one allocation and one copy.
constexpr int64_t IMG_MAGIC = 0x4f43424c55303034LL;   // "OCBLU004"
enum { M_MAGIC = 0, M_TOTAL, M_N, M_NEXT, M_NNZL, M_NNZU, M_NSUBL, M_NSUBU, M_NSUPER, M_MAXW, M_NSLICE,
       M_NROWS, M_NENT, M_OPR, M_OPC, M_CL, M_STAGEB, M_NSTAGES, M_KPSMEM, M_FLAT, M_NSUB,
       M_SHASH,   // hash of everything in the image that does not depend on the numbers (0: none)
       M_OBO = 24, M_OST = 32, M_NBATCH = 40,                  // one slot per cluster rank
       M_F_SLICE = 48, M_F_ROWSLICE, M_F_DST, M_F_INIT, M_F_SCALE, M_F_COL, M_F_VAL, M_F_SUBROW,
       M_P_PANEL = 56, M_P_SCALE, M_P_COL, M_P_VAL, M_P_SUB, M_NPANEL, M_PENT, M_PENT_ACTUAL,
       M_COUNT = 64 };

// Images of programs with the same structure differ in the two permutations and in the numbers
// only: keep the last images by structure and rewrite just those parts.
struct ImageTemplate {
    uint64_t structure_id = 0, stamp = 0;
    int max_smem_optin = 0, flags = 0;
    std::vector<unsigned char> img;
    std::vector<SliceDest> dests;
    int64_t o_pr = 0, o_pc = 0, o_fscale = -1, o_fval = -1;
    uint64_t shash = 0;       // of img with the numbers and the permutations zeroed (never 0)
};

static uint64_t hash_words(const unsigned char* p, size_t bytes) {
    uint64_t h = 0xcbf29ce484222325ULL;
    const size_t nw = bytes / 8;
    for (size_t i = 0; i < nw; ++i) {
        uint64_t w;
        memcpy(&w, p + 8 * i, 8);
        h = (h ^ w) * 0x100000001b3ULL;
        h ^= h >> 29;
    }
    return h ? h : 1;
}
static std::mutex g_img_mutex;
static std::vector<std::unique_ptr<ImageTemplate>> g_img_templates;
static uint64_t g_img_clock = 0;
constexpr size_t IMAGE_TEMPLATE_MAX_BYTES = 64u << 20;

// choose ring geometry + panel placement and pack the image
// dst / dst_capacity (optional): build the image there if it fits (*img_out == dst then),
// otherwise in a malloc'ed buffer
static int pack_image(const LuProgram& P, const int32_t* h_perm_r, const int32_t* h_perm_c,
                      int max_smem_optin, int flags, unsigned char** img_out, int64_t* bytes_out,
                      unsigned char* dst = nullptr, int64_t dst_capacity = 0) {
    const bool use_templates = P.structure_id != 0 && getenv("OCB_NO_TEMPLATE") == nullptr;
    if (use_templates) {
        std::unique_lock<std::mutex> lock(g_img_mutex);
        for (auto& t : g_img_templates)
            if (t->structure_id == P.structure_id && t->max_smem_optin == max_smem_optin && t->flags == flags) {
                t->stamp = ++g_img_clock;
                const size_t sz = t->img.size();
                unsigned char* img = (dst && (int64_t)sz <= dst_capacity) ? dst : (unsigned char*)malloc(sz);
                if (!img) {
                    set_error("lu_pack_host: out of memory");
                    return OCB_ERR_CAPACITY;
                }
                // A caller-provided buffer (a segment of the workers' pinned pool) that already
                // holds an image of this very structure keeps it: only the numbers and the
                // permutations are rewritten below (6 of 10 MB instead of 16).  The header goes
                // last, so that a half-written buffer never carries a valid hash.
                static const bool no_reuse = getenv("OCB_NO_SLOT_REUSE") != nullptr;
                const int64_t* old = (const int64_t*)img;
                const bool same = img == dst && !no_reuse && old[M_MAGIC] == IMG_MAGIC &&
                                  old[M_TOTAL] == (int64_t)sz && (uint64_t)old[M_SHASH] == t->shash;
                if (!same) {
                    ((int64_t*)img)[M_SHASH] = 0;
                    memcpy(img + M_COUNT * 8, t->img.data() + M_COUNT * 8, sz - M_COUNT * 8);
                    memcpy(img, t->img.data(), M_COUNT * 8);
                }
                if (P.n > 0) {
                    memcpy(img + t->o_pr, h_perm_r, (size_t)P.n * 4);
                    memcpy(img + t->o_pc, h_perm_c, (size_t)P.n * 4);
                }
                for (const SliceDest& d : t->dests) {
                    const Slice& sl = P.slices[d.slice];
                    memcpy(img + d.val_off, P.val.data() + sl.ebase, (size_t)sl.trips * 32 * 8);
                    memcpy(img + d.scale_off, P.scale.data() + sl.q0, (size_t)(sl.glog_nrows >> 8) * 8);
                }
                if (t->o_fscale >= 0) {
                    memcpy(img + t->o_fscale, P.scale.data(), (size_t)P.nrows() * 8);
                    memcpy(img + t->o_fval, P.val.data(), (size_t)P.nent() * 8);
                }
                *img_out = img;
                *bytes_out = (int64_t)t->img.size();
                return OCB_OK;
            }
    }
    std::unique_ptr<ImageTemplate> tmpl;
    if (use_templates) {
        tmpl.reset(new ImageTemplate());
        tmpl->structure_id = P.structure_id;
        tmpl->max_smem_optin = max_smem_optin;
        tmpl->flags = flags;
    }
    const int64_t smem_cap = (int64_t)max_smem_optin - 1024 - 192;
    const int64_t xe1 = (P.n_ext + 1) * 8;   // bytes of a one-column panel (+ the token slot)
    int kp_smem = 0, nst = 2, cl = 4;   // two large stages: a bulk copy has ~0.35 us of fixed cost
    int64_t cap = 0;
    const char* env = getenv("OCB_SPTRSM_FORCE_GLOBAL");
    const bool force_global = env && env[0] == '1';
    {   // cluster size: flags bits 4..7 (the caller's hint from the expected block width), else 4
        const int hint = (flags >> 4) & 15;
        if (hint == 1 || hint == 2 || hint == 3 || hint == 4 || hint == 8) cl = hint;
    }
    const char* e_cl = getenv("OCB_TRSM_CLUSTER");
    if (e_cl) {
        const int v = atoi(e_cl);
        if (v == 1 || v == 2 || v == 3 || v == 4 || v == 8) cl = v;
    }
    std::vector<std::vector<Piece>> batches[8];
    auto plan_all = [&](int64_t c, int ncl) {
        for (int r = 0; r < ncl; ++r)
            if (!plan_batches(P, c, r, ncl, &batches[r])) return false;
        return true;
    };
    // prefer a two-column panel next to three stages, else one column
    const int kps[2] = {2, 1};
    const char* e_st = getenv("OCB_RING_STAGES");
    const char* e_kb = getenv("OCB_RING_STAGE_KB");
    if (e_st && atoi(e_st) >= 2 && atoi(e_st) <= 8) nst = atoi(e_st);
    const int64_t want = (e_kb && atoi(e_kb) >= 4) ? (int64_t)atoi(e_kb) * 1024 : 60 * 1024;
    for (int ki = 0; ki < 2 && !force_global && kp_smem == 0; ++ki) {
        const int64_t left = smem_cap - xe1 * kps[ki];
        if (left < nst * 8192 || P.n_ext >= 65535) continue;
        cap = std::min<int64_t>(left / nst, want) & ~(int64_t)15;
        if (plan_all(cap, cl)) kp_smem = kps[ki];
    }
    if (kp_smem == 0) {   // the panel does not fit shared memory: only the wide executor applies
        cl = 0;
        for (int r = 0; r < 8; ++r) batches[r].clear();
    }
    const bool want_wide = (flags & 1) || kp_smem == 0;
    // both forms travel: the row program serves narrow blocks / small factors (most warps in
    // flight, lowest latency per sub-level), the panel program wide blocks on large factors
    static const bool no_panels = getenv("OCB_WIDE_LEGACY") != nullptr;
    const bool want_flat = want_wide;
    const bool want_panels = want_wide && !no_panels;
    PanelProgram Q;
    if (want_panels) {
        const char* e_pad = getenv("OCB_PANEL_MAXPAD");
        build_panels(P, e_pad ? atof(e_pad) : 1.6, &Q);
        tmpl.reset();   // the template refill below does not cover the panel values
    }
    std::vector<int64_t> off[8];
    int64_t maxb = 16;
    for (int r = 0; r < cl; ++r) {
        off[r].assign(batches[r].size() + 1, 0);
        for (size_t b = 0; b < batches[r].size(); ++b) {
            const int64_t rb = batch_bytes(P, batches[r][b], cl);
            off[r][b + 1] = off[r][b] + rb;
            maxb = std::max(maxb, rb);
        }
    }
    const int64_t o_pr = align_up(M_COUNT * 8, 256);
    const int64_t o_pc = align_up(o_pr + std::max<int64_t>(P.n, 1) * 4, 256);
    int64_t o = align_up(o_pc + std::max<int64_t>(P.n, 1) * 4, 256);
    int64_t o_bo[8], o_st[8];
    for (int r = 0; r < cl; ++r) {
        o_bo[r] = o;
        o = align_up(o + (int64_t)off[r].size() * 8, 256);
    }
    for (int r = 0; r < cl; ++r) {
        o_st[r] = o;
        o = align_up(o + std::max<int64_t>(off[r].back(), 16), 256);
    }
    // flat program (wide executor): slices | row -> slice | dst | init | scale | col | val | sub rows
    const int64_t nsl = (int64_t)P.slices.size(), nr = P.nrows(), ne = P.nent(), nsub = P.nsub();
    int64_t o_f[8] = {0};
    if (want_flat) {
        const int64_t sz[8] = {nsl * 16, nr * 4, nr * 4, nr * 4, nr * 8, ne * 4, ne * 8, 2 * (nsub + 1) * 4};
        for (int i = 0; i < 8; ++i) {
            o_f[i] = o;
            o = align_up(o + std::max<int64_t>(sz[i], 16), 256);
        }
    }
    int64_t o_p[5] = {0};
    if (want_panels) {
        const int64_t npn = (int64_t)Q.panels.size(), npc = (int64_t)Q.pcol.size(), nsb = Q.nsub();
        const int64_t sz[5] = {npn * (int64_t)sizeof(Panel), npn * PANEL_ROWS * 8, npc * 4,
                               npc * PANEL_ROWS * 8, 2 * (nsb + 1) * 4};
        for (int i = 0; i < 5; ++i) {
            o_p[i] = o;
            o = align_up(o + std::max<int64_t>(sz[i], 16), 256);
        }
    }
    const int64_t total = o;
    unsigned char* img;
    if (dst && total <= dst_capacity) {
        img = dst;
        memset(img, 0, (size_t)total);
    } else {
        img = (unsigned char*)calloc((size_t)total, 1);   // zero pages on demand
    }
    if (!img) {
        set_error("lu_pack_host: out of memory");
        return OCB_ERR_CAPACITY;
    }
    *img_out = img;
    *bytes_out = total;
    int64_t* meta = (int64_t*)img;
    meta[M_MAGIC] = IMG_MAGIC; meta[M_TOTAL] = total; meta[M_N] = P.n; meta[M_NEXT] = P.n_ext;
    meta[M_NNZL] = P.nnzL; meta[M_NNZU] = P.nnzU; meta[M_NSUBL] = P.nsub_L; meta[M_NSUBU] = P.nsub_U;
    meta[M_NSUPER] = P.nsuper; meta[M_MAXW] = P.max_w; meta[M_NSLICE] = (int64_t)P.slices.size();
    meta[M_NROWS] = P.nrows(); meta[M_NENT] = P.nent(); meta[M_OPR] = o_pr; meta[M_OPC] = o_pc;
    meta[M_CL] = cl; meta[M_STAGEB] = maxb; meta[M_NSTAGES] = nst; meta[M_KPSMEM] = kp_smem;
    for (int r = 0; r < cl; ++r) {
        meta[M_OBO + r] = o_bo[r];
        meta[M_OST + r] = o_st[r];
        meta[M_NBATCH + r] = (int64_t)batches[r].size();
    }
    if (P.n > 0) {
        memcpy(img + o_pr, h_perm_r, (size_t)P.n * 4);
        memcpy(img + o_pc, h_perm_c, (size_t)P.n * 4);
    }
    for (int r = 0; r < cl; ++r) {
        memcpy(img + o_bo[r], off[r].data(), off[r].size() * 8);
        for (size_t b = 0; b < batches[r].size(); ++b)
            write_record(P, batches[r][b], cl, img + o_st[r] + off[r][b], o_st[r] + off[r][b],
                         tmpl ? &tmpl->dests : nullptr);
    }
    meta[M_FLAT] = want_flat ? 1 : 0;
    meta[M_NSUB] = nsub;
    if (want_flat) {
        for (int i = 0; i < 8; ++i) meta[M_F_SLICE + i] = o_f[i];
        memcpy(img + o_f[0], P.slices.data(), (size_t)nsl * 16);
        int32_t* rowslice = (int32_t*)(img + o_f[1]);
        for (int64_t sidx = 0; sidx < nsl; ++sidx)
            for (int r = 0; r < slice_rows(P.slices[sidx]); ++r) rowslice[P.slices[sidx].q0 + r] = (int32_t)sidx;
        memcpy(img + o_f[2], P.dst.data(), (size_t)nr * 4);
        memcpy(img + o_f[3], P.init.data(), (size_t)nr * 4);
        memcpy(img + o_f[4], P.scale.data(), (size_t)nr * 8);
        memcpy(img + o_f[5], P.col.data(), (size_t)ne * 4);
        memcpy(img + o_f[6], P.val.data(), (size_t)ne * 8);
        int32_t* subrow = (int32_t*)(img + o_f[7]);
        for (int64_t sb = 0; sb < nsub; ++sb)
            subrow[sb] = P.sub_ptr[sb] < nsl ? P.slices[P.sub_ptr[sb]].q0 : (int32_t)nr;
        subrow[nsub] = (int32_t)nr;
        int32_t* submax = subrow + nsub + 1;   // the first slice of a sub-level holds its longest row
        for (int64_t sb = 0; sb < nsub; ++sb) {
            const Slice& f = P.slices[P.sub_ptr[sb]];
            submax[sb] = P.sub_ptr[sb] < nsl ? (f.trips << (f.glog_nrows & 255)) : 0;
        }
    }
    meta[M_NPANEL] = 0;
    if (want_panels) {
        const int64_t npn = (int64_t)Q.panels.size(), npc = (int64_t)Q.pcol.size(), nsb = Q.nsub();
        for (int i = 0; i < 5; ++i) meta[M_P_PANEL + i] = o_p[i];
        meta[M_NPANEL] = npn;
        meta[M_PENT] = npc;
        meta[M_PENT_ACTUAL] = Q.entries_actual;
        meta[M_NSUB] = nsb;
        if (npn) memcpy(img + o_p[0], Q.panels.data(), (size_t)npn * sizeof(Panel));
        if (npn) memcpy(img + o_p[1], Q.scale.data(), (size_t)npn * PANEL_ROWS * 8);
        if (npc) memcpy(img + o_p[2], Q.pcol.data(), (size_t)npc * 4);
        if (npc) memcpy(img + o_p[3], Q.pval.data(), (size_t)npc * PANEL_ROWS * 8);
        int32_t* sp = (int32_t*)(img + o_p[4]);
        memcpy(sp, Q.sub_ptr.data(), (size_t)(nsb + 1) * 4);
        int32_t* smax = sp + nsb + 1;   // panels of a sub-level are sorted: the first is the longest
        for (int64_t sb = 0; sb < nsb; ++sb)
            smax[sb] = Q.sub_ptr[sb] < Q.sub_ptr[sb + 1] ? Q.panels[Q.sub_ptr[sb]].ncol : 0;
    }
    if (tmpl && (size_t)total <= IMAGE_TEMPLATE_MAX_BYTES) {
        tmpl->img.assign(img, img + total);
        tmpl->o_pr = o_pr;
        tmpl->o_pc = o_pc;
        if (want_flat) {
            tmpl->o_fscale = o_f[4];
            tmpl->o_fval = o_f[6];
        }
        {   // the template keeps the structure only: numbers and permutations zeroed, then hashed
            unsigned char* ti = tmpl->img.data();
            if (P.n > 0) {
                memset(ti + o_pr, 0, (size_t)P.n * 4);
                memset(ti + o_pc, 0, (size_t)P.n * 4);
            }
            for (const SliceDest& d : tmpl->dests) {
                const Slice& sl = P.slices[d.slice];
                memset(ti + d.val_off, 0, (size_t)sl.trips * 32 * 8);
                memset(ti + d.scale_off, 0, (size_t)(sl.glog_nrows >> 8) * 8);
            }
            if (want_flat) {
                memset(ti + o_f[4], 0, (size_t)P.nrows() * 8);
                memset(ti + o_f[6], 0, (size_t)P.nent() * 8);
            }
            ((int64_t*)ti)[M_SHASH] = 0;
            tmpl->shash = hash_words(ti, (size_t)total);
            ((int64_t*)ti)[M_SHASH] = (int64_t)tmpl->shash;
            meta[M_SHASH] = (int64_t)tmpl->shash;
        }
        std::unique_lock<std::mutex> lock(g_img_mutex);
        tmpl->stamp = ++g_img_clock;
        if (g_img_templates.size() >= 4) {                      // keep the four most recently used
            size_t oldest = 0;
            for (size_t j = 1; j < g_img_templates.size(); ++j)
                if (g_img_templates[j]->stamp < g_img_templates[oldest]->stamp) oldest = j;
            g_img_templates.erase(g_img_templates.begin() + oldest);
        }
        g_img_templates.push_back(std::move(tmpl));
    }
    return OCB_OK;
}

static int upload_image(ocb_lu* lu, const unsigned char* img, int64_t bytes, void* d_arena, cudaStream_t st) {
    const int64_t* meta = (const int64_t*)img;
    if (bytes < M_COUNT * 8 || meta[M_MAGIC] != IMG_MAGIC || meta[M_TOTAL] != bytes) {
        set_error("lu_create_from_image: not a factor image (bad magic or size)");
        return OCB_ERR_ARG;
    }
    lu->n = meta[M_N]; lu->n_ext = meta[M_NEXT]; lu->nnzL = meta[M_NNZL]; lu->nnzU = meta[M_NNZU];
    lu->nsub_L = (int32_t)meta[M_NSUBL]; lu->nsub_U = (int32_t)meta[M_NSUBU];
    lu->nsuper = (int32_t)meta[M_NSUPER]; lu->max_w = (int32_t)meta[M_MAXW];
    lu->nseg = meta[M_NSLICE]; lu->nrows = meta[M_NROWS]; lu->nent = meta[M_NENT];
    if (d_arena) {
        if (((uintptr_t)d_arena & 255) != 0) {
            set_error("lu_create_from_image: the device buffer must be 256-byte aligned");
            return OCB_ERR_ARG;
        }
        lu->arena = (unsigned char*)d_arena;
        lu->arena_owned = false;
    } else {
        OCB_CUDA(cudaMalloc((void**)&lu->arena, (size_t)bytes));
        lu->arena_owned = true;
    }
    OCB_CUDA(cudaMemcpyAsync(lu->arena, img, (size_t)bytes, cudaMemcpyHostToDevice, st));
    OCB_CUDA(cudaStreamSynchronize(st));   // the caller's buffer may go away
    lu->perm_r = (int32_t*)(lu->arena + meta[M_OPR]);
    lu->perm_c = (int32_t*)(lu->arena + meta[M_OPC]);
    lu->cl = (int)meta[M_CL];
    for (int r = 0; r < lu->cl; ++r) {
        lu->batch_off[r] = (int64_t*)(lu->arena + meta[M_OBO + r]);
        lu->stream[r] = lu->arena + meta[M_OST + r];
        lu->nbatch[r] = (int)meta[M_NBATCH + r];
    }
    lu->bytes = bytes;
    lu->stage_bytes = (int)meta[M_STAGEB];
    lu->nstages = (int)meta[M_NSTAGES];
    lu->kp_smem_max = (int)meta[M_KPSMEM];
    lu->has_flat = meta[M_FLAT] != 0;
    if (lu->has_flat) {
        lu->f_slices = (const int4*)(lu->arena + meta[M_F_SLICE]);
        lu->f_rowslice = (const int32_t*)(lu->arena + meta[M_F_ROWSLICE]);
        lu->f_dst = (const int32_t*)(lu->arena + meta[M_F_DST]);
        lu->f_init = (const int32_t*)(lu->arena + meta[M_F_INIT]);
        lu->f_scale = (const double*)(lu->arena + meta[M_F_SCALE]);
        lu->f_col = (const int32_t*)(lu->arena + meta[M_F_COL]);
        lu->f_val = (const double*)(lu->arena + meta[M_F_VAL]);
        const int32_t* sr = (const int32_t*)(img + meta[M_F_SUBROW]);
        lu->sub_row.assign(sr, sr + meta[M_NSUB] + 1);
        lu->sub_maxlen.assign(sr + meta[M_NSUB] + 1, sr + 2 * meta[M_NSUB] + 1);
    }
    lu->has_panels = meta[M_NPANEL] > 0;
    if (lu->has_panels) {
        lu->p_panels = (const ocb::Panel*)(lu->arena + meta[M_P_PANEL]);
        lu->p_scale = (const double*)(lu->arena + meta[M_P_SCALE]);
        lu->p_col = (const int32_t*)(lu->arena + meta[M_P_COL]);
        lu->p_val = (const double*)(lu->arena + meta[M_P_VAL]);
        lu->npanels = meta[M_NPANEL];
        lu->panel_entries = meta[M_PENT];
        lu->panel_entries_actual = meta[M_PENT_ACTUAL];
        const int32_t* sp = (const int32_t*)(img + meta[M_P_SUB]);
        lu->p_sub_dev = (const int32_t*)(lu->arena + meta[M_P_SUB]);
        lu->sub_pan.assign(sp, sp + meta[M_NSUB] + 1);
        lu->sub_maxcol.assign(sp + meta[M_NSUB] + 1, sp + 2 * meta[M_NSUB] + 1);
    }
    return OCB_OK;
}

template <int KP, int CL>
static int launch_stream(const ocb_lu* lu, const SolveArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)lu->nstages * lu->stage_bytes + 192 +
                        (size_t)(lu->n_ext + 1) * KP * sizeof(double);
    OCB_CUDA(cudaFuncSetAttribute(sptrsm_stream_kernel<KP, CL>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned npanels = (unsigned)((a.k + KP - 1) / KP);
    const unsigned threads = trsm_threads() + 32;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(npanels * CL, 1, 1);
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    OCB_CUDA(cudaLaunchKernelEx(&cfg, sptrsm_stream_kernel<KP, CL>, a));
    g_launches.fetch_add(1);
    return OCB_OK;
}

// clusters of CL CTAs (one CTA per SM at this shared-memory size) that can be resident at once
template <int KP, int CL>
static int max_clusters(const ocb_lu* lu) {
    // per (KP, CL) instantiation and shared-memory size (factors of different problems differ)
    static size_t cache_smem[9] = {0};
    static int cache[9] = {0};
    const int slot = KP;
    const size_t smem = (size_t)lu->nstages * lu->stage_bytes + 192 + (size_t)(lu->n_ext + 1) * KP * sizeof(double);
    if (cache[slot] > 0 && cache_smem[slot] == smem) return cache[slot];
    cache_smem[slot] = smem;
    cudaFuncSetAttribute(sptrsm_stream_kernel<KP, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(CL * 64, 1, 1);
    cfg.blockDim = dim3(trsm_threads() + 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, sptrsm_stream_kernel<KP, CL>, &cfg) != cudaSuccess || nc <= 0) {
        cudaGetLastError();
        nc = sm_count() / CL;
    }
    cache[slot] = nc;
    return nc;
}

template <int CL>
static int launch_cluster(const ocb_lu* lu, const SolveArgs& a, cudaStream_t st) {
    // panel width: as narrow as possible while one wave of clusters covers all columns
    int kp = 1;
    const char* env = getenv("OCB_SPTRSM_KP");
    if (env && (atoi(env) == 1 || atoi(env) == 2) && atoi(env) <= lu->kp_smem_max) {
        kp = atoi(env);
    } else if (lu->kp_smem_max >= 2) {
        const int cap1 = CL > 1 ? max_clusters<1, CL>(lu) : sm_count();
        if (a.k > cap1) kp = 2;
    }
    if (kp == 2) return launch_stream<2, CL>(lu, a, st);
    return launch_stream<1, CL>(lu, a, st);
}

// ---------------------------------------------------------------------------------
// wide executor: ALL right-hand sides at once, one launch per sub-level
// ---------------------------------------------------------------------------------
// For factors whose panel does not fit shared memory (n_ext > ~27 000) and for very wide
// blocks (k >= OCB_WIDE_MIN_K = 640, image built with the flat program) the column-panel kernel
// above would stream the whole program once per panel.  Here the program is read ONCE per
// solve: the extended block xe (n_ext x ldx, row-major, in the caller's workspace) stays in
// HBM/L2, every sub-level is one kernel launch (the launch boundary is the barrier), and a warp
// owns one program row and 32*T consecutive COLUMNS: lane = column, so the gathers
// xe[col[p], c0 + lane] are 256-byte coalesced rows, col/val are warp-uniform (broadcast)
// loads, and there is no reduction at all.
struct WideArgs {
    const int4* slices;
    const int32_t *rowslice, *dst, *init, *col;
    const double *scale, *val;
    double* xe;
    int64_t ldx;
    int ntile;   // column tiles of 32*T per row
    const int* skip;
};

static int wide_min_k() {
    static int v = -1;
    if (v < 0) {
        const char* env = getenv("OCB_WIDE_MIN_K");
        v = env ? atoi(env) : 640;
        if (v < 1) v = 1;
    }
    return v;
}
static bool use_wide(const ocb_lu* lu, int64_t k) {
    return (lu->has_flat || lu->has_panels) && (lu->kp_smem_max == 0 || k >= wide_min_k());
}
static int wide_tiles(int64_t k) { return k <= 32 ? 1 : (k <= 64 ? 2 : 4); }
static int64_t wide_ldx(int64_t k) {
    const int64_t w = 32 * wide_tiles(k);
    return (k + w - 1) / w * w;
}

__global__ void __launch_bounds__(256) wide_load_kernel(const SolveArgs a, double* xe, int64_t ldx) {
    if (a.skip && *a.skip) return;
    // xe[perm_r[i], :] = [b[i, :], 0...]; rows >= nrows_b are zero
    const int64_t total = a.n * ldx;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / ldx, c = e - i * ldx;
        double v = 0.0;
        if (i < a.nrows_b && c < a.k) v = a.B[i * a.ldb + c];
        xe[(int64_t)__ldg(a.perm_r + i) * ldx + c] = v;
    }
}

__global__ void __launch_bounds__(256) wide_store_kernel(const SolveArgs a, const double* xe, int64_t ldx) {
    if (a.skip && *a.skip) return;
    const int64_t total = a.nrows_x * a.k;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = e / a.k, c = e - j * a.k;
        a.X[j * a.ldx + c] = xe[(int64_t)__ldg(a.perm_c + j) * ldx + c];
    }
}

// One CTA = 8 warps = (8 >> wlog) program rows x one tile of 32*T columns; the 2^wlog warps of a
// row take alternating chunks of UNR entries (a long row would otherwise be one serial chain
// of L2-latency-bound gathers) and their partial sums are combined through shared memory.
template <int T>
__global__ void __launch_bounds__(256) wide_level_kernel(const WideArgs a, int q0, int q1, int wlog) {
    __shared__ double red[8][T][32];
    if (a.skip && *a.skip) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpr = 1 << wlog, rows_per_cta = 8 >> wlog;
    const int rowgroup = blockIdx.x / a.ntile, tile = blockIdx.x - rowgroup * a.ntile;
    const int row = q0 + rowgroup * rows_per_cta + (warp >> wlog);
    const int wr = warp & (wpr - 1);
    const bool valid = row < q1;
    const int64_t c0 = (int64_t)tile * (32 * T) + lane;
    const double* xc = a.xe + c0;
    double acc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = 0.0;
    if (valid) {
        const int4 sl = __ldg(a.slices + __ldg(a.rowslice + row));
        const int gl = sl.z & 255, G = 1 << gl;
        const int r = row - sl.w;
        const int32_t* cp = a.col + sl.x + (r << gl);
        const double* vp = a.val + sl.x + (r << gl);
        const int nent = sl.y << gl;   // padded entries of this row (zero padding multiplies xe[0])
        constexpr int UNR = 8;
        for (int e0 = wr * UNR; e0 < nent; e0 += wpr * UNR) {
            int j[UNR];
            double v[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int e = e0 + u;
                const bool ok = e < nent;
                const int p = ((e >> gl) << 5) + (e & (G - 1));
                j[u] = ok ? __ldg(cp + p) : 0;
                v[u] = ok ? __ldg(vp + p) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const double* xr = xc + (int64_t)j[u] * a.ldx;
#pragma unroll
                for (int t = 0; t < T; ++t) acc[t] = fma(v[u], xr[32 * t], acc[t]);
            }
        }
    }
    if (wlog > 0) {
#pragma unroll
        for (int t = 0; t < T; ++t) red[warp][t][lane] = acc[t];
        __syncthreads();
        if (wr == 0) {
            for (int w2 = 1; w2 < wpr; ++w2)
#pragma unroll
                for (int t = 0; t < T; ++t) acc[t] += red[warp + w2][t][lane];
        }
    }
    if (valid && wr == 0) {
        const int i0 = __ldg(a.init + row);
        const double sc = __ldg(a.scale + row);
        double* xd = a.xe + (int64_t)__ldg(a.dst + row) * a.ldx + c0;
        if (i0 >= 0) {
            const double* xi = xc + (int64_t)i0 * a.ldx;
#pragma unroll
            for (int t = 0; t < T; ++t) xd[32 * t] = (xi[32 * t] - acc[t]) * sc;
        } else {
#pragma unroll
            for (int t = 0; t < T; ++t) xd[32 * t] = -acc[t] * sc;
        }
    }
}

static int wide_solve(const ocb_lu* lu, const SolveArgs& a, cudaStream_t st) {
    const int T = wide_tiles(a.k);
    const int64_t ldx = wide_ldx(a.k);
    double* xe = a.ws;
    const unsigned lblocks = (unsigned)std::min<int64_t>((a.n * ldx + 255) / 256, 148 * 16);
    wide_load_kernel<<<lblocks, 256, 0, st>>>(a, xe, ldx);
    OCB_LAUNCH_CHECK();
    WideArgs w;
    w.slices = lu->f_slices; w.rowslice = lu->f_rowslice; w.dst = lu->f_dst; w.init = lu->f_init;
    w.col = lu->f_col; w.scale = lu->f_scale; w.val = lu->f_val;
    w.xe = xe; w.ldx = ldx;
    w.skip = a.skip;
    w.ntile = (int)(ldx / (32 * T));
    const int nsub = (int)lu->sub_row.size() - 1;
    for (int sb = 0; sb < nsub; ++sb) {
        const int q0 = lu->sub_row[sb], q1 = lu->sub_row[sb + 1];
        if (q1 <= q0) continue;
        // warps per row from the longest row of the sub-level (rows are sorted by length)
        const int maxlen = lu->sub_maxlen[sb];
        const int wlog = maxlen > 256 ? 3 : (maxlen > 96 ? 2 : (maxlen > 32 ? 1 : 0));
        const int rows_per_cta = 8 >> wlog;
        const unsigned blocks = (unsigned)(((q1 - q0) + rows_per_cta - 1) / rows_per_cta) * (unsigned)w.ntile;
        if (T == 1) wide_level_kernel<1><<<blocks, 256, 0, st>>>(w, q0, q1, wlog);
        else if (T == 2) wide_level_kernel<2><<<blocks, 256, 0, st>>>(w, q0, q1, wlog);
        else wide_level_kernel<4><<<blocks, 256, 0, st>>>(w, q0, q1, wlog);
        OCB_LAUNCH_CHECK();
    }
    const unsigned sblocks = (unsigned)std::min<int64_t>((a.nrows_x * a.k + 255) / 256, 148 * 16);
    wide_store_kernel<<<std::max(1u, sblocks), 256, 0, st>>>(a, xe, ldx);
    OCB_LAUNCH_CHECK();
    return OCB_OK;
}

// ---------------------------------------------------------------------------------
// panel executor: the wide executor with REGISTER BLOCKING over the rows of a supernode
// ---------------------------------------------------------------------------------
// A panel = up to 8 rows of one sub-level that share one column list (lu_program.h).  A warp
// owns one panel and 32*T consecutive columns (lane = column): per list entry it loads ONE x row
// segment (coalesced 256 bytes per 32 columns) and the 8 row values (one uniform 64-byte load)
// and issues 8*T FMAs - 1 byte of x per FMA instead of the 8 of the row-by-row executor, which
// turns the kernel from L2-bandwidth bound into FP64 bound.  Long lists are split over 2^wlog
// warps of the CTA and combined through shared memory (fixed order: deterministic).
struct PanelArgs {
    const Panel* panels;
    const double *scale, *val;
    const int32_t* col;
    double* xe;
    int64_t ldx;
    int ntile;
    const int* skip;
};

// Inner loop of both panel executors: one warp accumulates acc[8][T] += val[e][8] * xe[col[e], tile]
// over the list entries [e0, e1) of one panel.  The gathers are latency bound (an x row segment
// comes out of the L2, ~600+ cycles), so every warp runs its OWN cp.async pipeline: the x row
// segments (32*T columns) and the 8 row values of the next NST-1 stages of 8 entries are in
// flight into the warp's private shared-memory ring while it computes the current stage from
// shared memory (conflict-free column reads, broadcast value reads).  No CTA-wide barrier.
constexpr int PSTAGE_E = 8;   // list entries per stage

template <int T>
__host__ __device__ constexpr int panel_stage_doubles() { return PSTAGE_E * 32 * T + PSTAGE_E * PANEL_ROWS; }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool l1) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (l1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// L1: x gathers may allocate in L1 (by-level executor: L1 is flushed at every launch) or must
// bypass it (persistent executor: other SMs rewrite x rows between sub-levels of ONE launch)
template <int T, int NST, bool L1>
__device__ __forceinline__ void panel_accumulate(double (&acc)[PANEL_ROWS][T], double* __restrict__ ring,
                                                 const int32_t* __restrict__ cp, const double* __restrict__ vp,
                                                 const double* __restrict__ xtile, int64_t ldx, int e0, int e1,
                                                 int lane) {
    constexpr int SD = panel_stage_doubles<T>();
    const int nstage = (e1 - e0 + PSTAGE_E - 1) / PSTAGE_E;
    if (nstage <= 0) return;
    // column indices one stage ahead of the copies that need them (uniform registers)
    int4 ja = __ldg((const int4*)(cp + e0));
    int4 jb = (e0 + 4 < e1) ? __ldg((const int4*)(cp + e0 + 4)) : make_int4(0, 0, 0, 0);
    auto issue = [&](int s) {
        const int e = e0 + s * PSTAGE_E;
        double* xs = ring + (size_t)(s % NST) * SD;
        double* vs = xs + PSTAGE_E * 32 * T;
        const int jj[PSTAGE_E] = {ja.x, ja.y, ja.z, ja.w, jb.x, jb.y, jb.z, jb.w};
        const int nv = min(PSTAGE_E, e1 - e);   // 4 or 8 (lists are padded to multiples of 4)
#pragma unroll
        for (int u = 0; u < PSTAGE_E; ++u) {
            if (u < nv) {
                const double* src = xtile + (int64_t)jj[u] * ldx;
#pragma unroll
                for (int h = 0; h < (T + 1) / 2; ++h) {
                    const int c = 2 * lane + 64 * h;
                    if (c < 32 * T) cp_async16(xs + u * 32 * T + c, src + c, L1);
                }
            }
        }
        if (2 * lane < nv * PANEL_ROWS) cp_async16(vs + 2 * lane, vp + (int64_t)e * PANEL_ROWS + 2 * lane, true);
        // indices of the stage after this one
        const int en = e + PSTAGE_E;
        ja = (en < e1) ? __ldg((const int4*)(cp + en)) : make_int4(0, 0, 0, 0);
        jb = (en + 4 < e1) ? __ldg((const int4*)(cp + en + 4)) : make_int4(0, 0, 0, 0);
    };
#pragma unroll
    for (int s = 0; s < NST - 1; ++s) {
        if (s < nstage) issue(s);
        cp_async_commit();
    }
    for (int s = 0; s < nstage; ++s) {
        if (s + NST - 1 < nstage) issue(s + NST - 1);
        cp_async_commit();
        cp_async_wait<NST - 1>();
        __syncwarp();
        const double* xs = ring + (size_t)(s % NST) * SD;
        const double* vs = xs + PSTAGE_E * 32 * T;
        const int nv = min(PSTAGE_E, e1 - (e0 + s * PSTAGE_E));
#pragma unroll
        for (int u = 0; u < PSTAGE_E; ++u) {
            if (u < nv) {
                double xv[T];
#pragma unroll
                for (int t = 0; t < T; ++t) xv[t] = xs[u * 32 * T + lane + 32 * t];
                const double2* v2 = (const double2*)(vs + u * PANEL_ROWS);
#pragma unroll
                for (int h = 0; h < PANEL_ROWS / 2; ++h) {
                    const double2 vv = v2[h];
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        acc[2 * h][t] = fma(vv.x, xv[t], acc[2 * h][t]);
                        acc[2 * h + 1][t] = fma(vv.y, xv[t], acc[2 * h + 1][t]);
                    }
                }
            }
        }
        __syncwarp();   // the ring slot is overwritten by the copies issued next
    }
    cp_async_wait<0>();
}

template <int T, int NST>
__global__ void __launch_bounds__(256) panel_level_kernel(const PanelArgs a, int p0, int p1, int wlog) {
    extern __shared__ __align__(16) double psm[];   // per warp: NST stages (also the split-list reduce buffer)
    if (a.skip && *a.skip) return;
    constexpr int WD = NST * panel_stage_doubles<T>();
    static_assert(WD >= PANEL_ROWS * T * 32, "reduce buffer must fit the ring");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpr = 1 << wlog, ppc = 8 >> wlog;
    const int grp = blockIdx.x / a.ntile, tile = blockIdx.x - grp * a.ntile;
    const int pi = p0 + grp * ppc + (warp >> wlog);
    const int wr = warp & (wpr - 1);
    const bool valid = pi < p1;
    double* xt = a.xe + (int64_t)tile * (32 * T);
    double* ring = psm + (size_t)warp * WD;
    double acc[PANEL_ROWS][T];
#pragma unroll
    for (int r = 0; r < PANEL_ROWS; ++r)
#pragma unroll
        for (int t = 0; t < T; ++t) acc[r][t] = 0.0;
    int4 pn0 = make_int4(0, 0, 0, -1);
    int nrows = 0;
    if (valid) {
        pn0 = __ldg((const int4*)(a.panels + pi));          // cbase, ncol, dst0, init0
        nrows = __ldg(&a.panels[pi].nrows);
        const int ncol = pn0.y;                               // multiple of 4
        const int per = (((ncol + 7) >> 3) + wpr - 1) / wpr * 8;    // chunk of this warp: whole stages
        const int e0 = wr * per, e1 = min(ncol, e0 + per);
        panel_accumulate<T, NST, true>(acc, ring, a.col + pn0.x, a.val + (int64_t)pn0.x * PANEL_ROWS, xt, a.ldx,
                                       e0, e1, lane);
    }
    if (wlog > 0) {
        __syncthreads();
        double* mine = ring;
#pragma unroll
        for (int r = 0; r < PANEL_ROWS; ++r)
#pragma unroll
            for (int t = 0; t < T; ++t) mine[(r * T + t) * 32 + lane] = acc[r][t];
        __syncthreads();
        if (wr == 0) {
            for (int w2 = 1; w2 < wpr; ++w2) {
                const double* other = psm + (size_t)(warp + w2) * WD;
#pragma unroll
                for (int r = 0; r < PANEL_ROWS; ++r)
#pragma unroll
                    for (int t = 0; t < T; ++t) acc[r][t] += other[(r * T + t) * 32 + lane];
            }
        }
    }
    if (valid && wr == 0) {
        const double* sc = a.scale + (int64_t)pi * PANEL_ROWS;
        double* xc = xt + lane;
#pragma unroll
        for (int r = 0; r < PANEL_ROWS; ++r) {
            if (r < nrows) {
                const double s = __ldg(sc + r);
                double* xd = xc + (int64_t)(pn0.z + r) * a.ldx;
                if (pn0.w >= 0) {
                    const double* xi = xc + (int64_t)(pn0.w + r) * a.ldx;
#pragma unroll
                    for (int t = 0; t < T; ++t) xd[32 * t] = (xi[32 * t] - acc[r][t]) * s;
                } else {
#pragma unroll
                    for (int t = 0; t < T; ++t) xd[32 * t] = -acc[r][t] * s;
                }
            }
        }
    }
}

template <int T, int NST>
static int panel_launch(const PanelArgs& w, int p0, int p1, int wlog, cudaStream_t st) {
    const int ppc = 8 >> wlog;
    const unsigned blocks = (unsigned)(((p1 - p0) + ppc - 1) / ppc) * (unsigned)w.ntile;
    const size_t smem = (size_t)8 * NST * panel_stage_doubles<T>() * sizeof(double);
    static bool attr = false;
    if (!attr) {
        OCB_CUDA(cudaFuncSetAttribute(panel_level_kernel<T, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    panel_level_kernel<T, NST><<<blocks, 256, smem, st>>>(w, p0, p1, wlog);
    OCB_LAUNCH_CHECK();
    return OCB_OK;
}

// column tiles of 32 * T per warp: narrow tiles (T = 1) give the most warps in flight per SM (the
// gathers are latency bound); OCB_PANEL_T / OCB_PANEL_U override for experiments
static int panel_tiles(int64_t k) {
    static int force = -1;
    if (force < 0) {
        const char* e = getenv("OCB_PANEL_T");
        force = e ? atoi(e) : 0;
    }
    if (force == 1 || force == 2 || force == 4) return (k <= 32) ? 1 : ((k <= 64 && force > 2) ? 2 : force);
    return k <= 32 ? 1 : 2;
}
static int64_t panel_ldx(int64_t k) {
    const int64_t w = 32 * panel_tiles(k);
    return (k + w - 1) / w * w;
}

// ---------------------------------------------------------------------------------
// persistent, column-chunked panel executor (the default for the wide path)
// ---------------------------------------------------------------------------------
// The level-by-level executor above streams the whole n_ext x k block through every
// sub-level: for n ~ 1e5 and k = 1024 the block (0.7 GB) is far larger than the L2, every x row
// is re-read from HBM dozens of times over the ~190 sub-levels and the solve runs at the speed
// of 256-byte random HBM reads (measured: 2.5 TFLOP/s).  Here the block is cut into CHUNKS of
// 32*T columns that stay resident in the 126 MB L2 for a whole pass over the program:
//   * the SMs are split into G GROUPS; a group takes one chunk through load -> all sub-levels
//     -> store, then the next chunk (chunk = group, group + G, ...); G chunks are in flight, sized
//     so that together they fit the L2;
//   * ONE launch per solve: the sub-level barrier is a counter/generation barrier among the
//     co-resident CTAs of a group (cooperative launch), not a kernel boundary;
//   * inside a sub-level the panels are dealt to the CTAs of the group; a warp owns a panel
//     (8 rows, register blocked) x the chunk's columns (lane = column).
struct GroupBar {
    unsigned int count, gen;
    unsigned int pad[30];
};

struct PersistArgs {
    SolveArgs a;
    PanelArgs p;
    const int32_t* sub_pan;      // device: first panel of every sub-level (nsub + 1), then longest list (nsub)
    int nsub, ngroups, cpg, nchunks;
    GroupBar* bars;
    int* err;
};

__device__ __forceinline__ bool group_barrier(GroupBar* bar, unsigned int ncta, unsigned int& gen, int* err) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                   // publish this CTA's x rows
        const unsigned int want = gen + 1;
        if (atomicAdd(&bar->count, 1u) == ncta - 1) {
            bar->count = 0;
            __threadfence();
            atomicExch(&bar->gen, want);
        } else {
            unsigned long long spins = 0;
            while (*((volatile unsigned int*)&bar->gen) != want) {
                if (++spins > (1ull << 31)) { atomicExch(err, 1); break; }   // watchdog: never hang the GPU
                __nanosleep(20);
            }
        }
        __threadfence();                                   // acquire: invalidates this SM's L1
    }
    ++gen;
    __syncthreads();
    return true;
}

template <int T, int NST>
__global__ void __launch_bounds__(256) panel_persist_kernel(const PersistArgs q) {
    extern __shared__ __align__(16) double psm[];   // per warp: NST stages (also the split-list reduce buffer)
    if (q.p.skip && *q.p.skip) return;
    constexpr int WD = NST * panel_stage_doubles<T>();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int group = blockIdx.x / q.cpg, gcta = blockIdx.x - group * q.cpg;
    if (group >= q.ngroups) return;
    GroupBar* bar = q.bars + group;
    unsigned int gen = 0;
    const SolveArgs& a = q.a;
    const int64_t ldx = q.p.ldx;
    constexpr int CW = 32 * T;
    double* ring = psm + (size_t)warp * WD;
    for (int chunk = group; chunk < q.nchunks; chunk += q.ngroups) {
        const int64_t c0 = (int64_t)chunk * CW;
        // ---- load: xe[perm_r[i], c0 + c] = B[i, c0 + c] (0 beyond nrows_b / k)
        for (int64_t e = (int64_t)gcta * 256 + tid; e < a.n * CW; e += (int64_t)q.cpg * 256) {
            const int64_t i = e / CW, c = c0 + (e - i * CW);
            double v = 0.0;
            if (i < a.nrows_b && c < a.k) v = a.B[i * a.ldb + c];
            q.p.xe[(int64_t)__ldg(a.perm_r + i) * ldx + c] = v;
        }
        group_barrier(bar, q.cpg, gen, q.err);
        double* xt = q.p.xe + c0;
        double* xc = xt + lane;
        for (int sb = 0; sb < q.nsub; ++sb) {
            const int p0 = __ldg(q.sub_pan + sb), p1 = __ldg(q.sub_pan + sb + 1);
            const int np = p1 - p0;
            if (np > 0) {
                const int maxcol = __ldg(q.sub_pan + q.nsub + 1 + sb);
                int wlog = 0;   // split long lists while the sub-level has fewer panels than the group has warps
                while (wlog < 3 && (np << wlog) < q.cpg * 8 && (maxcol >> (wlog + 1)) >= 32) ++wlog;
                const int wpr = 1 << wlog, ppc = 8 >> wlog;
                const int units = (np + ppc - 1) / ppc;
                const int wr = warp & (wpr - 1);
                for (int unit = gcta; unit < units; unit += q.cpg) {
                    const int pi = p0 + unit * ppc + (warp >> wlog);
                    const bool valid = pi < p1;
                    double acc[PANEL_ROWS][T];
#pragma unroll
                    for (int r = 0; r < PANEL_ROWS; ++r)
#pragma unroll
                        for (int t = 0; t < T; ++t) acc[r][t] = 0.0;
                    int4 pn0 = make_int4(0, 0, 0, -1);
                    int nrows = 0;
                    if (valid) {
                        pn0 = __ldg((const int4*)(q.p.panels + pi));
                        nrows = __ldg(&q.p.panels[pi].nrows);
                        const int ncol = pn0.y;
                        const int per = (((ncol + 7) >> 3) + wpr - 1) / wpr * 8;
                        const int e0 = wr * per, e1 = min(ncol, e0 + per);
                        panel_accumulate<T, NST, false>(acc, ring, q.p.col + pn0.x,
                                                        q.p.val + (int64_t)pn0.x * PANEL_ROWS, xt, ldx, e0, e1, lane);
                    }
                    if (wlog > 0) {
                        __syncthreads();
                        double* mine = ring;
#pragma unroll
                        for (int r = 0; r < PANEL_ROWS; ++r)
#pragma unroll
                            for (int t = 0; t < T; ++t) mine[(r * T + t) * 32 + lane] = acc[r][t];
                        __syncthreads();
                        if (wr == 0) {
                            for (int w2 = 1; w2 < wpr; ++w2) {
                                const double* other = psm + (size_t)(warp + w2) * WD;
#pragma unroll
                                for (int r = 0; r < PANEL_ROWS; ++r)
#pragma unroll
                                    for (int t = 0; t < T; ++t) acc[r][t] += other[(r * T + t) * 32 + lane];
                            }
                        }
                        __syncthreads();   // the ring is reused by the next unit
                    }
                    if (valid && wr == 0) {
                        const double* sc = q.p.scale + (int64_t)pi * PANEL_ROWS;
#pragma unroll
                        for (int r = 0; r < PANEL_ROWS; ++r) {
                            if (r < nrows) {
                                const double sv = __ldg(sc + r);
                                double* xd = xc + (int64_t)(pn0.z + r) * ldx;
                                if (pn0.w >= 0) {
                                    const double* xi = xc + (int64_t)(pn0.w + r) * ldx;
#pragma unroll
                                    for (int t = 0; t < T; ++t) xd[32 * t] = (__ldcg(xi + 32 * t) - acc[r][t]) * sv;
                                } else {
#pragma unroll
                                    for (int t = 0; t < T; ++t) xd[32 * t] = -acc[r][t] * sv;
                                }
                            }
                        }
                    }
                }
            }
            group_barrier(bar, q.cpg, gen, q.err);
        }
        // ---- store: X[j, c0 + c] = xe[perm_c[j], c0 + c]  (NaN if the watchdog fired)
        const bool bad = *((volatile int*)q.err) != 0;
        for (int64_t e = (int64_t)gcta * 256 + tid; e < a.nrows_x * CW; e += (int64_t)q.cpg * 256) {
            const int64_t j = e / CW, c = c0 + (e - j * CW);
            if (c < a.k) {
                const double v = __ldcg(q.p.xe + (int64_t)__ldg(a.perm_c + j) * ldx + c);
                a.X[j * a.ldx + c] = bad ? __longlong_as_double(0x7ff8000000000000LL) : v;
            }
        }
    }
}

constexpr int64_t PERSIST_TAIL_BYTES = 8192;   // group barriers + error flag behind the xe block

template <int T, int NST>
static int persist_launch(PersistArgs q, int ctas_per_sm_want, cudaStream_t st) {
    const size_t smem = (size_t)8 * NST * panel_stage_doubles<T>() * sizeof(double);
    static int per_sm = -1;
    if (per_sm < 0) {
        OCB_CUDA(cudaFuncSetAttribute(panel_persist_kernel<T, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, panel_persist_kernel<T, NST>, 256, smem));
    }
    if (per_sm < 1) {
        set_error("persistent solve: the kernel does not fit an SM");
        return OCB_ERR_CUDA;
    }
    // all CTAs must be co-resident (they wait for one another): as many per SM as fit, at most the wish
    q.cpg = std::max(1, std::min(ctas_per_sm_want, per_sm) * sm_count() / q.ngroups);
    void* args[] = {(void*)&q};
    OCB_CUDA(cudaLaunchCooperativeKernel((void*)panel_persist_kernel<T, NST>, dim3((unsigned)(q.ngroups * q.cpg)),
                                         dim3(256), args, smem, st));
    count_launch();
    return OCB_OK;
}

static int persist_ctas_per_sm() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("OCB_PERSIST_CTAS_PER_SM");
        v = e ? atoi(e) : 2;
        if (v < 1 || v > 4) v = 2;
    }
    return v;
}

// plan of a persistent solve: column tiles per warp (T), groups, chunks
struct PersistPlan { int T, ngroups, nchunks; int64_t ldx; };
static PersistPlan persist_plan(const ocb_lu* lu, int64_t k) {
    static int64_t l2_budget = 0;
    static int force_t = -1, force_g = -1;
    if (l2_budget == 0) {
        const char* e = getenv("OCB_PERSIST_L2_MB");
        l2_budget = (int64_t)(e ? atoi(e) : 80) << 20;
        const char* e2 = getenv("OCB_PANEL_T");
        force_t = e2 ? atoi(e2) : 0;
        const char* e3 = getenv("OCB_PERSIST_GROUPS");
        force_g = e3 ? atoi(e3) : 0;
    }
    PersistPlan pl;
    const int64_t bytes32 = lu->n_ext * 32 * 8;
    int T = 1;
    if (k > 32) {
        T = 2;
        if (k > 64 && 4 * bytes32 * 2 <= l2_budget) T = 4;
        while (T > 1 && T * bytes32 * 2 > l2_budget) T >>= 1;   // at least two chunks in the L2
    }
    if (force_t == 1 || force_t == 2 || force_t == 4) T = (k <= 32) ? 1 : force_t;
    pl.T = T;
    const int64_t cw = 32 * T;
    pl.nchunks = (int)((k + cw - 1) / cw);
    pl.ldx = (int64_t)pl.nchunks * cw;
    int64_t g = std::max<int64_t>(1, l2_budget / (T * bytes32));
    g = std::min<int64_t>(g, 8);
    g = std::min<int64_t>(g, pl.nchunks);
    if (force_g >= 1 && force_g <= 16) g = std::min<int64_t>(force_g, pl.nchunks);
    pl.ngroups = (int)g;
    return pl;
}

static int persist_solve(const ocb_lu* lu, const SolveArgs& a, cudaStream_t st) {
    const PersistPlan pl = persist_plan(lu, a.k);
    PersistArgs q;
    q.a = a;
    q.p.panels = lu->p_panels; q.p.scale = lu->p_scale; q.p.val = lu->p_val; q.p.col = lu->p_col;
    q.p.xe = a.ws; q.p.ldx = pl.ldx; q.p.skip = a.skip; q.p.ntile = 1;
    q.sub_pan = lu->p_sub_dev;
    q.nsub = (int)lu->sub_pan.size() - 1;
    q.ngroups = pl.ngroups;
    q.nchunks = pl.nchunks;
    q.cpg = 0;   // set by persist_launch from the occupancy of the instantiation
    unsigned char* tail = (unsigned char*)a.ws + lu->n_ext * pl.ldx * sizeof(double);
    tail = (unsigned char*)(((uintptr_t)tail + 255) & ~(uintptr_t)255);
    q.bars = (GroupBar*)tail;
    q.err = (int*)(tail + 16 * sizeof(GroupBar));
    OCB_CUDA(cudaMemsetAsync(tail, 0, 16 * sizeof(GroupBar) + 64, st));
    const int want = persist_ctas_per_sm();
    if (pl.T == 1) return persist_launch<1, 4>(q, want, st);
    if (pl.T == 2) return persist_launch<2, 3>(q, want, st);
    return persist_launch<4, 3>(q, want, st);
}

static int panel_solve(const ocb_lu* lu, const SolveArgs& a, cudaStream_t st) {
    const int T = panel_tiles(a.k);
    const int64_t ldx = panel_ldx(a.k);
    static int NST = 0;
    if (NST == 0) {
        const char* e = getenv("OCB_PANEL_STAGES");
        NST = e ? atoi(e) : 3;
        if (NST != 3 && NST != 4) NST = 3;
    }
    double* xe = a.ws;
    const unsigned lblocks = (unsigned)std::min<int64_t>((a.n * ldx + 255) / 256, 148 * 16);
    wide_load_kernel<<<lblocks, 256, 0, st>>>(a, xe, ldx);
    OCB_LAUNCH_CHECK();
    PanelArgs w;
    w.panels = lu->p_panels; w.scale = lu->p_scale; w.val = lu->p_val; w.col = lu->p_col;
    w.xe = xe; w.ldx = ldx; w.skip = a.skip;
    w.ntile = (int)(ldx / (32 * T));
    const int nsub = (int)lu->sub_pan.size() - 1;
    const int64_t want_warps = (int64_t)sm_count() * 16;
    for (int sb = 0; sb < nsub; ++sb) {
        const int p0 = lu->sub_pan[sb], p1 = lu->sub_pan[sb + 1];
        if (p1 <= p0) continue;
        // split long lists over the warps of a CTA while the sub-level has too few tasks to
        // fill the machine (the top separators: a few dozen panels with thousands of columns)
        const int64_t tasks = (int64_t)(p1 - p0) * w.ntile;
        const int maxcol = lu->sub_maxcol[sb];
        int wlog = 0;
        while (wlog < 3 && (tasks << wlog) < want_warps && (maxcol >> (wlog + 1)) >= 32) ++wlog;
        int rc;
        if (T == 1) rc = panel_launch<1, 4>(w, p0, p1, wlog, st);
        else if (T == 2) rc = NST == 4 ? panel_launch<2, 4>(w, p0, p1, wlog, st) : panel_launch<2, 3>(w, p0, p1, wlog, st);
        else rc = panel_launch<4, 3>(w, p0, p1, wlog, st);
        if (rc) return rc;
    }
    const unsigned sblocks = (unsigned)std::min<int64_t>((a.nrows_x * a.k + 255) / 256, 148 * 16);
    wide_store_kernel<<<std::max(1u, sblocks), 256, 0, st>>>(a, xe, ldx);
    OCB_LAUNCH_CHECK();
    return OCB_OK;
}

// optional per-launch timing of the solve kernel (bench.py roofline): CUDA events on the
// launching stream around every solve launch, summed on collect
struct SolveProf {
    bool on = false;
    std::vector<cudaEvent_t> ev;
    std::vector<double> bytes;   // algorithmic bytes of every timed launch
    size_t used = 0;
    double alg_bytes = 0.0;
    long long launches = 0;
};
static SolveProf g_prof;
static long long* g_trace_buf = nullptr;

static int lu_solve_dispatch(const ocb_lu* lu, const double* B, int64_t ldb, int64_t nrows_b, double* X,
                             int64_t ldx, int64_t nrows_x, int64_t k, void* ws, int64_t ws_bytes,
                             cudaStream_t st, const int* skip) {
    SolveArgs a;
    a.skip = skip;
    a.perm_r = lu->perm_r;
    a.perm_c = lu->perm_c;
    a.n = lu->n;
    a.n_ext = lu->n_ext;
    a.B = B;
    a.ldb = ldb;
    a.nrows_b = nrows_b;
    a.X = X;
    a.ldx = ldx;
    a.nrows_x = nrows_x;
    a.k = k;
    a.ws = (double*)ws;
    for (int r = 0; r < 8; ++r) {
        a.stream[r] = lu->stream[r];
        a.batch_off[r] = lu->batch_off[r];
        a.nbatch[r] = lu->nbatch[r];
    }
    a.stage_bytes = lu->stage_bytes;
    a.nstages = lu->nstages;
    {
        static int chunk = 0, dbg = -1;
        if (chunk == 0) {
            const char* e1 = getenv("OCB_TMA_CHUNK");
            chunk = e1 ? atoi(e1) : 65536;
            if (chunk < 16) chunk = 65536;
            chunk &= ~15;
            const char* e2 = getenv("OCB_TRSM_DEBUG_SKIP");
            dbg = e2 ? atoi(e2) : 0;
        }
        a.tma_chunk = chunk;
        a.debug_skip = dbg;
        static long long* trace_buf = nullptr;
        static int trace_on = -1;
        if (trace_on < 0) {
            const char* e3 = getenv("OCB_TRSM_TRACE");
            trace_on = (e3 && e3[0] == '1') ? 1 : 0;
            if (trace_on) cudaMalloc((void**)&trace_buf, 4096 * sizeof(long long));
        }
        a.trace = trace_on ? trace_buf : nullptr;
        g_trace_buf = trace_buf;
    }
    if (use_wide(lu, k)) {
        const int64_t need = ocb_lu_solve_ws_bytes(lu, k);
        if (ws == nullptr || ws_bytes < need) {
            set_error("lu_solve: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)need);
            return OCB_ERR_CAPACITY;
        }
        // which all-columns executor: measured on the cavity factors (profiles/r02_wide_executor.md) the
        // register-blocked panel kernel wins for wide blocks on large factors, the row kernel elsewhere
        static int force = -1;   // OCB_WIDE_EXECUTOR = rows | panels | persist
        static int64_t min_k = 96, min_n = 0;
        if (force < 0) {
            const char* e = getenv("OCB_WIDE_EXECUTOR");
            force = !e ? 0 : (e[0] == 'r' ? 1 : (e[0] == 'p' && e[1] == 'a' ? 2 : (e[0] == 'p' ? 3 : 0)));
            if (getenv("OCB_WIDE_BY_LEVEL")) force = 2;
            const char* ek = getenv("OCB_PANEL_MIN_K");
            const char* en = getenv("OCB_PANEL_MIN_N");
            if (ek) min_k = atoi(ek);
            if (en) min_n = atoi(en);
        }
        if (lu->has_panels && force == 3) return persist_solve(lu, a, st);
        if (lu->has_panels && (force == 2 || (force == 0 && k >= min_k && lu->n_ext >= min_n) || !lu->has_flat))
            return panel_solve(lu, a, st);
        return wide_solve(lu, a, st);
    }
    switch (lu->cl) {
        case 8: return launch_cluster<8>(lu, a, st);
        case 3: return launch_cluster<3>(lu, a, st);
        case 4: return launch_cluster<4>(lu, a, st);
        case 2: return launch_cluster<2>(lu, a, st);
        default: return launch_cluster<1>(lu, a, st);
    }
}

int lu_solve_impl(const ocb_lu* lu, const double* B, int64_t ldb, int64_t nrows_b, double* X,
                  int64_t ldx, int64_t nrows_x, int64_t k, void* ws, int64_t ws_bytes,
                  cudaStream_t st, const int* skip) {
    if (k == 0 || lu->n == 0) return OCB_OK;
    if (!g_prof.on) return lu_solve_dispatch(lu, B, ldb, nrows_b, X, ldx, nrows_x, k, ws, ws_bytes, st, skip);
    if (g_prof.used + 2 > g_prof.ev.size()) {
        const size_t old = g_prof.ev.size();
        g_prof.ev.resize(old + 1024);
        for (size_t i = old; i < g_prof.ev.size(); ++i) OCB_CUDA(cudaEventCreate(&g_prof.ev[i]));
    }
    OCB_CUDA(cudaEventRecord(g_prof.ev[g_prof.used], st));
    const int rc = lu_solve_dispatch(lu, B, ldb, nrows_b, X, ldx, nrows_x, k, ws, ws_bytes, st, skip);
    OCB_CUDA(cudaEventRecord(g_prof.ev[g_prof.used + 1], st));
    g_prof.used += 2;
    g_prof.launches += 1;
    const double ab = 12.0 * (double)(lu->nnzL + lu->nnzU) + 16.0 * (double)(lu->n + 1) +
                      32.0 * (double)lu->n * (double)k;
    g_prof.alg_bytes += ab;
    g_prof.bytes.push_back(ab);
    return rc;
}

// the last n timed launches were no-ops (their skip flag was set: the ADI loop had converged
// before they ran): they did no work, so they leave the statistics again
void lu_solve_prof_discard_last(int64_t n) {
    if (!g_prof.on) return;
    for (; n > 0 && g_prof.launches > 0 && !g_prof.bytes.empty(); --n) {
        g_prof.alg_bytes -= g_prof.bytes.back();
        g_prof.bytes.pop_back();
        g_prof.used -= 2;
        g_prof.launches -= 1;
    }
}

}  // namespace ocb

extern "C" {

// Common body of the three packing entry points.  dst != null: build there (OCB_ERR_CAPACITY if
// it does not fit); otherwise *img_out is malloc'ed.  A_* != null: residual guard - the program
// is executed on the host for one pseudo-random right-hand side and the normwise backward
// error ||b - A x|| / (||A||_F ||x|| + ||b||) of the ORIGINAL matrix (CSC) is returned.
static int pack_common(int64_t n, const int32_t* Lrp, const int32_t* Lci, const double* Lva,
                       const int32_t* Urp, const int32_t* Uci, const double* Uva, const int32_t* perm_r,
                       const int32_t* perm_c, int64_t max_smem_optin, int64_t flags, unsigned char* dst,
                       int64_t dst_capacity, unsigned char** img_out, int64_t* out_bytes,
                       const int32_t* A_colptr, const int32_t* A_rowidx, const double* A_vals,
                       double* out_backerr) {
    static thread_local ocb::LuProgram P;     // kept between calls: see the template hit in build_lu_program
    int rc = ocb::build_lu_program(n, Lrp, Lci, Lva, Urp, Uci, Uva, ocb::trsm_threads(), (flags & 2) != 0,
                                   (flags & 4) != 0 && !(flags & 8), &P, (flags & 8) ? 32 : 0);
    if (rc != OCB_OK) return rc;
    if (out_backerr) *out_backerr = 0.0;
    if (A_colptr && A_rowidx && A_vals && out_backerr && n > 0) {
        std::vector<double> b((size_t)n), x((size_t)n), r;
        uint64_t sd = 0x9e3779b97f4a7c15ULL;   // fixed pseudo-random right-hand side in [-1, 1)
        for (int64_t i = 0; i < n; ++i) {
            sd = sd * 6364136223846793005ULL + 1442695040888963407ULL;
            b[i] = (double)(int64_t)(sd >> 11) / 4503599627370496.0 - 1.0;
        }
        ocb::execute_program_host(P, perm_r, perm_c, b.data(), x.data());
        r = b;
        double a2 = 0.0, x2 = 0.0, b2 = 0.0, r2 = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            const double xj = x[j];
            x2 += xj * xj;
            for (int32_t p = A_colptr[j]; p < A_colptr[j + 1]; ++p) {
                r[A_rowidx[p]] -= A_vals[p] * xj;
                a2 += A_vals[p] * A_vals[p];
            }
        }
        for (int64_t i = 0; i < n; ++i) { r2 += r[i] * r[i]; b2 += b[i] * b[i]; }
        const double den = sqrt(a2) * sqrt(x2) + sqrt(b2);
        *out_backerr = (r2 == r2 && den > 0.0 && den == den) ? sqrt(r2) / den : 1.0;   // NaN -> 1
    }
    const auto t0 = std::chrono::steady_clock::now();
    unsigned char* img = nullptr;
    rc = ocb::pack_image(P, perm_r, perm_c, (int)max_smem_optin, (int)flags, &img, out_bytes, dst, dst_capacity);
    if (getenv("OCB_TIMING"))
        fprintf(stderr, "lu_pack_host: image packed in %.1f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    if (rc != OCB_OK) return rc;
    if (dst && img != dst) {   // did not fit: *out_bytes tells how much room the image needs
        free(img);
        ocb::set_error("lu_pack_host_into: the image needs %lld bytes, the buffer has %lld",
                       (long long)*out_bytes, (long long)dst_capacity);
        return OCB_ERR_CAPACITY;
    }
    if (img_out) *img_out = img;
    return OCB_OK;
}

int ocb_lu_pack_host(int64_t n, const int32_t* h_L_rowptr, const int32_t* h_L_colidx, const double* h_L_vals,
                     const int32_t* h_U_rowptr, const int32_t* h_U_colidx, const double* h_U_vals,
                     const int32_t* h_perm_r, const int32_t* h_perm_c, int64_t max_smem_optin, int64_t flags,
                     unsigned char** out_image, int64_t* out_bytes) {
    OCB_ARG(n >= 0 && out_image && out_bytes, "lu_pack_host");
    OCB_ARG(h_L_rowptr && h_U_rowptr && h_perm_r && h_perm_c, "lu_pack_host: null pointer");
    OCB_ARG(max_smem_optin >= 48 * 1024, "lu_pack_host: shared-memory size");
    return pack_common(n, h_L_rowptr, h_L_colidx, h_L_vals, h_U_rowptr, h_U_colidx, h_U_vals, h_perm_r,
                       h_perm_c, max_smem_optin, flags, nullptr, 0, out_image, out_bytes, nullptr, nullptr,
                       nullptr, nullptr);
}

int ocb_lu_pack_host_into(int64_t n, const int32_t* h_L_rowptr, const int32_t* h_L_colidx, const double* h_L_vals,
                          const int32_t* h_U_rowptr, const int32_t* h_U_colidx, const double* h_U_vals,
                          const int32_t* h_perm_r, const int32_t* h_perm_c, int64_t max_smem_optin, int64_t flags,
                          unsigned char* dst, int64_t dst_capacity, int64_t* out_bytes) {
    OCB_ARG(n >= 0 && dst && dst_capacity >= 0 && out_bytes, "lu_pack_host_into");
    OCB_ARG(h_L_rowptr && h_U_rowptr && h_perm_r && h_perm_c, "lu_pack_host_into: null pointer");
    OCB_ARG(max_smem_optin >= 48 * 1024, "lu_pack_host_into: shared-memory size");
    return pack_common(n, h_L_rowptr, h_L_colidx, h_L_vals, h_U_rowptr, h_U_colidx, h_U_vals, h_perm_r,
                       h_perm_c, max_smem_optin, flags, dst, dst_capacity, nullptr, out_bytes, nullptr, nullptr,
                       nullptr, nullptr);
}

int ocb_lu_pack_host_checked(int64_t n, const int32_t* h_L_rowptr, const int32_t* h_L_colidx,
                             const double* h_L_vals, const int32_t* h_U_rowptr, const int32_t* h_U_colidx,
                             const double* h_U_vals, const int32_t* h_perm_r, const int32_t* h_perm_c,
                             int64_t max_smem_optin, int64_t flags, unsigned char* dst, int64_t dst_capacity,
                             unsigned char** out_image, int64_t* out_bytes, const int32_t* h_A_colptr,
                             const int32_t* h_A_rowidx, const double* h_A_vals, double* out_backerr) {
    OCB_ARG(n >= 0 && out_bytes && (dst || out_image), "lu_pack_host_checked");
    OCB_ARG(h_L_rowptr && h_U_rowptr && h_perm_r && h_perm_c, "lu_pack_host_checked: null pointer");
    OCB_ARG(max_smem_optin >= 48 * 1024, "lu_pack_host_checked: shared-memory size");
    OCB_ARG(h_A_colptr && h_A_rowidx && h_A_vals && out_backerr, "lu_pack_host_checked: matrix");
    return pack_common(n, h_L_rowptr, h_L_colidx, h_L_vals, h_U_rowptr, h_U_colidx, h_U_vals, h_perm_r,
                       h_perm_c, max_smem_optin, flags, dst, dst_capacity, out_image, out_bytes, h_A_colptr,
                       h_A_rowidx, h_A_vals, out_backerr);
}

void ocb_host_free(void* p) { free(p); }

int ocb_lu_create_from_image(ocb_lu** out, const unsigned char* h_image, int64_t bytes, void* d_arena,
                             void* stream) {
    OCB_ARG(out && h_image && bytes > 0, "lu_create_from_image");
    ocb_lu* lu = new ocb_lu();
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&lu->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int rc = ocb::upload_image(lu, h_image, bytes, d_arena, (cudaStream_t)stream);
    if (rc != OCB_OK) {
        ocb_lu_destroy(lu);
        return rc;
    }
    *out = lu;
    return OCB_OK;
}

int ocb_lu_create(ocb_lu** out, int64_t n, const int32_t* h_L_rowptr, const int32_t* h_L_colidx,
                  const double* h_L_vals, const int32_t* h_U_rowptr, const int32_t* h_U_colidx,
                  const double* h_U_vals, const int32_t* h_perm_r, const int32_t* h_perm_c,
                  void* stream) {
    OCB_ARG(out && n >= 0, "lu_create");
    OCB_ARG(h_L_rowptr && h_U_rowptr && h_perm_r && h_perm_c, "lu_create: null pointer");
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (optin <= 0) {
        ocb::set_error("lu_create: no CUDA device");
        return OCB_ERR_CUDA;
    }
    unsigned char* img = nullptr;
    int64_t bytes = 0;
    int rc = ocb_lu_pack_host(n, h_L_rowptr, h_L_colidx, h_L_vals, h_U_rowptr, h_U_colidx, h_U_vals, h_perm_r,
                              h_perm_c, optin, 0, &img, &bytes);
    if (rc != OCB_OK) return rc;
    rc = ocb_lu_create_from_image(out, img, bytes, nullptr, stream);
    free(img);
    return rc;
}

int ocb_lu_destroy(ocb_lu* lu) {
    if (!lu) return OCB_OK;
    if (lu->arena_owned) cudaFree(lu->arena);
    delete lu;
    return OCB_OK;
}

int ocb_prof_enable(int on) {
    ocb::g_prof.on = on != 0;
    ocb::g_prof.used = 0;
    ocb::g_prof.alg_bytes = 0.0;
    ocb::g_prof.launches = 0;
    ocb::g_prof.bytes.clear();
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
    info8[1] = lu->nnzL;
    info8[2] = lu->nnzU;
    info8[3] = lu->nsub_L;
    info8[4] = lu->nsub_U;
    info8[5] = lu->bytes;
    info8[6] = lu->kp_smem_max;  // 0: panel in a global slab
    info8[7] = lu->nbatch[0];
    return OCB_OK;
}

int ocb_debug_trace(int64_t* h_out, int64_t count) {
    OCB_ARG(h_out && count > 0 && count <= 4096, "debug_trace");
    if (!ocb::g_trace_buf) {
        ocb::set_error("debug_trace: tracing is off (set OCB_TRSM_TRACE=1 before the first solve)");
        return OCB_ERR_ARG;
    }
    OCB_CUDA(cudaMemcpy(h_out, ocb::g_trace_buf, (size_t)count * sizeof(long long), cudaMemcpyDeviceToHost));
    return OCB_OK;
}

int64_t ocb_lu_panel_levels(const ocb_lu* lu, int64_t* h_out3, int64_t capacity_levels) {
    // debugging / profiling aid: per sub-level {panels, longest column list, -} of the panel program
    if (!lu || !lu->has_panels) return 0;
    const int64_t nsub = (int64_t)lu->sub_pan.size() - 1;
    if (h_out3)
        for (int64_t sb = 0; sb < nsub && sb < capacity_levels; ++sb) {
            h_out3[3 * sb] = lu->sub_pan[sb + 1] - lu->sub_pan[sb];
            h_out3[3 * sb + 1] = lu->sub_maxcol[sb];
            h_out3[3 * sb + 2] = 0;
        }
    return nsub;
}

int ocb_lu_stats(const ocb_lu* lu, int64_t* info8) {
    OCB_ARG(lu && info8, "lu_stats");
    info8[0] = lu->n_ext;
    info8[1] = lu->nsuper;
    info8[2] = lu->max_w;
    info8[3] = lu->nseg;
    info8[4] = lu->nrows;
    info8[5] = lu->nent;
    info8[6] = lu->stage_bytes;
    info8[7] = lu->nstages;
    return OCB_OK;
}

int64_t ocb_lu_solve_ws_bytes(const ocb_lu* lu, int64_t k) {
    if (!lu || !ocb::use_wide(lu, k)) return 0;
    int64_t ldx = std::max(ocb::wide_ldx(k), ocb::panel_ldx(k));
    if (lu->has_panels) ldx = std::max(ldx, ocb::persist_plan(lu, k).ldx);
    return lu->n_ext * ldx * (int64_t)sizeof(double) + ocb::PERSIST_TAIL_BYTES;
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
                              (cudaStream_t)stream, nullptr);
}
}
