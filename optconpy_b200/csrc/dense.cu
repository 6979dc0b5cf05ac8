// K3 / K4: the dense contractions of the hot path, on the FP64 tensor pipe (DMMA.8x8x4),
// and the small on-device symmetric eigen-solver.
//
//   gram       G = Z^T W            flops 2*n*ka*kb, bytes 8*n*(ka+kb) + 8*ka*kb
//   tall_gemm  C = Z T              flops 2*n*k*kc,  bytes 8*n*(k+kc)
//   sym_eig    parallel cyclic two-sided Jacobi, cooperative launch (grid.sync per round)
//   small_inv  Gauss-Jordan with partial pivoting for the m x m SMW core (m <= 32)
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <algorithm>

namespace cg = cooperative_groups;

namespace ocb {

// ---------------------------------------------------------------------------------
// Gram: CTA tile 64 x 64 of G, 8 warps (each 16 x 32 = 2 x 4 DMMA tiles), rows split
// over blockIdx.z; partial tiles go to a workspace and are summed in fixed order.
// ---------------------------------------------------------------------------------
constexpr int GR_TM = 64, GR_TN = 64, GR_RK = 32, GR_LD = 72;  // 72 % 16 == 8: conflict-free frags

__global__ void __launch_bounds__(256) gram_partial_kernel(const double* __restrict__ Z, int64_t ldz,
                                                          int64_t ka, const double* __restrict__ W,
                                                          int64_t ldw, int64_t kb, int64_t n,
                                                          int64_t rows_per_split,
                                                          double* __restrict__ P /*[split][ka][kb]*/,
                                                          int sym) {
    __shared__ double Zs[GR_RK * GR_LD];
    __shared__ double Ws[GR_RK * GR_LD];
    // sym: W == Z, G is symmetric: only the tiles on and below the diagonal are computed
    if (sym && blockIdx.x < blockIdx.y) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t a0 = (int64_t)blockIdx.x * GR_TM, b0 = (int64_t)blockIdx.y * GR_TN;
    const int64_t r_beg = (int64_t)blockIdx.z * rows_per_split;
    const int64_t r_end = min(n, r_beg + rows_per_split);
    const int m0 = (warp & 3) * 16, n0 = (warp >> 2) * 32;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int64_t r0 = r_beg; r0 < r_end; r0 += GR_RK) {
        // stage 32 rows x 64 cols of each operand (coalesced 512-byte row segments)
        for (int e = tid; e < GR_RK * GR_TM; e += 256) {
            const int rr = e >> 6, cc = e & 63;
            const int64_t r = r0 + rr;
            double zv = 0.0, wv = 0.0;
            if (r < r_end) {
                if (a0 + cc < ka) zv = __ldg(Z + r * ldz + a0 + cc);
                if (b0 + cc < kb) wv = __ldg(W + r * ldw + b0 + cc);
            }
            Zs[rr * GR_LD + cc] = zv;
            Ws[rr * GR_LD + cc] = wv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GR_RK / 4; ++kk) {
            const int kr = kk * 4 + (lane & 3);
            double af[2], bf[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) af[i] = Zs[kr * GR_LD + m0 + i * 8 + (lane >> 2)];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = Ws[kr * GR_LD + n0 + j * 8 + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
    double* Pz = P + (int64_t)blockIdx.z * ka * kb;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t gr = a0 + m0 + i * 8 + (lane >> 2);
            const int64_t gc = b0 + n0 + j * 8 + 2 * (lane & 3);
            if (gr < ka) {
                if (gc < kb) Pz[gr * kb + gc] = acc[i][j][0];
                if (gc + 1 < kb) Pz[gr * kb + gc + 1] = acc[i][j][1];
            }
        }
}

__global__ void __launch_bounds__(256) gram_reduce_kernel(const double* __restrict__ P, int64_t ka,
                                                         int64_t kb, int nsplit,
                                                         double* __restrict__ G, int64_t ldg, int sym,
                                                         int tile) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ka * kb) return;
    int64_t r = e / kb, c = e % kb;
    int64_t src = e;
    if (sym && (r / tile) < (c / tile)) src = c * kb + r;   // tile above the diagonal: mirror
    double s = 0.0;
    for (int z = 0; z < nsplit; ++z) s += P[(int64_t)z * ka * kb + src];
    G[r * ldg + c] = s;
}

// ---------------------------------------------------------------------------------
// Large Gram products (ka, kb >= 128: the K x K matrix of the compression): CTA tile 128 x 128,
// 8 warps of 32 x 64 (4 x 8 DMMA tiles, 64 accumulator registers), row slabs of 16 staged by
// cp.async through a 3-stage shared-memory ring - the global loads of the next two slabs are in
// flight while the tensor pipe works on the current one.  Leading dimension 132 (= 4 mod 16): the
// 4 rows x 4 columns a half-warp reads for one fragment fall into 32 distinct banks.
// ---------------------------------------------------------------------------------
constexpr int G2_T = 128, G2_RK = 16, G2_LD = 132, G2_ST = 3;

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(256) gram_partial_kernel2(const double* __restrict__ Z, int64_t ldz,
                                                           int64_t ka, const double* __restrict__ W,
                                                           int64_t ldw, int64_t kb, int64_t n,
                                                           int64_t rows_per_split,
                                                           double* __restrict__ P /*[split][ka][kb]*/,
                                                           int sym) {
    extern __shared__ __align__(16) double g2sm[];   // [stage][Z | W][G2_RK][G2_LD]
    if (sym && blockIdx.x < blockIdx.y) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t a0 = (int64_t)blockIdx.x * G2_T, b0 = (int64_t)blockIdx.y * G2_T;
    const int64_t r_beg = (int64_t)blockIdx.z * rows_per_split;
    const int64_t r_end = min(n, r_beg + rows_per_split);
    const int m0 = (warp & 3) * 32, n0 = (warp >> 2) * 64;
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int nslab = (int)((r_end - r_beg + G2_RK - 1) / G2_RK);
    auto issue = [&](int sl) {
        double* zs = g2sm + (size_t)(sl % G2_ST) * (2 * G2_RK * G2_LD);
        double* ws = zs + G2_RK * G2_LD;
        const int64_t r0 = r_beg + (int64_t)sl * G2_RK;
        // 16 rows x 128 columns per operand = 1024 chunks of 16 bytes: 4 per thread and operand
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int ch = tid + q * 256;
            const int rr = ch >> 6, cc = (ch & 63) * 2;
            const int64_t r = r0 + rr;
            const bool rok = r < r_end;
            const int za = rok ? (int)max((int64_t)0, min((int64_t)16, (ka - (a0 + cc)) * 8)) : 0;
            const int wa = rok ? (int)max((int64_t)0, min((int64_t)16, (kb - (b0 + cc)) * 8)) : 0;
            // clamp the source address into the block (it is not read when the size is 0)
            const double* zsrc = Z + (rok ? r : r_beg) * ldz + (za > 0 ? a0 + cc : 0);
            const double* wsrc = W + (rok ? r : r_beg) * ldw + (wa > 0 ? b0 + cc : 0);
            cp_async16_zfill(zs + rr * G2_LD + cc, zsrc, za);
            cp_async16_zfill(ws + rr * G2_LD + cc, wsrc, wa);
        }
    };
#pragma unroll
    for (int sl = 0; sl < G2_ST - 1; ++sl) {
        if (sl < nslab) issue(sl);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int sl = 0; sl < nslab; ++sl) {
        if (sl + G2_ST - 1 < nslab) issue(sl + G2_ST - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(G2_ST - 1) : "memory");
        __syncthreads();
        const double* zs = g2sm + (size_t)(sl % G2_ST) * (2 * G2_RK * G2_LD);
        const double* ws = zs + G2_RK * G2_LD;
#pragma unroll
        for (int kk = 0; kk < G2_RK / 4; ++kk) {
            const int kr = kk * 4 + (lane & 3);
            double af[4], bf[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = zs[kr * G2_LD + m0 + i * 8 + (lane >> 2)];
#pragma unroll
            for (int j = 0; j < 8; ++j) bf[j] = ws[kr * G2_LD + n0 + j * 8 + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();   // the slot is refilled by the copies issued next
    }
    double* Pz = P + (int64_t)blockIdx.z * ka * kb;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t gr = a0 + m0 + i * 8 + (lane >> 2);
            const int64_t gc = b0 + n0 + j * 8 + 2 * (lane & 3);
            if (gr < ka) {
                if (gc < kb) Pz[gr * kb + gc] = acc[i][j][0];
                if (gc + 1 < kb) Pz[gr * kb + gc + 1] = acc[i][j][1];
            }
        }
}

static bool gram_big_enabled() {
    static const bool on = getenv("OCB_GRAM_BIG_TILES") != nullptr;
    return on;
}

static bool gram_big(const double* Z, int64_t ldz, int64_t ka, const double* W, int64_t ldw, int64_t kb) {
    // opt-in (OCB_GRAM_BIG_TILES=1): measured equal to the 64 x 64 kernel (12.9 / 17.5 / 20.3 against
    // 12.5 / 17.3 / 19.7 TFLOP/s at K = 1296 / 1716 / 3072) - the Gram product is not bound by its staging
    static const bool on = getenv("OCB_GRAM_BIG_TILES") != nullptr;
    return on && ka >= 128 && kb >= 128 && (ldz & 1) == 0 && (ldw & 1) == 0 &&
           ((uintptr_t)Z & 15) == 0 && ((uintptr_t)W & 15) == 0;
}

static void gram_plan2(int64_t n, int64_t ka, int64_t kb, bool sym, int* nsplit, int64_t* rps) {
    const int64_t ta = (ka + G2_T - 1) / G2_T, tb = (kb + G2_T - 1) / G2_T;
    const int64_t tiles = sym ? ta * (ta + 1) / 2 : ta * tb;
    int64_t want = (2 * (int64_t)148 + tiles - 1) / tiles;          // ~2 CTAs per SM in total
    const int64_t maxsplit = (n + 8 * G2_RK - 1) / (8 * G2_RK);     // >= 128 rows per split
    want = std::max<int64_t>(1, std::min(want, std::max<int64_t>(1, maxsplit)));
    int64_t r = (n + want - 1) / want;
    r = align_up(std::max<int64_t>(r, 1), G2_RK);
    *rps = r;
    *nsplit = (int)std::max<int64_t>(1, (n + r - 1) / r);
}

static void gram_plan(int64_t n, int64_t ka, int64_t kb, int* nsplit, int64_t* rps) {
    const int64_t tiles = ((ka + GR_TM - 1) / GR_TM) * ((kb + GR_TN - 1) / GR_TN);
    int64_t want = (4 * (int64_t)148 + tiles - 1) / tiles;        // ~4 CTAs per SM in total
    const int64_t maxsplit = (n + 4 * GR_RK - 1) / (4 * GR_RK);   // >= 128 rows per split
    want = std::max<int64_t>(1, std::min(want, std::max<int64_t>(1, maxsplit)));
    int64_t r = (n + want - 1) / want;
    r = align_up(std::max<int64_t>(r, 1), GR_RK);
    *rps = r;
    *nsplit = (int)std::max<int64_t>(1, (n + r - 1) / r);
}

int gram_impl(const double* Z, int64_t ldz, int64_t ka, const double* W, int64_t ldw, int64_t kb,
              int64_t n, double* G, int64_t ldg, void* ws, int64_t ws_bytes, cudaStream_t st) {
    if (ka == 0 || kb == 0) return OCB_OK;
    const int sym = (Z == W && ldz == ldw && ka == kb && ka > GR_TM) ? 1 : 0;   // G = Z^T Z
    int nsplit;
    int64_t rps;
    const bool big = gram_big(Z, ldz, ka, W, ldw, kb);
    if (big) gram_plan2(n, ka, kb, sym != 0, &nsplit, &rps);
    else gram_plan(n, ka, kb, &nsplit, &rps);
    const int64_t need = (int64_t)nsplit * ka * kb * 8;
    if (!ws || ws_bytes < need) {
        set_error("gram: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)need);
        return OCB_ERR_CAPACITY;
    }
    if (big) {
        const int smem = G2_ST * 2 * G2_RK * G2_LD * (int)sizeof(double);
        static bool attr = false;
        if (!attr) {
            OCB_CUDA(cudaFuncSetAttribute(gram_partial_kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            attr = true;
        }
        dim3 grid2((unsigned)((ka + G2_T - 1) / G2_T), (unsigned)((kb + G2_T - 1) / G2_T), nsplit);
        gram_partial_kernel2<<<grid2, 256, smem, st>>>(Z, ldz, ka, W, ldw, kb, n, rps, (double*)ws, sym);
        OCB_LAUNCH_CHECK();
        gram_reduce_kernel<<<(unsigned)((ka * kb + 255) / 256), 256, 0, st>>>((const double*)ws, ka, kb,
                                                                             nsplit, G, ldg, sym, G2_T);
        OCB_LAUNCH_CHECK();
        return OCB_OK;
    }
    dim3 grid((unsigned)((ka + GR_TM - 1) / GR_TM), (unsigned)((kb + GR_TN - 1) / GR_TN), nsplit);
    gram_partial_kernel<<<grid, 256, 0, st>>>(Z, ldz, ka, W, ldw, kb, n, rps, (double*)ws, sym);
    OCB_LAUNCH_CHECK();
    gram_reduce_kernel<<<(unsigned)((ka * kb + 255) / 256), 256, 0, st>>>((const double*)ws, ka, kb,
                                                                         nsplit, G, ldg, sym, GR_TM);
    OCB_LAUNCH_CHECK();
    return OCB_OK;
}

// ---------------------------------------------------------------------------------
// tall_gemm: C (n x kc) = alpha Z (n x k) T (k x kc) + beta C.
// CTA tile 64 rows x 64 cols, 8 warps (16 x 32 each), inner dimension in chunks of 32.
// ---------------------------------------------------------------------------------
constexpr int TG_LDZ = 36;  // 36 % 16 == 4

__global__ void __launch_bounds__(256) tall_gemm_kernel(const double* __restrict__ Z, int64_t ldz,
                                                       int64_t n, int64_t k,
                                                       const double* __restrict__ T, int64_t ldt,
                                                       int64_t kc, double* __restrict__ C,
                                                       int64_t ldc, double alpha, double beta) {
    __shared__ double Zs[64 * TG_LDZ];
    __shared__ double Ts[32 * GR_LD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * 64, c0 = (int64_t)blockIdx.y * 64;
    const int m0 = (warp & 3) * 16, n0 = (warp >> 2) * 32;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int64_t k0 = 0; k0 < k; k0 += 32) {
        for (int e = tid; e < 64 * 32; e += 256) {
            const int rr = e >> 5, cc = e & 31;
            double v = 0.0;
            if (r0 + rr < n && k0 + cc < k) v = __ldg(Z + (r0 + rr) * ldz + k0 + cc);
            Zs[rr * TG_LDZ + cc] = v;
        }
        for (int e = tid; e < 32 * 64; e += 256) {
            const int rr = e >> 6, cc = e & 63;
            double v = 0.0;
            if (k0 + rr < k && c0 + cc < kc) v = __ldg(T + (k0 + rr) * ldt + c0 + cc);
            Ts[rr * GR_LD + cc] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            double af[2], bf[4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
                af[i] = Zs[(m0 + i * 8 + (lane >> 2)) * TG_LDZ + kk * 4 + (lane & 3)];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                bf[j] = Ts[(kk * 4 + (lane & 3)) * GR_LD + n0 + j * 8 + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t gr = r0 + m0 + i * 8 + (lane >> 2);
            const int64_t gc = c0 + n0 + j * 8 + 2 * (lane & 3);
            if (gr < n) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (gc + u < kc) {
                        double r = alpha * acc[i][j][u];
                        if (beta != 0.0) r = fma(beta, C[gr * ldc + gc + u], r);
                        C[gr * ldc + gc + u] = r;
                    }
                }
            }
        }
}

int tall_gemm_impl(const double* Z, int64_t ldz, int64_t n, int64_t k, const double* T, int64_t ldt,
                   int64_t kc, double* C, int64_t ldc, double alpha, double beta, cudaStream_t st) {
    if (n == 0 || kc == 0) return OCB_OK;
    dim3 grid((unsigned)((n + 63) / 64), (unsigned)((kc + 63) / 64));
    tall_gemm_kernel<<<grid, 256, 0, st>>>(Z, ldz, n, k, T, ldt, kc, C, ldc, alpha, beta);
    OCB_LAUNCH_CHECK();
    return OCB_OK;
}

// ---------------------------------------------------------------------------------
// Parallel cyclic Jacobi (two-sided), round-robin pairing, one cooperative kernel.
// ---------------------------------------------------------------------------------
constexpr int EIG_MAXK = 2048;
__device__ double g_eig_c[EIG_MAXK / 2], g_eig_s[EIG_MAXK / 2];
__device__ int g_eig_p[EIG_MAXK / 2], g_eig_q[EIG_MAXK / 2];
__device__ double g_eig_part[2 * 1024];
__device__ int g_eig_rank[EIG_MAXK];
__device__ int g_eig_sweeps;

__global__ void __launch_bounds__(256) jacobi_kernel(double* __restrict__ G, int64_t ldg, int k,
                                                    double* __restrict__ lam,
                                                    double* __restrict__ V, int64_t ldv,
                                                    int max_sweeps, double tol2) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double red[2][8];
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsize = (int64_t)gridDim.x * blockDim.x;
    const int kp = (k + 1) & ~1;  // padded to even; index k (if odd) is a bye
    const int np = kp / 2;
    // V = I
    for (int64_t e = gtid; e < (int64_t)k * k; e += gsize) {
        const int r = (int)(e / k), c = (int)(e % k);
        V[(int64_t)r * ldv + c] = (r == c) ? 1.0 : 0.0;
    }
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        // ---- convergence measure: off(G)^2 vs diag(G)^2, deterministic two-stage sum
        double off = 0.0, dg = 0.0;
        for (int64_t e = gtid; e < (int64_t)k * k; e += gsize) {
            const int r = (int)(e / k), c = (int)(e % k);
            const double v = G[(int64_t)r * ldg + c];
            if (r == c) dg = fma(v, v, dg); else off = fma(v, v, off);
        }
        for (int o = 16; o > 0; o >>= 1) {
            off += __shfl_xor_sync(0xffffffffu, off, o);
            dg += __shfl_xor_sync(0xffffffffu, dg, o);
        }
        if ((threadIdx.x & 31) == 0) {
            red[0][threadIdx.x >> 5] = off;
            red[1][threadIdx.x >> 5] = dg;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0, b = 0.0;
            for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
            g_eig_part[2 * blockIdx.x] = a;
            g_eig_part[2 * blockIdx.x + 1] = b;
        }
        grid.sync();
        double toff = 0.0, tdg = 0.0;
        for (int b = 0; b < (int)gridDim.x; ++b) {
            toff += g_eig_part[2 * b];
            tdg += g_eig_part[2 * b + 1];
        }
        if (toff <= tol2 * tdg) break;  // identical decision in every thread
        for (int round = 0; round < kp - 1; ++round) {
            // ---- phase A: rotations of this round's disjoint pairs
            for (int64_t t = gtid; t < np; t += gsize) {
                int p, q;
                if (t == 0) { p = kp - 1; q = round; }
                else {
                    p = (round + (int)t) % (kp - 1);
                    q = (round - (int)t + (kp - 1)) % (kp - 1);
                }
                if (p > q) { const int tmp = p; p = q; q = tmp; }
                double c = 1.0, s = 0.0;
                if (q < k) {
                    const double app = G[(int64_t)p * ldg + p], aqq = G[(int64_t)q * ldg + q];
                    const double apq = 0.5 * (G[(int64_t)p * ldg + q] + G[(int64_t)q * ldg + p]);
                    if (fabs(apq) > 1e-300 && fabs(apq) > 1e-18 * sqrt(fabs(app * aqq))) {
                        const double theta = (aqq - app) / (2.0 * apq);
                        const double tt = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(tt * tt + 1.0);
                        s = tt * c;
                    }
                } else { q = -1; }
                g_eig_p[t] = p; g_eig_q[t] = q; g_eig_c[t] = c; g_eig_s[t] = s;
            }
            grid.sync();
            // ---- phase B: G <- J^T G J on 2x2 blocks, V <- V J on row pairs
            const int64_t nblk = (int64_t)np * np;
            for (int64_t e = gtid; e < nblk + (int64_t)k * np; e += gsize) {
                if (e < nblk) {
                    const int a = (int)(e / np), b = (int)(e % np);
                    const int pa = g_eig_p[a], qa = g_eig_q[a], pb = g_eig_p[b], qb = g_eig_q[b];
                    const double ca = g_eig_c[a], sa = g_eig_s[a], cb = g_eig_c[b], sb = g_eig_s[b];
                    if (qa < 0 && qb < 0) {
                        // single element, identity rotations
                    } else if (qa < 0) {
                        double* gp = G + (int64_t)pa * ldg;
                        const double g0 = gp[pb], g1 = gp[qb];
                        gp[pb] = cb * g0 - sb * g1;
                        gp[qb] = sb * g0 + cb * g1;
                    } else if (qb < 0) {
                        double* g0p = G + (int64_t)pa * ldg + pb;
                        double* g1p = G + (int64_t)qa * ldg + pb;
                        const double g0 = *g0p, g1 = *g1p;
                        *g0p = ca * g0 - sa * g1;
                        *g1p = sa * g0 + ca * g1;
                    } else {
                        double* rp = G + (int64_t)pa * ldg;
                        double* rq = G + (int64_t)qa * ldg;
                        const double g00 = rp[pb], g01 = rp[qb], g10 = rq[pb], g11 = rq[qb];
                        const double h00 = cb * g00 - sb * g01, h01 = sb * g00 + cb * g01;
                        const double h10 = cb * g10 - sb * g11, h11 = sb * g10 + cb * g11;
                        double n00 = ca * h00 - sa * h10, n10 = sa * h00 + ca * h10;
                        double n01 = ca * h01 - sa * h11, n11 = sa * h01 + ca * h11;
                        if (a == b) { n01 = 0.0; n10 = 0.0; }
                        rp[pb] = n00; rp[qb] = n01; rq[pb] = n10; rq[qb] = n11;
                    }
                } else {
                    const int64_t f = e - nblk;
                    const int r = (int)(f / np), b = (int)(f % np);
                    const int pb = g_eig_p[b], qb = g_eig_q[b];
                    if (qb >= 0) {
                        double* vr = V + (int64_t)r * ldv;
                        const double cb = g_eig_c[b], sb = g_eig_s[b];
                        const double v0 = vr[pb], v1 = vr[qb];
                        vr[pb] = cb * v0 - sb * v1;
                        vr[qb] = sb * v0 + cb * v1;
                    }
                }
            }
            grid.sync();
        }
    }
    // ---- eigenvalues sorted descending; permute the columns of V through G as scratch
    for (int64_t i = gtid; i < k; i += gsize) {
        const double li = G[i * ldg + i];
        int rank = 0;
        for (int j = 0; j < k; ++j) {
            const double lj = G[(int64_t)j * ldg + j];
            rank += (lj > li) || (lj == li && j < i);
        }
        g_eig_rank[i] = rank;
        lam[rank] = li;
    }
    if (gtid == 0) g_eig_sweeps = sweep;
    grid.sync();
    for (int64_t e = gtid; e < (int64_t)k * k; e += gsize) {
        const int r = (int)(e / k), c = (int)(e % k);
        G[(int64_t)r * ldg + g_eig_rank[c]] = V[(int64_t)r * ldv + c];
    }
    grid.sync();
    for (int64_t e = gtid; e < (int64_t)k * k; e += gsize) {
        const int r = (int)(e / k), c = (int)(e % k);
        V[(int64_t)r * ldv + c] = G[(int64_t)r * ldg + c];
    }
}

// Small matrices (k <= EIG_SMALL_MAX): the same cyclic Jacobi in ONE CTA with G and V in shared
// memory - a round costs two __syncthreads instead of two grid-wide barriers (the cooperative
// kernel above spends ~1 us per barrier: 0.9 ms at k = 58, this one ~0.1 ms).
constexpr int EIG_SMALL_MAX = 104;   // 2 * 104 * 105 doubles = 171 KB of shared memory

__global__ void __launch_bounds__(1024) jacobi_small_kernel(double* __restrict__ Gg, int64_t ldg, int k,
                                                           double* __restrict__ lam,
                                                           double* __restrict__ Vg, int64_t ldv,
                                                           int max_sweeps, double tol2,
                                                           int* __restrict__ sweeps_out) {
    extern __shared__ double sm[];
    const int ld = k | 1;                        // odd leading dimension: fewer bank conflicts
    double* G = sm;
    double* V = G + (size_t)k * ld;
    double* cs = V + (size_t)k * ld;             // c[np], s[np]
    int* pq = (int*)(cs + 2 * ((k + 1) / 2 + 1));   // p[np], q[np]
    __shared__ double red[2][32];
    __shared__ double tot[2];
    const int tid = threadIdx.x, nth = blockDim.x;
    const int kp = (k + 1) & ~1, np = kp / 2;
    for (int e = tid; e < k * k; e += nth) {
        const int r = e / k, c = e % k;
        G[r * ld + c] = Gg[(int64_t)r * ldg + c];
        V[r * ld + c] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int e = tid; e < k * k; e += nth) {
            const int r = e / k, c = e % k;
            const double v = G[r * ld + c];
            if (r == c) dg = fma(v, v, dg); else off = fma(v, v, off);
        }
        for (int o = 16; o > 0; o >>= 1) {
            off += __shfl_xor_sync(0xffffffffu, off, o);
            dg += __shfl_xor_sync(0xffffffffu, dg, o);
        }
        if ((tid & 31) == 0) { red[0][tid >> 5] = off; red[1][tid >> 5] = dg; }
        __syncthreads();
        if (tid == 0) {
            double a = 0.0, b = 0.0;
            for (int w = 0; w < (nth >> 5); ++w) { a += red[0][w]; b += red[1][w]; }
            tot[0] = a; tot[1] = b;
        }
        __syncthreads();
        if (tot[0] <= tol2 * tot[1]) break;
        for (int round = 0; round < kp - 1; ++round) {
            for (int t = tid; t < np; t += nth) {
                int p, q;
                if (t == 0) { p = kp - 1; q = round; }
                else {
                    p = (round + t) % (kp - 1);
                    q = (round - t + (kp - 1)) % (kp - 1);
                }
                if (p > q) { const int tmp = p; p = q; q = tmp; }
                double c = 1.0, s = 0.0;
                if (q < k) {
                    const double app = G[p * ld + p], aqq = G[q * ld + q];
                    const double apq = 0.5 * (G[p * ld + q] + G[q * ld + p]);
                    if (fabs(apq) > 1e-300 && fabs(apq) > 1e-18 * sqrt(fabs(app * aqq))) {
                        const double theta = (aqq - app) / (2.0 * apq);
                        const double tt = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(tt * tt + 1.0);
                        s = tt * c;
                    }
                } else { q = -1; }
                pq[t] = p; pq[np + t] = q; cs[t] = c; cs[np + t] = s;
            }
            __syncthreads();
            const int nblk = np * np;
            for (int e = tid; e < nblk + k * np; e += nth) {
                if (e < nblk) {
                    const int a = e / np, b = e % np;
                    const int pa = pq[a], qa = pq[np + a], pb = pq[b], qb = pq[np + b];
                    const double ca = cs[a], sa = cs[np + a], cb = cs[b], sb = cs[np + b];
                    if (qa < 0 && qb < 0) {
                    } else if (qa < 0) {
                        double* gp = G + pa * ld;
                        const double g0 = gp[pb], g1 = gp[qb];
                        gp[pb] = cb * g0 - sb * g1;
                        gp[qb] = sb * g0 + cb * g1;
                    } else if (qb < 0) {
                        double* g0p = G + pa * ld + pb;
                        double* g1p = G + qa * ld + pb;
                        const double g0 = *g0p, g1 = *g1p;
                        *g0p = ca * g0 - sa * g1;
                        *g1p = sa * g0 + ca * g1;
                    } else {
                        double* rp = G + pa * ld;
                        double* rq = G + qa * ld;
                        const double g00 = rp[pb], g01 = rp[qb], g10 = rq[pb], g11 = rq[qb];
                        const double h00 = cb * g00 - sb * g01, h01 = sb * g00 + cb * g01;
                        const double h10 = cb * g10 - sb * g11, h11 = sb * g10 + cb * g11;
                        double n00 = ca * h00 - sa * h10, n10 = sa * h00 + ca * h10;
                        double n01 = ca * h01 - sa * h11, n11 = sa * h01 + ca * h11;
                        if (a == b) { n01 = 0.0; n10 = 0.0; }
                        rp[pb] = n00; rp[qb] = n01; rq[pb] = n10; rq[qb] = n11;
                    }
                } else {
                    const int f = e - nblk;
                    const int r = f / np, b = f % np;
                    const int pb = pq[b], qb = pq[np + b];
                    if (qb >= 0) {
                        double* vr = V + r * ld;
                        const double cb = cs[b], sb = cs[np + b];
                        const double v0 = vr[pb], v1 = vr[qb];
                        vr[pb] = cb * v0 - sb * v1;
                        vr[qb] = sb * v0 + cb * v1;
                    }
                }
            }
            __syncthreads();
        }
    }
    // eigenvalues sorted descending (ties: lower index first), columns of V permuted accordingly
    for (int i = tid; i < k; i += nth) {
        const double li = G[i * ld + i];
        int rank = 0;
        for (int j = 0; j < k; ++j) {
            const double lj = G[j * ld + j];
            rank += (lj > li) || (lj == li && j < i);
        }
        pq[i] = rank;                // np + np >= k entries available
        lam[rank] = li;
    }
    __syncthreads();
    for (int e = tid; e < k * k; e += nth) {
        const int r = e / k, c = e % k;
        Vg[(int64_t)r * ldv + pq[c]] = V[r * ld + c];
    }
    if (tid == 0) *sweeps_out = sweep;
}

// d_sweeps (optional device int): when given, the sweep count is left on the device and the call
// does not synchronise (the caller reads it together with its own results)
int sym_eig_async(double* G, int64_t ldg, int64_t k, double* lam, double* V, int64_t ldv,
                  int* d_sweeps, cudaStream_t st) {
    if (k == 0) return OCB_OK;
    if (k > EIG_SMALL_MAX) return -100;   // caller falls back to the cooperative kernel
    const int ld = (int)k | 1;
    const size_t smem = (size_t)(2 * k * ld + 2 * ((k + 1) / 2 + 1)) * sizeof(double) +
                        (size_t)(2 * ((k + 1) / 2) + 2) * sizeof(int) + 64;
    static bool attr_set = false;
    if (!attr_set) {
        OCB_CUDA(cudaFuncSetAttribute(jacobi_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    jacobi_small_kernel<<<1, 1024, smem, st>>>(G, ldg, (int)k, lam, V, ldv, 40, 1e-30, d_sweeps);
    OCB_LAUNCH_CHECK();
    return OCB_OK;
}

int sym_eig_impl(double* G, int64_t ldg, int64_t k, double* lam, double* V, int64_t ldv,
                 int32_t* h_sweeps, cudaStream_t st) {
    if (k == 0) { if (h_sweeps) *h_sweeps = 0; return OCB_OK; }
    if (k > EIG_MAXK) { set_error("sym_eig: k=%lld > %d", (long long)k, EIG_MAXK); return OCB_ERR_ARG; }
    static const bool no_small = getenv("OCB_NO_SMALL_EIG") != nullptr;
    if (k <= EIG_SMALL_MAX && !no_small) {
        static int* d_sw = nullptr;
        if (!d_sw) OCB_CUDA(cudaMalloc((void**)&d_sw, sizeof(int)));
        const int rc = sym_eig_async(G, ldg, k, lam, V, ldv, d_sw, st);
        if (rc) return rc;
        if (h_sweeps) {
            int sw = 0;
            OCB_CUDA(cudaMemcpyAsync(&sw, d_sw, sizeof(int), cudaMemcpyDeviceToHost, st));
            OCB_CUDA(cudaStreamSynchronize(st));
            *h_sweeps = sw;
            if (sw >= 40) { set_error("sym_eig: no convergence in 40 sweeps"); return OCB_ERR_NOCONV; }
        }
        return OCB_OK;
    }
    int per_sm = 0;
    OCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_kernel, 256, 0));
    const int64_t np = (k + 1) / 2;
    const int64_t work = np * np + k * np;
    int64_t blocks = (work + 255) / 256;
    blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, std::min(per_sm, 4) * (int64_t)sm_count()));
    blocks = std::min<int64_t>(blocks, 1024);
    int kk = (int)k, max_sweeps = 40;
    double tol2 = 1e-30;  // off(G)^2 <= tol2 * diag(G)^2
    void* args[] = {&G, &ldg, &kk, &lam, &V, &ldv, &max_sweeps, &tol2};
    OCB_CUDA(cudaLaunchCooperativeKernel((void*)jacobi_kernel, dim3((unsigned)blocks), dim3(256), args, 0, st));
    count_launch();
    if (h_sweeps) {
        int sw = 0;
        OCB_CUDA(cudaMemcpyFromSymbolAsync(&sw, g_eig_sweeps, sizeof(int), 0, cudaMemcpyDeviceToHost, st));
        OCB_CUDA(cudaStreamSynchronize(st));
        *h_sweeps = sw;
        if (sw >= max_sweeps) { set_error("sym_eig: no convergence in %d sweeps", max_sweeps); return OCB_ERR_NOCONV; }
    }
    return OCB_OK;
}

// ---------------------------------------------------------------------------------
// m x m inverse (SMW core  (I - V A^-1 U)^-1 ), one warp, Gauss-Jordan with partial pivoting.
// In: C (m x m row-major, ld) holds  V A^-1 U ; Out: Sinv = (I - C)^-1.  flag[0]=1 if singular.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) smw_core_inv_kernel(const double* __restrict__ C, int64_t ldc,
                                                         int m, double* __restrict__ Sinv,
                                                         int* __restrict__ flag) {
    __shared__ double A[32][65];
    const int lane = threadIdx.x;
    for (int r = 0; r < m; ++r)
        for (int c = lane; c < 2 * m; c += 32)
            A[r][c] = (c < m) ? ((r == c ? 1.0 : 0.0) - C[(int64_t)r * ldc + c]) : ((c - m) == r ? 1.0 : 0.0);
    __syncwarp();
    for (int col = 0; col < m; ++col) {
        int piv = col;
        double best = fabs(A[col][col]);
        for (int r = col + 1; r < m; ++r)
            if (fabs(A[r][col]) > best) { best = fabs(A[r][col]); piv = r; }
        if (best == 0.0) { if (lane == 0) *flag = 1; return; }
        if (piv != col)
            for (int c = lane; c < 2 * m; c += 32) { const double t = A[col][c]; A[col][c] = A[piv][c]; A[piv][c] = t; }
        __syncwarp();
        const double d = 1.0 / A[col][col];
        __syncwarp();
        for (int c = lane; c < 2 * m; c += 32) A[col][c] *= d;
        __syncwarp();
        for (int r = 0; r < m; ++r) {
            if (r == col) continue;
            const double f = A[r][col];
            __syncwarp();
            for (int c = lane; c < 2 * m; c += 32) A[r][c] = fma(-f, A[col][c], A[r][c]);
            __syncwarp();
        }
    }
    for (int r = 0; r < m; ++r)
        for (int c = lane; c < m; c += 32) Sinv[r * m + c] = A[r][m + c];
}

int smw_core_inv(const double* C, int64_t ldc, int m, double* Sinv, int* flag, cudaStream_t st) {
    if (m > 32) { set_error("SMW rank m=%d > 32", m); return OCB_ERR_ARG; }
    smw_core_inv_kernel<<<1, 32, 0, st>>>(C, ldc, m, Sinv, flag);
    OCB_LAUNCH_CHECK();
    return OCB_OK;
}

// ---------------------------------------------------------------------------------
// FP64 peak micro-benchmark (roofline denominator of the FP64-bound kernels): register-only
// chains of DMMA.8x8x4 (kind 0) or DFMA (kind 1), 8 independent accumulator pairs per thread.
// ---------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) fp64_peak_kernel(int64_t iters, double seed, double* __restrict__ sink) {
    double c[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = seed * (threadIdx.x + j); c[j][1] = seed; }
    double a = 1.0 + seed, b = 1.0 - seed;
    for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (KIND == 0) {
                dmma_8x8x4(c[j][0], c[j][1], a, b);
            } else {
                c[j][0] = fma(a, c[j][0], b);
                c[j][1] = fma(b, c[j][1], a);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    if (s == 123.456) sink[0] = s;   // keeps the chains alive
}

}  // namespace ocb

extern "C" {

int ocb_fp64_peak(int kind, int64_t iters, int64_t ctas_per_sm, double* h_tflops, double* d_sink, void* stream) {
    using namespace ocb;
    OCB_ARG((kind == 0 || kind == 1) && iters > 0 && ctas_per_sm >= 1 && ctas_per_sm <= 8 && h_tflops && d_sink,
            "fp64_peak");
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    OCB_CUDA(cudaEventCreate(&e0));
    OCB_CUDA(cudaEventCreate(&e1));
    const unsigned blocks = (unsigned)(sm_count() * ctas_per_sm);
    for (int rep = 0; rep < 2; ++rep) {   // first pass warms up
        OCB_CUDA(cudaEventRecord(e0, st));
        if (kind == 0) fp64_peak_kernel<0><<<blocks, 256, 0, st>>>(iters, 1e-9, d_sink);
        else fp64_peak_kernel<1><<<blocks, 256, 0, st>>>(iters, 1e-9, d_sink);
        OCB_LAUNCH_CHECK();
        OCB_CUDA(cudaEventRecord(e1, st));
        OCB_CUDA(cudaEventSynchronize(e1));
    }
    float ms = 0.f;
    OCB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // per thread and iteration: kind 0: 8 DMMA = 8 * 512 flops per WARP; kind 1: 16 DFMA = 32 flops
    const double per_block = kind == 0 ? 8.0 * 8.0 * 512.0 : 256.0 * 32.0;
    *h_tflops = per_block * (double)iters * (double)blocks / ((double)ms * 1e-3) / 1e12;
    return OCB_OK;
}

int64_t ocb_gram_ws_bytes(int64_t n, int64_t ka, int64_t kb) {
    int nsplit, nsplit2 = 0, nsplit3 = 0;
    int64_t rps;
    ocb::gram_plan(n, std::max<int64_t>(ka, 1), std::max<int64_t>(kb, 1), &nsplit, &rps);
    if (ocb::gram_big_enabled() && ka >= 128 && kb >= 128) {
        ocb::gram_plan2(n, ka, kb, false, &nsplit2, &rps);
        ocb::gram_plan2(n, ka, kb, true, &nsplit3, &rps);
    }
    return (int64_t)std::max(nsplit, std::max(nsplit2, nsplit3)) * ka * kb * 8;
}

int ocb_gram(const double* d_Z, int64_t ldz, int64_t ka, const double* d_W, int64_t ldw, int64_t kb,
             int64_t n, double* d_G, int64_t ldg, void* d_ws, int64_t ws_bytes, void* stream) {
    OCB_ARG(ka >= 0 && kb >= 0 && n >= 0 && ldz >= ka && ldw >= kb && ldg >= kb, "gram sizes");
    OCB_ARG(ka == 0 || kb == 0 || (d_Z && d_W && d_G), "gram null");
    return ocb::gram_impl(d_Z, ldz, ka, d_W, ldw, kb, n, d_G, ldg, d_ws, ws_bytes, (cudaStream_t)stream);
}

int ocb_tall_gemm(const double* d_Z, int64_t ldz, int64_t n, int64_t k, const double* d_T, int64_t ldt,
                  int64_t kc, double* d_C, int64_t ldc, double alpha, double beta, void* stream) {
    OCB_ARG(n >= 0 && k >= 0 && kc >= 0 && ldz >= k && ldt >= kc && ldc >= kc, "tall_gemm sizes");
    OCB_ARG(n == 0 || kc == 0 || (d_C && (k == 0 || (d_Z && d_T))), "tall_gemm null");
    return ocb::tall_gemm_impl(d_Z, ldz, n, k, d_T, ldt, kc, d_C, ldc, alpha, beta, (cudaStream_t)stream);
}

int ocb_sym_eig(double* d_G, int64_t ldg, int64_t k, double* d_lam, double* d_V, int64_t ldv,
                int32_t* h_sweeps, void* stream) {
    OCB_ARG(k >= 0 && ldg >= k && ldv >= k, "sym_eig sizes");
    OCB_ARG(k == 0 || (d_G && d_lam && d_V), "sym_eig null");
    return ocb::sym_eig_impl(d_G, ldg, k, d_lam, d_V, ldv, h_sweeps, (cudaStream_t)stream);
}
}
