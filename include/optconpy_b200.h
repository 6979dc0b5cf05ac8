/*
 * optconpy_b200 — C ABI of the B200 (sm_100a) kernels behind the
 * sadptprj_riclyap_adi interface that highlando/optconpy drives.
 *
 * The reference has no FFI of its own: its boundary is the Python module
 * functions of sadptprj_riclyap_adi.{lin_alg_utils,proj_ric_utils} that
 * optcont_main.py:13-14 and solve_dae_ric.py:3-4 import.  Each entry point below
 * names the reference call site(s) whose numerical work it carries.  The Python
 * shim optconpy_b200/{lin_alg_utils,proj_ric_utils}.py binds these with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, a negative ocb_status otherwise;
 *     ocb_last_error() gives the message of the last failure on this thread.
 *   - dense blocks are FP64, ROW-major, leading dimension in elements (ld >= k):
 *     element (i, c) of an (n x k) block is ptr[i*ld + c].
 *   - sparse matrices are CSR, int32 indices, FP64 values.
 *   - pointers named d_* are DEVICE pointers (e.g. torch.Tensor.data_ptr()),
 *     h_* are HOST pointers.  No torch types cross this boundary.
 *   - stream is a cudaStream_t passed as void*; nothing uses a hidden stream.
 */
#ifndef OPTCONPY_B200_H
#define OPTCONPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    OCB_OK = 0,
    OCB_ERR_ARG = -1,      /* bad argument */
    OCB_ERR_CUDA = -2,     /* CUDA runtime error (message has the detail) */
    OCB_ERR_SINGULAR = -3, /* zero pivot / singular small system */
    OCB_ERR_NOCONV = -4,   /* Jacobi eigen-solver hit its sweep limit */
    OCB_ERR_CAPACITY = -5  /* a caller-provided buffer is too small */
} ocb_status;

const char* ocb_last_error(void);
int ocb_version(void);
/* number of kernel launches issued by this library since process start (bench.py: gpu_launches) */
int64_t ocb_launch_count(void);
/* How host threads of this process wait for the CURRENT device (cudaStreamSynchronize and friends):
 * blocking = 1 sleeps on an interrupt instead of spinning.  With one process per GPU and fewer than
 * ~3 host cores per process the spinning main threads otherwise take the cores the host LU workers
 * need (8 GPUs on a 16-core host: e2e 55 -> see DESIGN.md section 7). */
int ocb_set_sync_mode(int blocking);

/* ---- K2: CSR x dense block ---------------------------------------------------
 * Y = alpha * S * X + beta * Y,  S (nrows x ncols) CSR, X (ncols x k), Y (nrows x k).
 * Carries `MT*Zc`, `Mt*V` inside the ADI, `jmat*mic`, `tb_mat.T*Y`
 * (solve_dae_ric.py:78,149,187; the SpMM of every ADI step). */
int ocb_spmm(int64_t nrows, int64_t ncols,
             const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_vals,
             const double* d_X, int64_t ldx, double* d_Y, int64_t ldy, int64_t k,
             double alpha, double beta, void* stream);

/* ---- K1: multi-RHS sparse triangular solves on an LU factorisation ------------
 * Replaces the SuperLU solve behind spsla.factorized(...) that
 * lau.solve_sadpnt_smw / pru.solve_proj_lyap_stein call once per column
 * (solve_dae_ric.py:192-194; every ADI step).  The factorisation itself
 * (Pr*A*Pc = L*U, L unit lower, both CSR on the host) is the separately timed
 * setup step; ocb_lu_create analyses the dependency levels and uploads. */
typedef struct ocb_lu ocb_lu;

int ocb_lu_create(ocb_lu** out, int64_t n,
                  const int32_t* h_L_rowptr, const int32_t* h_L_colidx, const double* h_L_vals,
                  const int32_t* h_U_rowptr, const int32_t* h_U_colidx, const double* h_U_vals,
                  const int32_t* h_perm_r, const int32_t* h_perm_c, void* stream);
int ocb_lu_destroy(ocb_lu* lu);
/* The same in two halves, so that the analysis can run where the host LU ran (a worker
 * process without a CUDA context): ocb_lu_pack_host builds the self-describing device image
 * (malloc'ed; release with ocb_host_free) for a GPU with max_smem_optin bytes of opt-in shared
 * memory per block (flags bit 1: TRANSPOSED layout - the factors come from an LU factorisation
 * of A^T handed over column-wise, A = U^T L^T: the first three arrays are then the rows of the
 * lower factor U^T, which carries the pivots, the next three the rows of the unit upper factor
 * L^T, and h_perm_r / h_perm_c are that factorisation's perm_c / perm_r; SuperLU's compressed-
 * column output is used as is, no transposition on the host.  flags bit 2: supernodes up to 64
 * rows wide are solved in ONE sub-level, x_t = inv(T_tt) b_t - (inv(T_tt) T[t,off]) x, instead of
 * two (about a third fewer sub-level barriers for ~2 % more entries).  flags bits 4..7: cluster size of the column-panel kernel, 1 / 2 / 3 / 4 / 8 (0 = default 4;
 * 4 is fastest up to 33 right-hand sides, 3 up to 44, 2 from 45 to ~150).  flags bit 0: also include the flat program of the wide, all-columns-at-once
 * executor (its PANEL form, see ocb_lu_program_solve_host) that ocb_lu_solve uses for k >= 640
 * right-hand sides; it is always included when the column panel does not fit shared memory); ocb_lu_create_from_image uploads it with one copy into d_arena (bytes long,
 * 256-byte aligned, owned by the caller and kept alive until ocb_lu_destroy; NULL: the library
 * allocates and frees its own). */
int ocb_lu_pack_host(int64_t n,
                     const int32_t* h_L_rowptr, const int32_t* h_L_colidx, const double* h_L_vals,
                     const int32_t* h_U_rowptr, const int32_t* h_U_colidx, const double* h_U_vals,
                     const int32_t* h_perm_r, const int32_t* h_perm_c, int64_t max_smem_optin,
                     int64_t flags, unsigned char** out_image, int64_t* out_bytes);
/* The same, building the image directly in the caller's buffer (e.g. a page-locked shared-memory
 * segment): no intermediate copy.  OCB_ERR_CAPACITY if it does not fit; *out_bytes then holds
 * the size the image needs. */
int ocb_lu_pack_host_into(int64_t n,
                          const int32_t* h_L_rowptr, const int32_t* h_L_colidx, const double* h_L_vals,
                          const int32_t* h_U_rowptr, const int32_t* h_U_colidx, const double* h_U_vals,
                          const int32_t* h_perm_r, const int32_t* h_perm_c, int64_t max_smem_optin,
                          int64_t flags, unsigned char* dst, int64_t dst_capacity, int64_t* out_bytes);
/* The same with a RESIDUAL GUARD (what the LU workers call): the finished gather program is
 * executed on the host for one fixed pseudo-random right-hand side and *out_backerr receives the
 * normwise backward error ||b - A x|| / (||A||_F ||x|| + ||b||) against the ORIGINAL matrix A
 * (CSC arrays).  It covers both error sources the reference's spsla.factorized does not have -
 * the relaxed pivoting of the host LU and the explicit inverses of the supernode blocks - so
 * the caller can re-factorise with full partial pivoting and flags bit 3 (no supernode wider
 * than 32 rows, no one-step blocks) when it is too large.  dst != NULL: build in the caller's
 * buffer (as ocb_lu_pack_host_into); dst == NULL: *out_image is malloc'ed. */
int ocb_lu_pack_host_checked(int64_t n,
                             const int32_t* h_L_rowptr, const int32_t* h_L_colidx, const double* h_L_vals,
                             const int32_t* h_U_rowptr, const int32_t* h_U_colidx, const double* h_U_vals,
                             const int32_t* h_perm_r, const int32_t* h_perm_c, int64_t max_smem_optin,
                             int64_t flags, unsigned char* dst, int64_t dst_capacity,
                             unsigned char** out_image, int64_t* out_bytes,
                             const int32_t* h_A_colptr, const int32_t* h_A_rowidx, const double* h_A_vals,
                             double* out_backerr);
void ocb_host_free(void* p);
int ocb_lu_create_from_image(ocb_lu** out, const unsigned char* h_image, int64_t bytes,
                             void* d_arena, void* stream);
/* info[0..7] = n, nnz(L) strictly lower, nnz(U) incl. diagonal, #sub-levels of the L sweep,
 *              #sub-levels of the U sweep (one CTA barrier each), device bytes held, widest
 *              column panel that fits shared memory (0 = panel lives in a global slab),
 *              number of stream batches */
int ocb_lu_info(const ocb_lu* lu, int64_t* info8);
/* info[0..7] = rows of the extended vector (n + y scratch), #supernodes, widest supernode,
 *              #slices, #program rows, #program entries (padded), ring stage bytes, ring stages */
int ocb_lu_stats(const ocb_lu* lu, int64_t* info8);
/* profiling aid: number of sub-levels of the panel program (0: none); h_out3 (optional, 3 per
 * sub-level) = panels, longest column list, 0 */
int64_t ocb_lu_panel_levels(const ocb_lu* lu, int64_t* h_out3, int64_t capacity_levels);
/* debugging aid (OCB_TRSM_TRACE=1): SM clock of CTA 0 after the panel load and after every
 * sub-level barrier of the most recent solve */
int ocb_debug_trace(int64_t* h_out, int64_t count);

/* Host-only view of the same analysis (no CUDA call; usable without a GPU): the GATHER
 * PROGRAM that ocb_lu_create packs and uploads.  Supernodes of U (rows with nested structure)
 * get their diagonal blocks inverted on the host, which turns the two triangular solves into
 * a short sequence of sub-levels of independent rows
 *     xe[dst] = ((init >= 0 ? xe[init] : 0) - sum_p val[p]*xe[col[p]]) * scale
 * on the extended vector xe = [x (n) | y scratch].  Sub-level s owns the slices
 * [sub_ptr[s], sub_ptr[s+1]); slice = {ebase, trips, glog | nrows << 8, q0}: nrows <= 32 >> glog
 * rows starting at row q0, 2^glog lanes per row, entries stored trip-major (sliced ELLPACK):
 * entry (trip u, lane l = r * 2^glog + g) at ebase + 32 u + l is entry u * 2^glog + g of row
 * q0 + r, zero padded.  tests/test_lu_program.py executes it with numpy against SuperLU. */
typedef struct ocb_lu_program ocb_lu_program;
int ocb_lu_program_create(ocb_lu_program** out, int64_t n,
                          const int32_t* h_L_rowptr, const int32_t* h_L_colidx, const double* h_L_vals,
                          const int32_t* h_U_rowptr, const int32_t* h_U_colidx, const double* h_U_vals,
                          int64_t flags);
int ocb_lu_program_destroy(ocb_lu_program* prog);
/* The builder keeps the STRUCTURE of the last four programs by the index arrays they were built
 * from; a factor with the same index arrays (other shift, other time step: same ordering and
 * pivots) only has its numbers recomputed (OCB_NO_TEMPLATE=1 switches that off).  Number of
 * builds of this process served that way: */
int64_t ocb_lu_program_template_hits(void);
/* info[0..11] = n, n_ext, ymax, sub-levels L, sub-levels U, #supernodes, widest supernode,
 *               #slices, #rows, #entries (padded), nnz(L) strictly lower, nnz(U) */
int ocb_lu_program_info(const ocb_lu_program* prog, int64_t* info12);
/* h_sub_ptr: sub-levels + 1 ints; h_slice4: 4 ints per slice; the rest per row / per entry */
int ocb_lu_program_export(const ocb_lu_program* prog, int32_t* h_sub_ptr, int32_t* h_slice4,
                          int32_t* h_dst, int32_t* h_init, double* h_scale,
                          int32_t* h_col, double* h_val);
/* Host execution of the program for ONE right-hand side, x = A^-1 b (what the kernels do, serially;
 * the residual guard and the CPU tests use it).  mode 0: the row program; mode 1: its PANEL form
 * (up to 8 rows of a supernode share one zero-padded column list, values interleaved - the layout of
 * the register-blocked all-columns-at-once executor; max_pad bounds the padding, <= 1 = default 1.6);
 * h_stats4 (optional, mode 1) = panels, stored values, actual entries, sub-levels. */
int ocb_lu_program_solve_host(const ocb_lu_program* prog, const int32_t* h_perm_r, const int32_t* h_perm_c,
                              const double* h_b, double* h_x, int64_t mode, double max_pad,
                              int64_t* h_stats4);
/* ---- numeric-only refactorisation with symbolic reuse (host side) ----------
 * Replaces, for the second and every later matrix of one sparsity pattern, the per-matrix
 * `spsla.splu` of the reference's hot path (proj_ric_utils.py:108-111 factorises A + p_j M for every
 * shift j; solve_dae_ric.py:147-163 rebuilds A for every time step: same pattern, new numbers).
 * ocb_refactor_create takes the pattern of A (CSC) and the two permutations a first, pivoting
 * factorisation produced, (P A Q)[perm_r[i], perm_c[j]] = A[i, j] with the pivots on the diagonal,
 * and computes everything that depends on them only (elimination tree, supernodes, front
 * structures, index maps, CSR structure of both factors).  ocb_refactor_numeric then factorises a
 * matrix with these static pivots: one multifrontal pass over dense fronts; OCB_ERR_SINGULAR on a
 * zero pivot.  Its outputs, together with ocb_refactor_structure's index arrays and permutations
 * (the input ones composed with an elimination-tree postorder), are the L / U arguments of
 * ocb_lu_pack_host* with flags bit 1 CLEAR (P A Q = L U, L unit lower).  The caller's residual
 * guard (ocb_lu_pack_host_checked) decides whether the static pivots were good enough. */
/* Fill- AND depth-reducing ordering of a symmetric sparsity pattern (adjacency in CSR, both
 * triangles, diagonal entries ignored): nested dissection by BFS level structures with thinned
 * separators, sets of <= leaf nodes kept as they are.  h_order_out[k] = the node eliminated k-th.
 * Replaces SuperLU's minimum-degree ordering inside the reference's `spsla.factorized` /
 * `spsla.splu` calls (lin_alg_utils.py:95-96, proj_ric_utils.py:108-111): the solve kernels are
 * bound by the height of the elimination tree, which this ordering halves. */
int ocb_order_nd(int64_t n, const int32_t* h_adj_rowptr, const int32_t* h_adj_colidx, int64_t leaf,
                 int32_t* h_order_out);
/* Saddle-point constraint on an elimination order: nodes with a zero diagonal (h_diag_is_zero[v] != 0: the
 * pressure block of [[A, J^T], [J, 0]]) are delayed until right after their first neighbour with a
 * non-zero diagonal, so that their pivot is the Schur-complement entry and never an exact zero. */
int ocb_order_delay_zero_diagonals(int64_t n, const int32_t* h_adj_rowptr, const int32_t* h_adj_colidx,
                                   const uint8_t* h_diag_is_zero, const int32_t* h_order_in,
                                   int32_t* h_order_out);
typedef struct ocb_refactor ocb_refactor;
int ocb_refactor_create(ocb_refactor** out, int64_t n, const int32_t* h_A_colptr, const int32_t* h_A_rowidx,
                        const int32_t* h_perm_r, const int32_t* h_perm_c);
int ocb_refactor_destroy(ocb_refactor* rf);
/* info8: n, nnz(L) incl. the unit diagonal, nnz(U), supernodes, largest front, peak stack
 * entries, flops of one numeric pass, nnz(A) */
int ocb_refactor_info(const ocb_refactor* rf, int64_t* info8);
/* any pointer may be null; sizes from ocb_refactor_info */
int ocb_refactor_structure(const ocb_refactor* rf, int32_t* h_L_rowptr, int32_t* h_L_colidx,
                           int32_t* h_U_rowptr, int32_t* h_U_colidx, int32_t* h_perm_r, int32_t* h_perm_c);
/* h_A_vals in the order of the CSC arrays given to ocb_refactor_create; not re-entrant per handle */
int ocb_refactor_numeric(ocb_refactor* rf, const double* h_A_vals, double* h_L_vals, double* h_U_vals);

/* bytes of device workspace ocb_lu_solve needs for k right-hand sides (0 when the column-panel
 * kernel is used; n_ext x roundup(k) doubles for the wide executor) */
int64_t ocb_lu_solve_ws_bytes(const ocb_lu* lu, int64_t k);
/* X[0:nrows_x, 0:k] = (A^-1 [B[0:nrows_b, 0:k]; 0])[0:nrows_x].  B and X may alias. */
int ocb_lu_solve(const ocb_lu* lu, const double* d_B, int64_t ldb, int64_t nrows_b,
                 double* d_X, int64_t ldx, int64_t nrows_x, int64_t k,
                 void* d_ws, int64_t ws_bytes, void* stream);

/* Per-launch timing of the solve kernel: when enabled, every ocb_lu_solve launch (also the
 * ones inside ocb_adi_run / ocb_smw_solve) is bracketed by CUDA events on its stream.
 * ocb_prof_collect waits for them and returns the summed kernel time, the launch count and the
 * summed ALGORITHMIC bytes (12*(nnzL+nnzU) + 16*(n+1) + 32*n*k per launch). */
int ocb_prof_enable(int on);
int ocb_prof_collect(double* total_ms, int64_t* launches, double* alg_bytes);

/* ---- K3: Gram product on FP64 tensor cores (DMMA) -----------------------------
 * G (ka x kb) = Z^T W over n rows; deterministic (fixed split + ordered reduce).
 * Carries np.dot(Z.T, M*Z), Z.T*tB, the factored norms
 * (tests/test_units_compfacres_compress.py:85-86; solve_dae_ric.py:101,183,189). */
int64_t ocb_gram_ws_bytes(int64_t n, int64_t ka, int64_t kb);
int ocb_gram(const double* d_Z, int64_t ldz, int64_t ka,
             const double* d_W, int64_t ldw, int64_t kb, int64_t n,
             double* d_G, int64_t ldg, void* d_ws, int64_t ws_bytes, void* stream);

/* Measured FP64 peak of the device this runs on (roofline denominator of the FP64-bound
 * kernels; MEASURED_PEAKS.json only has HBM and bf16): register-only chains of DMMA.8x8x4
 * (kind 0, the FP64 tensor pipe the Gram / tall products use) or DFMA (kind 1), ctas_per_sm
 * CTAs of 256 threads per SM, timed with CUDA events.  d_sink: one device double. */
int ocb_fp64_peak(int kind, int64_t iters, int64_t ctas_per_sm, double* h_tflops, double* d_sink,
                  void* stream);

/* C (n x kc) = alpha * Z (n x k) * T (k x kc) + beta * C   (DMMA; `Z*V_k`, `Z*(Z^T tB)`) */
int ocb_tall_gemm(const double* d_Z, int64_t ldz, int64_t n, int64_t k,
                  const double* d_T, int64_t ldt, int64_t kc,
                  double* d_C, int64_t ldc, double alpha, double beta, void* stream);

/* ---- K4: small symmetric eigen-decomposition (parallel cyclic Jacobi) ----------
 * G (k x k, symmetric, destroyed) = V diag(lam) V^T, lam sorted descending,
 * V row-major with eigenvectors in its columns.  Cooperative launch. */
int ocb_sym_eig(double* d_G, int64_t ldg, int64_t k, double* d_lam, double* d_V, int64_t ldv,
                int32_t* h_sweeps, void* stream);

/* ---- K3+K4: pru.compress_Zsvd (solve_dae_ric.py:162; optcont_main.py:498) -------
 * Zc = Z V_keep with the singular values > thresh, at most kmax (kmax <= 0: no cap;
 * thresh < 0: no threshold).  Rank-revealing pivoted Cholesky of Z^T Z (relative
 * stop eta), Jacobi on the r x r core, two tall products.  h_info[0]=kept columns,
 * [1]=Cholesky rank r, [2]=Jacobi sweeps; d_sigma (optional, >= rmax) gets the
 * singular values of the core. */
int64_t ocb_compress_ws_bytes(int64_t n, int64_t K, int64_t rmax);
int ocb_compress(const double* d_Z, int64_t ldz, int64_t n, int64_t K,
                 double thresh, int64_t kmax, double eta, int64_t rmax,
                 double* d_Zc, int64_t ldzc, int64_t zc_capacity_cols,
                 double* d_sigma, int64_t* h_info3,
                 void* d_ws, int64_t ws_bytes, void* stream);

/* The replicated half of the same compression, from a K x K Gram matrix G = Z^T Z the caller
 * already holds (column-sharded runs: G is the all-reduced sum of the ranks' partial products,
 * SURVEY 8e): pivoted Cholesky + core eigen-decomposition -> T (K x keep, row-major) with
 * Zc = Z T, so every rank can apply T to its own rows of Z.  h_info3 as ocb_compress. */
int64_t ocb_compress_gram_ws_bytes(int64_t K, int64_t rmax);
int ocb_compress_from_gram(const double* d_G, int64_t ldg, int64_t K, double thresh, int64_t kmax,
                           double eta, int64_t rmax, double* d_T, int64_t ldt, int64_t t_capacity_cols,
                           double* d_sigma, int64_t* h_info3, void* d_ws, int64_t ws_bytes, void* stream);

/* ---- peer-memory movement for column-sharded runs -------------------------
 * d_dst_peer / h_peer_ptrs are DEVICE pointers into buffers of other GPUs mapped into this
 * process (symmetric memory over NVLink / NVSwitch); ordering between the ranks is the caller's
 * (a symmetric-memory barrier between the phases).
 *   ocb_p2p_put2d:     dst[r*ldd + c] = src[r*lds + c] (P2P stores): the column -> row re-shard of
 *                      the factor before the Gram product, the all-gather of the compressed rows
 *   ocb_p2p_sum_peers: out[e] = sum_g peer_g[e], ranks in fixed order (P2P loads): the K x K Gram
 *                      all-reduce as the epilogue of the local partial products; bitwise the same
 *                      result on every rank. */
int ocb_p2p_put2d(const double* d_src, int64_t lds, int64_t nrows, int64_t ncols, double* d_dst_peer,
                  int64_t ldd, void* stream);
int ocb_p2p_sum_peers(const double* const* h_peer_ptrs, int64_t world, int64_t count, double* d_out,
                      void* stream);

/* ---- K6/K7 + the LR-ADI loop: pru.solve_proj_lyap_stein -------------------------
 * (tests/test_units_compfacres_compress.py:62-64; called by
 * pru.proj_alg_ric_newtonadi, solve_dae_ric.py:152-159, optcont_main.py:488-492).
 * Li-White recurrence on saddle-point solves with per-shift LU handles:
 *   V_1 = sqrt(-2 mu_1) S_1(W),  V_i = sqrt(mu_i/mu_{i-1}) (V_{i-1} - (mu_i+mu_{i-1}) S_i(Mt V_{i-1})),
 *   S_i(R) = first NV rows of (A_i - Ue Ve)^-1 [R; 0],  A_i = [[At + mu_i Mt, J^T],[J, 0]].
 * Low-rank part (optional, m columns): Ue = [d_Ufb (NV x m dense); 0],
 * Ve = [Vt (m x NV) CSR, 0]  applied by Sherman-Morrison-Woodbury.
 * Blocks V_i are written side by side into d_Z (NV x ldz), block i at column i*k.
 * Stops after step i when ||V_i||_F / ||[V_1..V_i]||_F <= reltol or i == maxsteps
 * (maxsteps <= 8192).  The test is evaluated ON THE DEVICE (last CTA of the fused update kernel,
 * same arithmetic and order as a host loop: z += v; rel = sqrt(v / z)); the host enqueues
 * iterations in batches and every kernel of an iteration past the converged one returns at
 * once, so there is no host round trip per step (OCB_ADI_CHUNK=1 restores one).
 * h_relnorms (maxsteps doubles) receives the ratios; *h_steps the number of blocks.
 * Threading: the library keeps one pinned scratch block and one side stream per process - call
 * its entry points from one host thread at a time per device (distinct streams are fine). */
/* Column-sharded runs: every rank iterates on its own slice of the right-hand-side
 * columns; the stopping test needs the GLOBAL ||V_i||_F^2.  When a hook is set (per calling
 * thread; NULL clears it) ocb_adi_run passes the local value through it after every step and
 * uses what comes back - the caller all-reduces the scalar (torch.distributed / NCCL). */
int ocb_adi_set_norm_hook(void (*hook)(double* v_nsq, void* ctx), void* ctx);
int64_t ocb_adi_ws_bytes(int64_t n_sad, int64_t k, int64_t m, int64_t nshifts,
                         ocb_lu* const* lus);
int ocb_adi_run(ocb_lu* const* lus, const double* h_shifts, int64_t nshifts,
                int64_t NV, int64_t NP,
                const int32_t* d_Mt_rowptr, const int32_t* d_Mt_colidx, const double* d_Mt_vals,
                const double* d_W, int64_t ldw, int64_t k,
                const double* d_Ufb, int64_t ldu, int64_t m,
                const int32_t* d_Vt_rowptr, const int32_t* d_Vt_colidx, const double* d_Vt_vals,
                int64_t maxsteps, double reltol,
                double* d_Z, int64_t ldz, int64_t z_capacity_cols,
                double* h_relnorms, int64_t* h_steps,
                void* d_ws, int64_t ws_bytes, void* stream);

/* One saddle-point solve with SMW low-rank update: lau.solve_sadpnt_smw
 * (solve_dae_ric.py:192-194; optcont_main.py:510-514), lau.app_prj_via_sadpnt
 * (optcont_main.py:405-408).  X (n_sad x k) = (A - Ue Ve)^-1 [B (nrows_b x k); 0]. */
int64_t ocb_smw_solve_ws_bytes(const ocb_lu* lu, int64_t k, int64_t m);
int ocb_smw_solve(const ocb_lu* lu, int64_t NV,
                  const double* d_B, int64_t ldb, int64_t nrows_b, int64_t k,
                  const double* d_Ufb, int64_t ldu, int64_t m,
                  const int32_t* d_Vt_rowptr, const int32_t* d_Vt_colidx, const double* d_Vt_vals,
                  double* d_X, int64_t ldx, int64_t nrows_x,
                  void* d_ws, int64_t ws_bytes, void* stream);

/* ---- K5: pru.get_mTzzTtb (solve_dae_ric.py:101,183,189; optcont_main.py:505-506) --
 * Out (NV x m) = alpha * Mt * (Z * (Z^T tB)),  Z (NV x kz), tB (NV x m) dense. */
int64_t ocb_feedback_ws_bytes(int64_t NV, int64_t kz, int64_t m);
int ocb_feedback(const int32_t* d_Mt_rowptr, const int32_t* d_Mt_colidx, const double* d_Mt_vals,
                 int64_t NV, const double* d_Z, int64_t ldz, int64_t kz,
                 const double* d_tB, int64_t ldb, int64_t m,
                 double* d_Out, int64_t ldo, double alpha,
                 void* d_ws, int64_t ws_bytes, void* stream);

/* sum of squares of an (n x k) block -> *d_out (one double); deterministic */
int ocb_sqnorm(const double* d_X, int64_t ldx, int64_t n, int64_t k, double* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OPTCONPY_B200_H */
