// Host-side "gather program" of a sparse LU solve (no CUDA in this header).
//
// x = Pc U^-1 L^-1 Pr b is rewritten as a short sequence of SUB-LEVELS.  Rows are
// grouped into supernodes (consecutive rows of U with nested structure); the diagonal
// block of every supernode is inverted on the host, so that a whole supernode costs two
// dependent steps whatever its width:
//     A:  y_t = b_t - T[t, off-block] x          (plain sparse row gathers)
//     B:  x_t = inv(T[t,t]) y_t                   (dense triangular rows, gathers from y)
// Both are instances of one primitive executed by the CUDA kernels,
//     xe[dst] = ( (init >= 0 ? xe[init] : 0) - sum_p val[p] * xe[col[p]] ) * scale,
// on an EXTENDED vector xe = [x (n rows) | y scratch (2 * ymax rows)].  Rows of one
// sub-level are independent; a barrier separates sub-levels.  Within a sub-level the rows
// are sorted by length and cut into SLICES of 32 >> g rows (2^g lanes per row, g from the
// row length); a slice is the unit of work of one warp and its entries are stored
// trip-major (sliced ELLPACK), so that the 32 lanes of a warp read 32 consecutive entries.  
// For the N=25 cavity factor this is 139-143 sub-levels instead of 1784 scalar dependency levels.
#pragma once
#include <stdint.h>
#include <vector>

namespace ocb {

struct Slice {        // up to 32 >> glog rows processed by ONE warp, 2^glog lanes per row
    int32_t ebase;    // first entry; entry (trip u, lane l) is at ebase + 32*u + l, where lane
                      // l = r * 2^glog + g holds entry u * 2^glog + g of the slice's row r
                      // (zero padded: SELL-32 layout, conflict-free shared-memory reads)
    int32_t trips;    // 32-wide trips = ceil(longest row / 2^glog)
    int32_t glog_nrows;  // glog | nrows << 8
    int32_t q0;       // first row
};

struct LuProgram {
    int64_t n = 0, n_ext = 0, ymax = 0;
    int64_t nnzL = 0, nnzU = 0;          // strictly lower / upper incl. diagonal (input factors)
    int64_t nent_actual = 0;             // program entries without the SELL padding
    int32_t nsub_L = 0, nsub_U = 0;      // sub-levels (barriers) of the two sweeps
    int32_t nsuper = 0, max_w = 0;       // supernodes, widest supernode
    std::vector<int32_t> sub_ptr;        // nsub + 1: slices of sub-level s are [sub_ptr[s], sub_ptr[s+1])
    std::vector<Slice> slices;
    std::vector<int32_t> dst, init;      // per row (indices into xe; init = -1: none)
    std::vector<double> scale;           // per row
    std::vector<int32_t> col;            // per padded entry (indices into xe)
    std::vector<double> val;
    uint64_t structure_id = 0;           // programs built from the same index arrays share it (0: none)
    int64_t nrows() const { return (int64_t)dst.size(); }
    int64_t nent() const { return (int64_t)col.size(); }
    int64_t nsub() const { return (int64_t)sub_ptr.size() - 1; }
};

// Build the program from CSR factors (L unit lower, diagonal optional; U upper with the
// diagonal).  Returns 0 or a negative ocb_status (message via set_error).
// transposed = true: the factors come from an LU factorisation of A^T handed over column-wise
// (A = U^T L^T): the LOWER factor (first three arrays, rows of U^T) carries the pivots and the
// UPPER factor (rows of L^T) has the unit diagonal.
int build_lu_program(int64_t n, const int32_t* Lrp, const int32_t* Lci, const double* Lva,
                     const int32_t* Urp, const int32_t* Uci, const double* Uva, int max_lanes,
                     bool transposed, bool merge, LuProgram* out, int wmax_cap = 0);
// wmax_cap > 0: no supernode wider than that (the "safe" layout of the residual guard: the
// inverse of a narrow triangular block is far better conditioned than that of a 512-row one).

// ---------------------------------------------------------------------------------
// PANEL form of the same program, for the all-columns-at-once executor (large n / wide blocks).
// Up to 8 rows of one sub-level with consecutive destinations (rows of one supernode) form a
// PANEL that shares ONE column list: the union of the rows' lists, values zero padded and stored
// interleaved (val[p][8]).  The executor then loads every x row once per panel and uses it for
// 8 rows (register blocking): 1 byte of x per FMA instead of 8.  Rows of the upper factor
// inside a supernode have identical lists (no padding); the lower factor pads ~1.4x.
// ---------------------------------------------------------------------------------
constexpr int PANEL_ROWS = 8;
struct Panel {
    int32_t cbase;    // first entry of the column list (pcol); values at pval[(cbase + p) * 8 + r]
    int32_t ncol;
    int32_t dst0;     // row r writes xe[dst0 + r]
    int32_t init0;    // row r starts from xe[init0 + r] (-1: from zero)
    int32_t nrows;    // 1..8
    int32_t pad[3];
};
struct PanelProgram {
    int64_t n = 0, n_ext = 0;
    std::vector<int32_t> sub_ptr;     // panels of sub-level s: [sub_ptr[s], sub_ptr[s+1]), longest first
    std::vector<Panel> panels;
    std::vector<double> scale;        // 8 per panel
    std::vector<int32_t> pcol;
    std::vector<double> pval;         // 8 per column entry
    int64_t entries_actual = 0;       // non-padding entries (= flops / 2 per right-hand side)
    int64_t nsub() const { return (int64_t)sub_ptr.size() - 1; }
};
// max_pad: a row joins a panel only while (union size * rows) <= max_pad * (sum of row lengths)
void build_panels(const LuProgram& P, double max_pad, PanelProgram* out);
void execute_panels_host(const PanelProgram& Q, const int32_t* perm_r, const int32_t* perm_c,
                         const double* b, double* x);

// Host execution of the program for ONE right-hand side (what the CUDA kernels do, serially):
// x = Pc U^-1 L^-1 Pr b.  Used by the residual guard of ocb_lu_pack_host_checked and by tests.
void execute_program_host(const LuProgram& P, const int32_t* perm_r, const int32_t* perm_c,
                          const double* b, double* x);

// merge = true: supernodes up to 64 rows wide (OCB_MERGE_W) are solved in ONE sub-level instead
// of two, in inverse-multiplied form  x_t = inv(T_tt) b_t - (inv(T_tt) T[t,off]) x  (as long as
// the product rows hold at most 1.3x the entries, OCB_MERGE_GROWTH).  The results are written to
// the y region and copied home one sub-level later by zero-length rows, off the critical path;
// a reader one sub-level after the block takes them from the y region.  N=25 cavity factor:
// 140 -> 90 sub-levels for 2 % more entries and 2 ms more host time.
//
// P = X * T for a w x w triangular X and a dense w x m T (row-major).  host_dense.cpp.
void tri_times_dense(const double* X, const double* T, double* P, int w, int m, bool upper);
// X = inverse of the w x w triangular D (row-major; X zero on entry).  host_dense.cpp.
void tri_inverse(const double* D, double* X, int w, bool upper, bool unit);
// C[M x N] += alpha * A[M x K] * B[K x N], row-major with leading dimensions.  host_dense.cpp.
void gemm_acc(int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C, int ldc,
              double alpha);

}  // namespace ocb
