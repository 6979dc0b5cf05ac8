// Numeric-only sparse LU refactorisation with symbolic reuse (host side).
//
// Every shifted saddle-point matrix of a run, (F_k^T + p_j M^T, J^T; J, 0), has the SAME
// sparsity pattern (reference solve_dae_ric.py:147,173 builds F_k from M, A, N(t) and the step
// length only; proj_ric_utils.py:108-111 adds the shift), and SuperLU - threshold pivoting with
// a preference for the diagonal - picks the same pivots for all of them.  The first
// factorisation of a pattern therefore fixes the two permutations; everything that depends on
// the pattern and the pivot order only is computed ONCE here,
//     elimination tree, supernodes, front structures, assembly and extend-add index maps,
//     the CSR structure of both factors,
// and every further matrix of the pattern costs one multifrontal pass over dense fronts with
// static pivots: no ordering, no symbolic analysis, no pivot search, no format conversion.
// The caller's residual guard (ocb_lu_pack_host_checked) decides whether the static pivots
// were good enough for a given matrix; if not, that matrix goes through SuperLU again.
//
// Structure: the pattern of C = P A Q is symmetrised (C + C^T) - the cavity / channel systems
// are structurally symmetric anyway - so that L and U^T share one structure, as in the
// classical "symmetric pattern, unsymmetric values" multifrontal method.  Fronts are dense
// row-major m x m arrays; the first w pivots of a front are eliminated by a blocked
// right-looking LU without pivoting whose trailing update is a register-blocked GEMM.  With
// row-major fronts both factors leave the front as contiguous row segments:
//     row k of U  = F[k, k:m],        row R[i] of L (columns of this supernode) = F[i, 0:w].
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../include/optconpy_b200.h"
#include "lu_program.h"

namespace ocb {
void set_error(const char* fmt, ...);

namespace {

inline void gemm_sub(int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* Cm, int ldc) {
    gemm_acc(M, N, K, A, lda, B, ldb, Cm, ldc, -1.0);      // host_dense.cpp
}

// First w pivots of the row-major m x m front F (no pivoting).  On return
//   F[k, k:m] = row k of U (k < w),   F[i, 0:min(i,w)] = multipliers (unit lower L),
//   F[w:m, w:m] = Schur complement.
// Blocked right-looking: per block of <= 32 pivots the diagonal block is factorised in place,
// its two triangular factors are inverted (32^3 / 3 flops each), and the row panel
// U12 = inv(L11) A12, the column panel L21 = A21 inv(U11) and the trailing update all run
// through the register-blocked GEMM (the substitution loops they replace ran at 1-2 GFLOP/s and
// took 2/3 of the time of the large fronts).
// Returns the index of the first pivot with |pivot| <= tiny, or -1.  work: >= 3*32*32 + 32*m.
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target_clones("avx512f", "avx2,fma", "default")))
#endif
int partial_lu(double* F, int m, int w, double tiny, double* work) {
    constexpr int NB = 32;
    double* nLi = work;                 // -(inv(L11) - I), b x b row-major (strictly lower)
    double* nUi = work + NB * NB;       // -inv(U11), b x b (upper)
    double* T = work + 2 * NB * NB;     // copy of the column panel, rows x b
    for (int kb = 0; kb < w; kb += NB) {
        const int b = std::min(NB, w - kb);
        const int ke = kb + b;
        // diagonal block, in place
        for (int k = kb; k < ke; ++k) {
            const double* rk = F + (size_t)k * m;
            const double piv = rk[k];
            if (!(fabs(piv) > tiny)) return k;
            const double ip = 1.0 / piv;
            for (int i = k + 1; i < ke; ++i) {
                double* ri = F + (size_t)i * m;
                const double l = ri[k] * ip;
                ri[k] = l;
                if (l == 0.0) continue;
                for (int j = k + 1; j < ke; ++j) ri[j] -= l * rk[j];
            }
        }
        if (ke == m) break;
        if (b <= 4) {
            // narrow block: plain substitution
            for (int k = kb; k < ke; ++k) {
                const double* rk = F + (size_t)k * m;
                for (int i = k + 1; i < ke; ++i) {
                    double* ri = F + (size_t)i * m;
                    const double l = ri[k];
                    if (l == 0.0) continue;
                    for (int j = ke; j < m; ++j) ri[j] -= l * rk[j];
                }
            }
            for (int i = ke; i < m; ++i) {
                double* ri = F + (size_t)i * m;
                for (int k = kb; k < ke; ++k) {
                    const double* rk = F + (size_t)k * m;
                    const double l = ri[k] / rk[k];
                    ri[k] = l;
                    for (int j = k + 1; j < ke; ++j) ri[j] -= l * rk[j];
                }
            }
        } else {
            // nLi = I - inv(L11): row i of inv(L) = e_i - sum_{k<i} L[i,k] inv(L)[k,:]
            for (int i = 0; i < b; ++i) {
                double* xi = nLi + i * b;
                for (int j = 0; j < b; ++j) xi[j] = 0.0;
                const double* li = F + (size_t)(kb + i) * m + kb;
                for (int k = 0; k < i; ++k) {
                    const double d = li[k];
                    if (d == 0.0) continue;
                    const double* xk = nLi + k * b;
                    xi[k] += d;                               // -(e_k part) negated: + L[i,k]
                    for (int j = 0; j < k; ++j) xi[j] -= d * xk[j];
                }
            }
            // nUi = -inv(U11): row i of inv(U) = (e_i - sum_{k>i} U[i,k] inv(U)[k,:]) / U[i,i]
            for (int i = b - 1; i >= 0; --i) {
                double* xi = nUi + i * b;
                for (int j = 0; j < b; ++j) xi[j] = 0.0;
                const double* ui = F + (size_t)(kb + i) * m + kb;
                xi[i] = -1.0;
                for (int k = i + 1; k < b; ++k) {
                    const double d = ui[k];
                    if (d == 0.0) continue;
                    const double* xk = nUi + k * b;
                    for (int j = k; j < b; ++j) xi[j] -= d * xk[j];
                }
                const double ip = 1.0 / ui[i];
                for (int j = i; j < b; ++j) xi[j] *= ip;
            }
            // U12 = inv(L11) A12 = A12 - nLi A12, in place from the bottom tile up (a tile only
            // reads rows above itself or, inside the tile, through zero coefficients)
            for (int i = (b - 1) & ~3; i >= 0; i -= 4) {
                const int rows = std::min(4, b - i);
                if (rows == 4) {
                    gemm_sub(4, m - ke, i + 3, nLi + i * b, b, F + (size_t)kb * m + ke, m,
                             F + (size_t)(kb + i) * m + ke, m);
                } else {        // the kernel's single-row path updates in place: one row at a time
                    for (int r = i + rows - 1; r >= std::max(i, 1); --r)
                        gemm_sub(1, m - ke, r, nLi + r * b, b, F + (size_t)kb * m + ke, m,
                                 F + (size_t)(kb + r) * m + ke, m);
                }
            }
            // L21 = A21 inv(U11) = 0 - T nUi
            const int rows = m - ke;
            for (int i = 0; i < rows; ++i) {
                double* ri = F + (size_t)(ke + i) * m + kb;
                double* ti = T + (size_t)i * b;
                for (int j = 0; j < b; ++j) { ti[j] = ri[j]; ri[j] = 0.0; }
            }
            gemm_sub(rows, b, b, T, b, nUi, b, F + (size_t)ke * m + kb, m);
        }
        // trailing update
        gemm_sub(m - ke, m - ke, b, F + (size_t)ke * m + kb, m, F + (size_t)kb * m + ke, m,
                 F + (size_t)ke * m + ke, m);
    }
    return -1;
}

}  // namespace
}  // namespace ocb

// ---------------------------------------------------------------------------------------------
// Nested-dissection ordering by breadth-first level structures (George's automatic nested
// dissection with thinned separators).  The solve kernels are bound by the NUMBER of dependent
// sub-levels of the gather program, i.e. by the height of the supernodal elimination tree: minimum
// degree keeps the fill low but grows tall trees (N=25 cavity saddle matrix: 90 sub-levels); this
// ordering gives 37 sub-levels AND 13 % less fill on the same matrix (n = 22k: 130 -> 69, -24 %).
//   set S:  components are ordered one after the other; a connected set is cut at the level of a
//   BFS from a pseudo-peripheral node that balances the halves; separator nodes without a
//   neighbour on one side move to the other side; order(S) = order(side 0), order(side 1), separator.
// ---------------------------------------------------------------------------------------------
namespace ocb {
namespace {

struct NdWork {
    const int32_t* ap;
    const int32_t* ai;
    int leaf;
    std::vector<int32_t> tag;       // current set id of a node (-1: already ordered / other set)
    std::vector<int32_t> dist, queue, side;
    int32_t next_tag = 0;
    int32_t* out;
    int64_t nout = 0;
};

// BFS inside the set tagged `tg`, from `start`; returns the last node reached, fills dist for the
// reached nodes and W.queue with them in BFS order
int32_t nd_bfs(NdWork& W, int32_t tg, int32_t start, int32_t* nreached, int32_t* depth) {
    W.queue.clear();
    W.queue.push_back(start);
    W.dist[start] = 0;
    size_t head = 0;
    // visited marker: dist >= 0 while the node's tag is tg and it sits in the queue; reset by the caller
    while (head < W.queue.size()) {
        const int32_t v = W.queue[head++];
        for (int32_t p = W.ap[v]; p < W.ap[v + 1]; ++p) {
            const int32_t u = W.ai[p];
            if (W.tag[u] == tg && W.dist[u] < 0) {
                W.dist[u] = W.dist[v] + 1;
                W.queue.push_back(u);
            }
        }
    }
    *nreached = (int32_t)W.queue.size();
    *depth = W.dist[W.queue.back()];
    return W.queue.back();
}

void nd_order(NdWork& W, std::vector<int32_t>& nodes) {
    const int32_t m = (int32_t)nodes.size();
    if (m == 0) return;
    if (m <= W.leaf) {
        for (int32_t v : nodes) { W.out[W.nout++] = v; W.tag[v] = -1; }
        return;
    }
    const int32_t tg = W.next_tag++;
    for (int32_t v : nodes) { W.tag[v] = tg; W.dist[v] = -1; }
    // connected components: the first one is treated below, the others recursively
    int32_t reached = 0, depth = 0;
    int32_t far = nd_bfs(W, tg, nodes[0], &reached, &depth);
    if (reached < m) {
        std::vector<int32_t> comp(W.queue.begin(), W.queue.end()), rest;
        rest.reserve(m - reached);
        for (int32_t v : nodes)
            if (W.dist[v] < 0) rest.push_back(v);
        for (int32_t v : nodes) W.tag[v] = -1;
        nd_order(W, comp);
        nd_order(W, rest);
        return;
    }
    // pseudo-peripheral start
    int32_t start = nodes[0];
    for (int it = 0; it < 4 && far != start; ++it) {
        start = far;
        for (int32_t v : nodes) W.dist[v] = -1;
        const int32_t d0 = depth;
        far = nd_bfs(W, tg, start, &reached, &depth);
        if (depth <= d0 && it > 0) break;
    }
    if (depth < 2) {     // (nearly) complete graph: nothing to dissect
        for (int32_t v : nodes) { W.out[W.nout++] = v; W.tag[v] = -1; }
        return;
    }
    // the level that balances the two sides
    std::vector<int32_t> cnt(depth + 1, 0);
    for (int32_t v : nodes) ++cnt[W.dist[v]];
    int32_t lev = 1;
    {
        double bestd = 1e300, cum = 0.0;
        for (int32_t l = 0; l <= depth; ++l) {
            cum += cnt[l];
            const double d = fabs(cum - 0.5 * cnt[l] - 0.5 * m);
            if (d < bestd && l >= 1 && l <= depth - 1) { bestd = d; lev = l; }
        }
    }
    for (int32_t v : nodes) W.side[v] = W.dist[v] < lev ? 0 : (W.dist[v] > lev ? 1 : 2);
    // thin the separator
    for (int pass = 0; pass < 2; ++pass) {
        const int other = pass == 0 ? 1 : 0;
        for (int32_t v : nodes) {
            if (W.side[v] != 2) continue;
            bool touches = false;
            for (int32_t p = W.ap[v]; p < W.ap[v + 1] && !touches; ++p) {
                const int32_t u = W.ai[p];
                touches = W.tag[u] == tg && W.side[u] == other;
            }
            if (!touches) W.side[v] = pass == 0 ? 0 : 1;
        }
    }
    std::vector<int32_t> s0, s1, sep;
    for (int32_t v : nodes) (W.side[v] == 0 ? s0 : (W.side[v] == 1 ? s1 : sep)).push_back(v);
    for (int32_t v : nodes) W.tag[v] = -1;
    std::vector<int32_t>().swap(nodes);
    nd_order(W, s0);
    nd_order(W, s1);
    for (int32_t v : sep) W.out[W.nout++] = v;
}

}  // namespace
}  // namespace ocb

extern "C" int ocb_order_nd(int64_t n, const int32_t* adj_rowptr, const int32_t* adj_colidx, int64_t leaf,
                            int32_t* order_out) {
    using namespace ocb;
    if (n < 0 || (n > 0 && (!adj_rowptr || !order_out)) || n >= INT32_MAX) {
        set_error("order_nd: bad argument");
        return OCB_ERR_ARG;
    }
    if (n == 0) return OCB_OK;
    for (int64_t i = 0; i < n; ++i)
        for (int32_t p = adj_rowptr[i]; p < adj_rowptr[i + 1]; ++p)
            if (adj_colidx[p] < 0 || adj_colidx[p] >= n) {
                set_error("order_nd: column index out of range");
                return OCB_ERR_ARG;
            }
    NdWork W;
    W.ap = adj_rowptr;
    W.ai = adj_colidx;
    W.leaf = (int)std::max<int64_t>(leaf, 1);
    W.tag.assign(n, -1);
    W.dist.assign(n, -1);
    W.side.assign(n, 0);
    W.queue.reserve(n);
    W.out = order_out;
    std::vector<int32_t> all(n);
    std::iota(all.begin(), all.end(), 0);
    nd_order(W, all);
    if (W.nout != n) {
        set_error("order_nd: internal error (%lld of %lld nodes ordered)", (long long)W.nout, (long long)n);
        return OCB_ERR_ARG;
    }
    return OCB_OK;
}

// Constrained ordering for saddle-point patterns (see _lu_worker._delay_zero_diagonals, whose
// Python loop this replaces for large systems: 0.7 s at n = 89 402): a node with a zero diagonal is
// delayed until right after the first of its neighbours with a non-zero diagonal has been
// eliminated; zero-diagonal nodes without such a neighbour go last, in their original order.
extern "C" int ocb_order_delay_zero_diagonals(int64_t n, const int32_t* adj_rowptr, const int32_t* adj_colidx,
                                              const uint8_t* diag_is_zero, const int32_t* order_in,
                                              int32_t* order_out) {
    using namespace ocb;
    if (n < 0 || (n > 0 && (!adj_rowptr || !diag_is_zero || !order_in || !order_out))) {
        set_error("order_delay_zero_diagonals: bad argument");
        return OCB_ERR_ARG;
    }
    std::vector<uint8_t> touched(n, 0), pending(n, 0), seen(n, 0);
    std::vector<int32_t> rel;
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t v = order_in[i];
        if (v < 0 || v >= n || seen[v]) {
            set_error("order_delay_zero_diagonals: order_in is not a permutation");
            return OCB_ERR_ARG;
        }
        seen[v] = 1;
        if (diag_is_zero[v]) {
            if (touched[v]) order_out[m++] = v; else pending[v] = 1;
            continue;
        }
        order_out[m++] = v;
        rel.clear();
        for (int32_t p = adj_rowptr[v]; p < adj_rowptr[v + 1]; ++p) {
            const int32_t u = adj_colidx[p];
            if (u < 0 || u >= n || !diag_is_zero[u]) continue;
            touched[u] = 1;
            if (pending[u]) { pending[u] = 0; rel.push_back(u); }
        }
        if (!rel.empty()) {
            std::sort(rel.begin(), rel.end());      // released nodes in index order (np.unique)
            rel.erase(std::unique(rel.begin(), rel.end()), rel.end());
            for (int32_t u : rel) order_out[m++] = u;
        }
    }
    for (int64_t i = 0; i < n; ++i) {               // the rest, in the order of order_in
        const int32_t v = order_in[i];
        if (pending[v]) { pending[v] = 0; order_out[m++] = v; }
    }
    if (m != n) {
        set_error("order_delay_zero_diagonals: order_in is not a permutation (%lld of %lld placed)",
                  (long long)m, (long long)n);
        return OCB_ERR_ARG;
    }
    return OCB_OK;
}

struct ocb_refactor {
    int64_t n = 0;
    int64_t nnzA = 0;
    std::vector<int32_t> perm_r, perm_c;      // final permutations (input ones composed with the postorder)
    // supernodes
    int32_t nsn = 0;
    std::vector<int32_t> sn_start;            // nsn + 1
    std::vector<int64_t> rs_ptr;              // nsn + 1: rows of supernode J are rows[rs_ptr[J] .. rs_ptr[J+1])
    std::vector<int32_t> rows;
    std::vector<int32_t> nchild;              // children per supernode (their blocks are on top of the stack)
    std::vector<int32_t> rel;                 // per entry of rows beyond the diagonal block: index in the parent's rows
    std::vector<int64_t> lseg;                // per entry of rows: start of this row's segment in Lva
    // assembly: entries of A grouped by supernode
    std::vector<int64_t> asm_ptr;             // nsn + 1
    std::vector<int32_t> asm_src;
    std::vector<int64_t> asm_dst;             // offset in the front
    // factor structure (CSR)
    std::vector<int32_t> Lrp, Lci, Urp, Uci;
    int32_t max_front = 0;
    int64_t stack_peak = 0;
    double flops = 0.0;
    // work space
    std::vector<double> front, stack, work;
};

extern "C" {

int ocb_refactor_create(ocb_refactor** out, int64_t n, const int32_t* A_colptr, const int32_t* A_rowidx,
                        const int32_t* perm_r, const int32_t* perm_c) {
    using namespace ocb;
    if (!out || n < 0 || (n > 0 && (!A_colptr || !A_rowidx || !perm_r || !perm_c))) {
        set_error("refactor_create: bad argument");
        return OCB_ERR_ARG;
    }
    ocb_refactor* R = new ocb_refactor();
    R->n = n;
    R->nnzA = n > 0 ? A_colptr[n] : 0;
    const int64_t nnzA = R->nnzA;
    if (nnzA >= INT32_MAX / 4) {
        delete R;
        set_error("refactor_create: matrix too large for int32 indices");
        return OCB_ERR_ARG;
    }
    {   // the inputs must be permutations
        std::vector<char> seen(n, 0);
        for (int pass = 0; pass < 2; ++pass) {
            const int32_t* p = pass ? perm_c : perm_r;
            std::fill(seen.begin(), seen.end(), 0);
            for (int64_t i = 0; i < n; ++i) {
                if (p[i] < 0 || p[i] >= n || seen[p[i]]) {
                    delete R;
                    set_error("refactor_create: perm_%c is not a permutation", pass ? 'c' : 'r');
                    return OCB_ERR_ARG;
                }
                seen[p[i]] = 1;
            }
        }
    }
    std::vector<int32_t> pr(perm_r, perm_r + n), pc(perm_c, perm_c + n);
    std::vector<int32_t> parent(n), adj_ptr, adj;
    // symmetric adjacency of C = P A Q (C[pr[i], pc[j]] = A[i, j]) without the diagonal, sorted
    auto build_adjacency = [&]() {
        std::vector<int32_t> cnt(n + 1, 0);
        for (int64_t j = 0; j < n; ++j)
            for (int32_t p = A_colptr[j]; p < A_colptr[j + 1]; ++p) {
                const int32_t r = pr[A_rowidx[p]], c = pc[j];
                if (r != c) { ++cnt[r + 1]; ++cnt[c + 1]; }
            }
        adj_ptr.assign(n + 1, 0);
        for (int64_t i = 0; i < n; ++i) adj_ptr[i + 1] = adj_ptr[i] + cnt[i + 1];
        std::vector<int32_t> raw(adj_ptr[n]), fill(adj_ptr.begin(), adj_ptr.end() - 1);
        for (int64_t j = 0; j < n; ++j)
            for (int32_t p = A_colptr[j]; p < A_colptr[j + 1]; ++p) {
                const int32_t r = pr[A_rowidx[p]], c = pc[j];
                if (r != c) { raw[fill[r]++] = c; raw[fill[c]++] = r; }
            }
        adj.clear();
        std::vector<int32_t> np(n + 1, 0);
        for (int64_t i = 0; i < n; ++i) {
            std::sort(raw.begin() + adj_ptr[i], raw.begin() + adj_ptr[i + 1]);
            int32_t last = -1;
            for (int32_t p = adj_ptr[i]; p < adj_ptr[i + 1]; ++p)
                if (raw[p] != last) { adj.push_back(raw[p]); last = raw[p]; }
            np[i + 1] = (int32_t)adj.size();
        }
        adj_ptr.swap(np);
    };
    auto build_etree = [&]() {     // Liu's algorithm with path compression
        std::vector<int32_t> anc(n, -1);
        for (int64_t i = 0; i < n; ++i) {
            parent[i] = -1;
            for (int32_t p = adj_ptr[i]; p < adj_ptr[i + 1] && adj[p] < i; ++p) {
                int32_t k = adj[p];
                while (k != -1 && k < i) {
                    const int32_t nx = anc[k];
                    anc[k] = (int32_t)i;
                    if (nx == -1) parent[k] = (int32_t)i;
                    k = nx;
                }
            }
        }
    };
    build_adjacency();
    build_etree();
    {   // postorder (children in increasing order), then relabel: subtrees become contiguous
        std::vector<int32_t> head(n, -1), next(n, -1), post(n), stack;
        for (int64_t i = n - 1; i >= 0; --i)
            if (parent[i] >= 0) { next[i] = head[parent[i]]; head[parent[i]] = (int32_t)i; }
        int32_t k = 0;
        for (int64_t root = 0; root < n; ++root) {
            if (parent[root] >= 0) continue;
            stack.push_back((int32_t)root);
            while (!stack.empty()) {
                const int32_t v = stack.back();
                const int32_t c = head[v];
                if (c >= 0) { head[v] = next[c]; stack.push_back(c); }
                else { post[v] = k++; stack.pop_back(); }
            }
        }
        for (int64_t i = 0; i < n; ++i) { pr[i] = post[pr[i]]; pc[i] = post[pc[i]]; }
        build_adjacency();
        build_etree();
    }
    R->perm_r = pr;
    R->perm_c = pc;
    // column structures of L (rows >= j), children merged into the parent
    std::vector<std::vector<int32_t>> cs(n);
    {
        std::vector<int32_t> mark(n, -1);
        std::vector<std::vector<int32_t>> kids(n);
        for (int64_t j = 0; j < n; ++j)
            if (parent[j] >= 0) kids[parent[j]].push_back((int32_t)j);
        for (int64_t j = 0; j < n; ++j) {
            std::vector<int32_t>& s = cs[j];
            s.push_back((int32_t)j);
            mark[j] = (int32_t)j;
            for (int32_t p = adj_ptr[j]; p < adj_ptr[j + 1]; ++p)
                if (adj[p] > j && mark[adj[p]] != j) { mark[adj[p]] = (int32_t)j; s.push_back(adj[p]); }
            for (int32_t c : kids[j]) {
                for (int32_t r : cs[c])
                    if (r > j && mark[r] != j) { mark[r] = (int32_t)j; s.push_back(r); }
            }
            std::sort(s.begin(), s.end());
        }
    }
    // (fundamental) supernodes: column j+1 joins column j if it is j's parent and has the same structure below
    R->sn_start.push_back(0);
    for (int64_t j = 1; j < n; ++j)
        if (!(parent[j - 1] == j && cs[j].size() + 1 == cs[j - 1].size())) R->sn_start.push_back((int32_t)j);
    if (n > 0) R->sn_start.push_back((int32_t)n);
    R->nsn = (int32_t)R->sn_start.size() - 1;
    const int32_t nsn = R->nsn;
    std::vector<int32_t> sn_of(n);
    R->rs_ptr.assign(nsn + 1, 0);
    for (int32_t J = 0; J < nsn; ++J) {
        for (int32_t j = R->sn_start[J]; j < R->sn_start[J + 1]; ++j) sn_of[j] = J;
        R->rs_ptr[J + 1] = R->rs_ptr[J] + (int64_t)cs[R->sn_start[J]].size();
    }
    R->rows.resize(R->rs_ptr[nsn]);
    for (int32_t J = 0; J < nsn; ++J)
        std::copy(cs[R->sn_start[J]].begin(), cs[R->sn_start[J]].end(), R->rows.begin() + R->rs_ptr[J]);
    std::vector<std::vector<int32_t>>().swap(cs);
    // supernodal tree, relative indices, stack depth, flops
    R->nchild.assign(nsn, 0);
    R->rel.assign(R->rows.size(), -1);
    std::vector<int32_t> sparent(nsn, -1);
    for (int32_t J = 0; J < nsn; ++J) {
        const int32_t w = R->sn_start[J + 1] - R->sn_start[J];
        const int64_t a = R->rs_ptr[J], m = R->rs_ptr[J + 1] - a;
        R->max_front = std::max<int32_t>(R->max_front, (int32_t)m);
        for (int32_t k = 0; k < w; ++k) R->flops += 2.0 * (double)(m - k - 1) * (double)(m - k - 1) + (double)(m - k - 1);
        if (m == w) continue;
        const int32_t Pn = sn_of[R->rows[a + w]];
        sparent[J] = Pn;
        ++R->nchild[Pn];
        const int64_t pa = R->rs_ptr[Pn], pm = R->rs_ptr[Pn + 1] - pa;
        int64_t q = 0;
        for (int64_t i = w; i < m; ++i) {
            const int32_t r = R->rows[a + i];
            while (q < pm && R->rows[pa + q] < r) ++q;
            if (q >= pm || R->rows[pa + q] != r) {
                delete R;
                set_error("refactor_create: internal error (front structure not nested)");
                return OCB_ERR_ARG;
            }
            R->rel[a + i] = (int32_t)q;
        }
    }
    {   // the contribution blocks of the children of J must be the top of the stack when J is
        // assembled: true for a postordered tree; verify, and record the peak
        std::vector<int32_t> st;
        int64_t cur = 0;
        for (int32_t J = 0; J < nsn; ++J) {
            for (int32_t c = 0; c < R->nchild[J]; ++c) {
                if (st.empty() || sparent[st.back()] != J) {
                    delete R;
                    set_error("refactor_create: internal error (supernodes not in postorder)");
                    return OCB_ERR_ARG;
                }
                const int32_t K = st.back();
                st.pop_back();
                const int64_t mk = R->rs_ptr[K + 1] - R->rs_ptr[K] - (R->sn_start[K + 1] - R->sn_start[K]);
                cur -= mk * mk;
            }
            const int64_t mj = R->rs_ptr[J + 1] - R->rs_ptr[J] - (R->sn_start[J + 1] - R->sn_start[J]);
            if (mj > 0) {
                st.push_back(J);
                cur += mj * mj;
                R->stack_peak = std::max(R->stack_peak, cur);
            }
        }
    }
    // CSR structure of U (row j0+k: rows[a+k .. a+m)) and of the unit lower L (diagonal stored)
    R->Urp.assign(n + 1, 0);
    R->Lrp.assign(n + 1, 0);
    std::vector<int32_t> lcnt(n, 1);
    for (int32_t J = 0; J < nsn; ++J) {
        const int32_t j0 = R->sn_start[J], w = R->sn_start[J + 1] - j0;
        const int64_t a = R->rs_ptr[J], m = R->rs_ptr[J + 1] - a;
        for (int32_t k = 0; k < w; ++k) R->Urp[j0 + k + 1] = (int32_t)(m - k);
        for (int64_t i = 0; i < m; ++i) lcnt[R->rows[a + i]] += (int32_t)std::min<int64_t>(i, w);
    }
    for (int64_t i = 0; i < n; ++i) {
        R->Urp[i + 1] += R->Urp[i];
        R->Lrp[i + 1] = R->Lrp[i] + lcnt[i];
    }
    R->Uci.resize(R->Urp[n]);
    R->Lci.resize(R->Lrp[n]);
    R->lseg.assign(R->rows.size(), 0);
    std::vector<int32_t> lfill(R->Lrp.begin(), R->Lrp.end() - 1);
    for (int32_t J = 0; J < nsn; ++J) {
        const int32_t j0 = R->sn_start[J], w = R->sn_start[J + 1] - j0;
        const int64_t a = R->rs_ptr[J], m = R->rs_ptr[J + 1] - a;
        for (int32_t k = 0; k < w; ++k)
            std::copy(R->rows.begin() + a + k, R->rows.begin() + a + m, R->Uci.begin() + R->Urp[j0 + k]);
        for (int64_t i = 0; i < m; ++i) {
            const int32_t r = R->rows[a + i];
            const int32_t len = (int32_t)std::min<int64_t>(i, w);
            R->lseg[a + i] = lfill[r];
            for (int32_t k = 0; k < len; ++k) R->Lci[lfill[r]++] = j0 + k;
        }
    }
    for (int64_t i = 0; i < n; ++i) R->Lci[lfill[i]++] = (int32_t)i;      // unit diagonal, last in its row
    // assembly map
    {
        std::vector<int64_t> cnt(nsn + 1, 0);
        for (int64_t j = 0; j < n; ++j)
            for (int32_t p = A_colptr[j]; p < A_colptr[j + 1]; ++p)
                ++cnt[sn_of[std::min(pr[A_rowidx[p]], pc[j])] + 1];
        R->asm_ptr.assign(nsn + 1, 0);
        for (int32_t J = 0; J < nsn; ++J) R->asm_ptr[J + 1] = R->asm_ptr[J] + cnt[J + 1];
        R->asm_src.resize(nnzA);
        R->asm_dst.resize(nnzA);
        std::vector<int64_t> fill(R->asm_ptr.begin(), R->asm_ptr.end() - 1);
        for (int64_t j = 0; j < n; ++j)
            for (int32_t p = A_colptr[j]; p < A_colptr[j + 1]; ++p) {
                const int32_t r = pr[A_rowidx[p]], c = pc[j];
                const int32_t J = sn_of[std::min(r, c)];
                const int32_t j0 = R->sn_start[J], j1 = R->sn_start[J + 1];
                const int64_t a = R->rs_ptr[J], m = R->rs_ptr[J + 1] - a;
                auto pos = [&](int32_t x) -> int64_t {
                    if (x < j1) return x - j0;
                    const int32_t* b = R->rows.data() + a;
                    const int32_t* it = std::lower_bound(b, b + m, x);
                    return (it != b + m && *it == x) ? it - b : -1;
                };
                const int64_t lr = pos(r), lc = pos(c);
                if (lr < 0 || lc < 0) {
                    delete R;
                    set_error("refactor_create: internal error (entry outside its front)");
                    return OCB_ERR_ARG;
                }
                const int64_t f = fill[J]++;
                R->asm_src[f] = p;
                R->asm_dst[f] = lr * m + lc;
            }
    }
    R->front.resize((size_t)R->max_front * R->max_front);
    R->stack.resize((size_t)R->stack_peak);
    R->work.resize((size_t)3 * 32 * 32 + (size_t)32 * R->max_front);
    *out = R;
    return OCB_OK;
}

int ocb_refactor_destroy(ocb_refactor* R) {
    delete R;
    return OCB_OK;
}

// info8: n, nnz(L) incl. unit diagonal, nnz(U), supernodes, largest front, peak stack entries,
// flops of one numeric pass, nnz(A)
int ocb_refactor_info(const ocb_refactor* R, int64_t* info8) {
    if (!R || !info8) return OCB_ERR_ARG;
    info8[0] = R->n;
    info8[1] = R->n ? R->Lrp[R->n] : 0;
    info8[2] = R->n ? R->Urp[R->n] : 0;
    info8[3] = R->nsn;
    info8[4] = R->max_front;
    info8[5] = R->stack_peak;
    info8[6] = (int64_t)R->flops;
    info8[7] = R->nnzA;
    return OCB_OK;
}

// Structure of the factors (fixed for the life of the handle): CSR index arrays of the unit
// lower L (diagonal stored last in its row) and of U (diagonal first), and the permutations
// with  (P A Q)[perm_r[i], perm_c[j]] = A[i, j] = (L U)[perm_r[i], perm_c[j]].
int ocb_refactor_structure(const ocb_refactor* R, int32_t* L_rowptr, int32_t* L_colidx, int32_t* U_rowptr,
                           int32_t* U_colidx, int32_t* perm_r, int32_t* perm_c) {
    if (!R) return OCB_ERR_ARG;
    const int64_t n = R->n;
    if (n == 0) return OCB_OK;
    if (L_rowptr) memcpy(L_rowptr, R->Lrp.data(), (size_t)(n + 1) * 4);
    if (L_colidx) memcpy(L_colidx, R->Lci.data(), R->Lci.size() * 4);
    if (U_rowptr) memcpy(U_rowptr, R->Urp.data(), (size_t)(n + 1) * 4);
    if (U_colidx) memcpy(U_colidx, R->Uci.data(), R->Uci.size() * 4);
    if (perm_r) memcpy(perm_r, R->perm_r.data(), (size_t)n * 4);
    if (perm_c) memcpy(perm_c, R->perm_c.data(), (size_t)n * 4);
    return OCB_OK;
}

// One numeric pass: A_vals in the order of the CSC arrays given to ocb_refactor_create.
// OCB_ERR_SINGULAR if a static pivot is (numerically) zero - the caller then falls back to a
// pivoting factorisation.  Not re-entrant on one handle (the work space belongs to it).
int ocb_refactor_numeric(ocb_refactor* R, const double* A_vals, double* L_vals, double* U_vals) {
    using namespace ocb;
    if (!R || (R->n > 0 && (!A_vals || !L_vals || !U_vals))) {
        set_error("refactor_numeric: bad argument");
        return OCB_ERR_ARG;
    }
    const int64_t n = R->n;
    if (n == 0) return OCB_OK;
    double amax = 0.0;
    for (int64_t p = 0; p < R->nnzA; ++p) amax = std::max(amax, fabs(A_vals[p]));
    if (!(amax == amax) || amax > 1e300) {
        set_error("refactor_numeric: matrix has non-finite entries");
        return OCB_ERR_ARG;
    }
    const double tiny = amax * 1e-300;
    double* F = R->front.data();
    double* S = R->stack.data();
    int64_t top = 0;                       // entries on the stack
    // (rows beyond the diagonal block)^2 of the blocks on the stack, innermost last
    std::vector<int32_t> onstack;
    onstack.reserve(256);
    for (int32_t J = 0; J < R->nsn; ++J) {
        const int32_t j0 = R->sn_start[J], w = R->sn_start[J + 1] - j0;
        const int64_t a = R->rs_ptr[J];
        const int m = (int)(R->rs_ptr[J + 1] - a);
        memset(F, 0, (size_t)m * m * sizeof(double));
        for (int64_t e = R->asm_ptr[J]; e < R->asm_ptr[J + 1]; ++e) F[R->asm_dst[e]] += A_vals[R->asm_src[e]];
        for (int32_t c = 0; c < R->nchild[J]; ++c) {       // extend-add, last child first
            const int32_t K = onstack.back();
            onstack.pop_back();
            const int32_t wk = R->sn_start[K + 1] - R->sn_start[K];
            const int64_t ka = R->rs_ptr[K] + wk;
            const int mk = (int)(R->rs_ptr[K + 1] - ka);
            top -= (int64_t)mk * mk;
            const double* cb = S + top;
            const int32_t* rel = R->rel.data() + ka;
            for (int i = 0; i < mk; ++i) {
                double* fr = F + (size_t)rel[i] * m;
                const double* cr = cb + (size_t)i * mk;
                for (int j = 0; j < mk; ++j) fr[rel[j]] += cr[j];
            }
        }
        const int bad = partial_lu(F, m, w, tiny, R->work.data());
        if (bad >= 0) {
            set_error("refactor_numeric: zero static pivot in column %d", j0 + bad);
            return OCB_ERR_SINGULAR;
        }
        for (int32_t k = 0; k < w; ++k)
            memcpy(U_vals + R->Urp[j0 + k], F + (size_t)k * m + k, (size_t)(m - k) * sizeof(double));
        for (int i = 1; i < m; ++i) {
            const int len = std::min(i, (int)w);
            memcpy(L_vals + R->lseg[a + i], F + (size_t)i * m, (size_t)len * sizeof(double));
        }
        const int mk = m - w;
        if (mk > 0) {
            double* cb = S + top;
            for (int i = 0; i < mk; ++i) memcpy(cb + (size_t)i * mk, F + (size_t)(w + i) * m + w, (size_t)mk * sizeof(double));
            top += (int64_t)mk * mk;
            onstack.push_back(J);
        }
    }
    for (int64_t i = 0; i < n; ++i) L_vals[R->Lrp[i + 1] - 1] = 1.0;
    return OCB_OK;
}

}  // extern "C"
