// Host-side construction of the gather program of a sparse LU solve (see lu_program.h).
// Pure host code: no CUDA calls, so it is unit-tested on CPU through
// ocb_lu_program_create / ocb_lu_program_export (tests/test_lu_program.py).
#include "common.cuh"
#include "lu_program.h"
#include <algorithm>
#include <numeric>
#include <string.h>
#include <stdlib.h>
#include <chrono>
#include <memory>
#include <mutex>
#include <stdio.h>

namespace ocb {

namespace {

static int wmax() {           // widest supernode (its inverse block has w(w+1)/2 entries)
    static int v = 0;
    if (v == 0) {
        const char* env = getenv("OCB_SUPERNODE_WMAX");
        v = env ? atoi(env) : 512;
        if (v < 1 || v > 512) v = 512;
    }
    return v;
}
// y-scratch rows one sub-level may use (two such regions exist): a supernode that finds its
// sub-level full moves to the next one.  1024 rows keep the extended vector of the column-panel
// kernel small (shared memory); large systems, whose vector lives in global memory anyway, get
// n/8 rows - with nested dissection thousands of supernodes are ready at the same time, and a
// small cap would spread them over many artificial sub-levels (N=100 cavity: 195 instead of ~100)
static int ycap_for(int64_t n) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("OCB_YCAP");
        forced = e ? atoi(e) : 0;
    }
    if (forced > 0) return forced;
    return n <= 16384 ? 1024 : (int)std::min<int64_t>(n / 8, 1 << 20);
}

struct BlockPlan {
    int32_t sA = -1;   // sub-level of the A rows (-1: block needs no work)
    int32_t yoff = -1; // offset into the y region (w > 1 only)
    bool merged = false;   // one-step block: x_t = inv(T_tt) b_t - (inv(T_tt) T[t,off]) x at sA
};

struct MergeRule {
    int max_w = 0;         // widest block solved in one step (0: never)
    double growth = 1.3;   // ... if its product rows hold at most growth x the entries of T[t,off]
};

struct Tri {
    const int32_t* rp;
    const int32_t* ci;
    const double* va;
    bool upper;
    bool unit;   // unit diagonal (stored entries on the diagonal are ignored)
};

inline int pow2ceil(int64_t v) {
    int g = 0;
    while ((1LL << g) < v) ++g;
    return g;
}

// supernodes: maximal runs of rows i, i+1, ... of U with  struct(U_i) \ {i} == struct(U_{i+1})
int find_supernodes(int64_t n, const int32_t* rp, const int32_t* ci, int wcap, std::vector<int32_t>* starts) {
    starts->clear();
    starts->push_back(0);
    int w = 1;
    for (int64_t i = 0; i + 1 < n; ++i) {
        const int32_t a0 = rp[i], a1 = rp[i + 1], b0 = rp[i + 1], b1 = rp[i + 2];
        bool same = false;
        if (a1 > a0 && ci[a0] == i && (a1 - a0 - 1) == (b1 - b0) && b1 > b0 && w < wcap)
            same = memcmp(ci + a0 + 1, ci + b0, (size_t)(b1 - b0) * sizeof(int32_t)) == 0;
        if (same) {
            ++w;
        } else {
            starts->push_back((int32_t)(i + 1));
            w = 1;
        }
    }
    if (n > 0) starts->push_back((int32_t)n);
    return OCB_OK;
}

// assign the A sub-level (and y offset) of every block of one triangular factor
int plan_factor(int64_t n, const Tri& T, const std::vector<int32_t>& starts, const MergeRule& rule,
                std::vector<BlockPlan>* plan, int32_t* nsub, int64_t* ymax) {
    const int YCAP = ycap_for(n);
    const int nb = (int)starts.size() - 1;
    plan->assign(nb, BlockPlan());
    std::vector<int32_t> done(n, -1);
    std::vector<int32_t> seen(rule.max_w > 0 ? n : 0, -1);   // last block that touched a column
    std::vector<int32_t> yuse;
    int32_t top = -1;
    for (int bb = 0; bb < nb; ++bb) {
        const int t = T.upper ? nb - 1 - bb : bb;
        const int32_t r0 = starts[t], r1 = starts[t + 1], w = r1 - r0;
        int32_t dep = -1;
        int64_t noff = 0;
        for (int32_t i = r0; i < r1; ++i) {
            bool diag = false;
            for (int32_t p = T.rp[i]; p < T.rp[i + 1]; ++p) {
                const int32_t c = T.ci[p];
                if (c < 0 || c >= n) {
                    set_error("factor has a column index out of range (row %d)", i);
                    return OCB_ERR_ARG;
                }
                if (c >= r0 && c < r1) {
                    if (T.upper ? c < i : c > i) {
                        set_error("%s has an entry on the wrong side of the diagonal (row %d col %d)",
                                  T.upper ? "U" : "L", i, c);
                        return OCB_ERR_ARG;
                    }
                    if (c == i) diag = T.va[p] != 0.0;
                    continue;
                }
                if (T.upper ? c < i : c > i) {
                    set_error("%s has an entry on the wrong side of the diagonal (row %d col %d)",
                              T.upper ? "U" : "L", i, c);
                    return OCB_ERR_ARG;
                }
                dep = std::max(dep, done[c]);
                ++noff;
            }
            if (!T.unit && !diag) {
                set_error("the %s factor has a zero pivot in row %d", T.upper ? "upper" : "lower", i);
                return OCB_ERR_SINGULAR;
            }
        }
        BlockPlan& bp = (*plan)[t];
        if (w > 1 && w <= rule.max_w) {
            // entries of the product rows inv(T_tt) T[t,off]: row k holds the union of the
            // off-block patterns of the rows it depends on (s <= k lower, s >= k upper)
            int64_t uni = 0, mcost = 0;
            for (int32_t kk = 0; kk < w; ++kk) {
                const int32_t i = T.upper ? r1 - 1 - kk : r0 + kk;
                for (int32_t p = T.rp[i]; p < T.rp[i + 1]; ++p) {
                    const int32_t c = T.ci[p];
                    if ((c < r0 || c >= r1) && seen[c] != t) { seen[c] = t; ++uni; }
                }
                mcost += uni;
            }
            bp.merged = (double)mcost <= rule.growth * (double)noff + 8.0 * w;
        }
        if (w == 1) {
            if (T.unit && noff == 0) continue;  // unit diagonal, nothing to subtract
            bp.sA = dep + 1;
            done[r0] = bp.sA;
            top = std::max(top, bp.sA);
        } else if (bp.merged) {
            // results go to the y region at sA and are copied home at sA + 1; a reader at
            // sA + 1 takes them from the y region, later ones from their home rows
            int32_t s = dep + 1;
            for (;;) {
                if ((int)yuse.size() <= s) yuse.resize(s + 1, 0);
                if (yuse[s] + w <= YCAP || yuse[s] == 0) break;
                ++s;
            }
            bp.sA = s;
            bp.yoff = yuse[s];
            yuse[s] += w;
            *ymax = std::max<int64_t>(*ymax, yuse[s]);
            for (int32_t i = r0; i < r1; ++i) done[i] = s;
            top = std::max(top, s + 1);
        } else {
            int32_t s = dep + 1;
            for (;;) {
                if ((int)yuse.size() <= s) yuse.resize(s + 1, 0);
                if (yuse[s] + w <= YCAP || yuse[s] == 0) break;
                ++s;
            }
            bp.sA = s;
            bp.yoff = yuse[s];
            yuse[s] += w;
            *ymax = std::max<int64_t>(*ymax, yuse[s]);
            for (int32_t i = r0; i < r1; ++i) done[i] = s + 1;
            top = std::max(top, s + 1);
        }
    }
    *nsub = top + 1;
    return OCB_OK;
}

// inverse of the w x w diagonal block of rows [r0, r1) (row-major, dense)
void invert_block(const Tri& T, int32_t r0, int32_t r1, std::vector<double>* Dbuf,
                  std::vector<double>* Xbuf) {
    const int w = r1 - r0;
    std::vector<double>& D = *Dbuf;
    std::vector<double>& X = *Xbuf;
    D.assign((size_t)w * w, 0.0);
    X.assign((size_t)w * w, 0.0);
    for (int32_t i = r0; i < r1; ++i)
        for (int32_t p = T.rp[i]; p < T.rp[i + 1]; ++p) {
            const int32_t c = T.ci[p];
            if (c >= r0 && c < r1) D[(size_t)(i - r0) * w + (c - r0)] = T.va[p];
        }
    tri_inverse(D.data(), X.data(), w, T.upper, T.unit);   // host_dense.cpp
}

static double g_inv_ms = 0.0;

struct RowRef {     // a row of the program before it is laid out
    int32_t block;  // supernode index
    int32_t k;      // row within the supernode
    int32_t len;    // number of entries
    uint8_t kind;   // 0: A row, 1: B row, 2: one-step (merged) row, 3: copy of a merged result
};

// What a fresh build records about one triangular factor so that a later matrix with the SAME
// index arrays only has to recompute numbers (refill_factor): the rows of every sub-level in
// their final order and where their entries live in the sliced-ELLPACK arrays.
struct FactorTemplate {
    std::vector<int32_t> level_ptr;   // rows of recorded level j: [level_ptr[j], level_ptr[j+1])
    std::vector<RowRef> rows;
    std::vector<int32_t> q;           // program row
    std::vector<int32_t> base;        // position of the row's entry 0
    std::vector<uint8_t> g;           // log2(lanes per row)
    // rows of kind 1 / 2 grouped by supernode (index_blocks): the refill computes one block
    // inverse at a time into scratch buffers and writes its rows straight away
    std::vector<int32_t> bl_level_ptr;   // block entries of recorded level j
    std::vector<int32_t> bl_block, bl_ptr, bl_x;
    // merged (one-step) blocks: where every stored entry of the block's rows goes - index into the
    // dense off-block matrix To (>= 0) or into the diagonal block D (-1 - index) -, the width of To
    // and the column count of every row: all of it depends on the index arrays only
    std::vector<int64_t> mb_ptr;         // per block entry (kind 2 only; equal neighbours otherwise)
    std::vector<int32_t> mb_dst;
    std::vector<int32_t> mb_cu;
    std::vector<int64_t> mb_ncol_ptr;
    std::vector<int32_t> mb_ncol;
    void index_blocks(int nb, const Tri& T, const std::vector<int32_t>& starts);
};

void FactorTemplate::index_blocks(int nb, const Tri& T, const std::vector<int32_t>& starts) {
    bl_level_ptr.assign(1, 0);
    bl_ptr.assign(1, 0);
    mb_ptr.assign(1, 0);
    mb_ncol_ptr.assign(1, 0);
    std::vector<int32_t> pos;            // column -> index in the block's column list
    std::vector<int32_t> slot(nb, -1);
    std::vector<std::vector<int32_t>> bucket;
    std::vector<int32_t> order;
    for (size_t lv = 0; lv + 1 < level_ptr.size(); ++lv) {
        bucket.clear();
        order.clear();
        for (int32_t x = level_ptr[lv]; x < level_ptr[lv + 1]; ++x) {
            const RowRef& rr = rows[x];
            if (rr.kind != 1 && rr.kind != 2) continue;
            if (slot[rr.block] < 0) {
                slot[rr.block] = (int32_t)bucket.size();
                bucket.emplace_back();
                order.push_back(rr.block);
            }
            bucket[slot[rr.block]].push_back(x);
        }
        for (size_t j = 0; j < order.size(); ++j) {
            bl_block.push_back(order[j]);
            bl_x.insert(bl_x.end(), bucket[j].begin(), bucket[j].end());
            bl_ptr.push_back((int32_t)bl_x.size());
            slot[order[j]] = -1;
            int32_t cu = 0;
            if (rows[bucket[j][0]].kind == 2) {
                // the same walk as multiply_block: columns in order of first appearance over the
                // rows in dependency order, then the destination of every stored entry
                const int32_t r0 = starts[order[j]], r1 = starts[order[j] + 1], w = r1 - r0;
                if (pos.empty()) pos.assign(T.rp ? (size_t)starts.back() : 0, -1);
                std::vector<int32_t> cols, ncol(w, 0);
                for (int kk = 0; kk < w; ++kk) {
                    const int k = T.upper ? w - 1 - kk : kk;
                    for (int32_t p = T.rp[r0 + k]; p < T.rp[r0 + k + 1]; ++p) {
                        const int32_t c = T.ci[p];
                        if ((c < r0 || c >= r1) && pos[c] < 0) {
                            pos[c] = (int32_t)cols.size();
                            cols.push_back(c);
                        }
                    }
                    ncol[k] = (int32_t)cols.size();
                }
                cu = (int32_t)cols.size();
                for (int k = 0; k < w; ++k)
                    for (int32_t p = T.rp[r0 + k]; p < T.rp[r0 + k + 1]; ++p) {
                        const int32_t c = T.ci[p];
                        mb_dst.push_back((c < r0 || c >= r1) ? k * cu + pos[c] : -1 - (k * w + (c - r0)));
                    }
                for (int32_t c : cols) pos[c] = -1;
                mb_ncol.insert(mb_ncol.end(), ncol.begin(), ncol.end());
            }
            mb_cu.push_back(cu);
            mb_ptr.push_back((int64_t)mb_dst.size());
            mb_ncol_ptr.push_back((int64_t)mb_ncol.size());
        }
        bl_level_ptr.push_back((int32_t)bl_block.size());
    }
}

struct MergedBlock {            // inverse-multiplied form of one merged block
    std::vector<int32_t> cols;  // off-block columns in order of first appearance
    std::vector<int32_t> ncol;  // per row k: how many of them row k uses
    std::vector<double> X;      // w x w inverse of the diagonal block
    std::vector<double> Pm;     // w x cols.size(): X * T[t, cols]
};

// inverse-multiplied form of the merged block t: X = inv(T_tt), Pm = X * T[t, off-block]
void multiply_block(int64_t n, const Tri& T, int32_t r0, int32_t r1, std::vector<int32_t>* posbuf,
                    std::vector<double>* Dbuf, std::vector<double>* Tbuf, MergedBlock* mb) {
    const int w = r1 - r0;
    std::vector<int32_t>& pos = *posbuf;     // column -> index in mb->cols (-1 outside this call)
    if ((int64_t)pos.size() < n) pos.assign(n, -1);
    invert_block(T, r0, r1, Dbuf, &mb->X);
    mb->cols.clear();
    mb->ncol.assign(w, 0);
    for (int kk = 0; kk < w; ++kk) {         // rows in dependency order
        const int k = T.upper ? w - 1 - kk : kk;
        for (int32_t p = T.rp[r0 + k]; p < T.rp[r0 + k + 1]; ++p) {
            const int32_t c = T.ci[p];
            if ((c < r0 || c >= r1) && pos[c] < 0) {
                pos[c] = (int32_t)mb->cols.size();
                mb->cols.push_back(c);
            }
        }
        mb->ncol[k] = (int32_t)mb->cols.size();
    }
    const int cu = (int)mb->cols.size();
    std::vector<double>& To = *Tbuf;
    To.assign((size_t)w * cu, 0.0);
    for (int k = 0; k < w; ++k)
        for (int32_t p = T.rp[r0 + k]; p < T.rp[r0 + k + 1]; ++p) {
            const int32_t c = T.ci[p];
            if (c < r0 || c >= r1) To[(size_t)k * cu + pos[c]] = T.va[p];
        }
    mb->Pm.assign((size_t)w * cu, 0.0);
    tri_times_dense(mb->X.data(), To.data(), mb->Pm.data(), w, cu, T.upper);   // host_dense.cpp
    for (int32_t c : mb->cols) pos[c] = -1;
}

int emit_factor(int64_t n, const Tri& T, const std::vector<int32_t>& starts,
                const std::vector<BlockPlan>& plan, int32_t nsub, int64_t ymax, int max_lanes,
                LuProgram* P, FactorTemplate* rec) {
    const int nb = (int)starts.size() - 1;
    if (rec) rec->level_ptr.assign(1, 0);
    std::vector<std::vector<RowRef>> lev(nsub);
    // rows of merged blocks: the sub-level that produces them and where a reader one sub-level
    // later finds them (the y region; everybody later reads the home row)
    std::vector<int32_t> mlev, maddr;
    bool any_merged = false;
    for (int t = 0; t < nb; ++t) any_merged = any_merged || plan[t].merged;
    if (any_merged) {
        mlev.assign(n, -1);
        maddr.assign(n, -1);
    }
    auto ybase_of = [&](const BlockPlan& bp) {
        return (int32_t)(n + (bp.sA & 1) * ymax + (bp.yoff < 0 ? 0 : bp.yoff));
    };
    // off-block entry counts per row
    for (int t = 0; t < nb; ++t) {
        const BlockPlan& bp = plan[t];
        if (bp.sA < 0) continue;
        const int32_t r0 = starts[t], r1 = starts[t + 1], w = r1 - r0;
        if (bp.merged) {
            for (int k = 0; k < w; ++k) {
                mlev[r0 + k] = bp.sA;
                maddr[r0 + k] = ybase_of(bp) + k;
                lev[bp.sA].push_back(RowRef{t, k, 0, 2});      // length: set when the block is formed
                lev[bp.sA + 1].push_back(RowRef{t, k, 0, 3});
            }
            continue;
        }
        for (int32_t i = r0; i < r1; ++i) {
            int32_t len = 0;
            for (int32_t p = T.rp[i]; p < T.rp[i + 1]; ++p) len += (T.ci[p] < r0 || T.ci[p] >= r1);
            lev[bp.sA].push_back(RowRef{t, i - r0, len, 0});
        }
        if (w > 1)
            for (int k = 0; k < w; ++k) lev[bp.sA + 1].push_back(RowRef{t, k, T.upper ? w - k : k + 1, 1});
    }
    std::vector<double> D, X, Tbuf;
    std::vector<int32_t> posbuf;
    std::vector<int32_t> inv_slot(nb, -1);
    std::vector<std::vector<double>> inv;   // inverses of the diagonal blocks of this sub-level
    std::vector<MergedBlock> mblocks;       // merged blocks of this sub-level
    for (int32_t s = 0; s < nsub; ++s) {
        std::vector<RowRef>& rows = lev[s];
        if (rows.empty()) continue;
        inv.clear();
        mblocks.clear();
        const auto ti0 = std::chrono::steady_clock::now();
        for (RowRef& rr : rows) {
            if (rr.kind == 1 && rr.k == 0) {
                invert_block(T, starts[rr.block], starts[rr.block + 1], &D, &X);
                inv_slot[rr.block] = (int32_t)inv.size();
                inv.push_back(X);
            } else if (rr.kind == 2) {
                if (rr.k == 0) {
                    inv_slot[rr.block] = (int32_t)mblocks.size();
                    mblocks.emplace_back();
                    multiply_block(n, T, starts[rr.block], starts[rr.block + 1], &posbuf, &D, &Tbuf,
                                   &mblocks.back());
                }
                const MergedBlock& mb = mblocks[inv_slot[rr.block]];
                const int w = starts[rr.block + 1] - starts[rr.block];
                rr.len = (T.upper ? w - rr.k : rr.k + 1) + mb.ncol[rr.k];
            }
        }
        g_inv_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ti0).count();
        std::stable_sort(rows.begin(), rows.end(), [](const RowRef& a, const RowRef& b) {
            if (a.len != b.len) return a.len > b.len;
            return a.block < b.block;
        });
        // lanes per row by length class (about lane_entries entries per lane), boosted while
        // the whole sub-level still fits one pass of the CTA
        std::vector<int> gl(rows.size());
        int64_t slots = 0;
        static int lane_entries = 0;
        if (lane_entries == 0) {
            const char* env = getenv("OCB_LANE_ENTRIES");
            lane_entries = env ? atoi(env) : 16;
            if (lane_entries < 1) lane_entries = 16;
        }
        for (size_t r = 0; r < rows.size(); ++r) {
            gl[r] = std::min(5, pow2ceil((rows[r].len + lane_entries - 1) / lane_entries));
            slots += 1LL << gl[r];
        }
        while (slots * 2 <= max_lanes) {
            bool any = false;
            slots = 0;
            for (size_t r = 0; r < rows.size(); ++r) {
                if (gl[r] < 5 && (2 << gl[r]) <= std::max(rows[r].len, 1)) { ++gl[r]; any = true; }
                slots += 1LL << gl[r];
            }
            if (!any) break;
        }
        for (size_t r = 1; r < rows.size(); ++r) gl[r] = std::min(gl[r], gl[r - 1]);  // monotone
        // where a row of sub-level s reads the value of column c
        auto addr = [&](int32_t c) {
            return (any_merged && mlev[c] >= 0 && mlev[c] + 1 == s) ? maddr[c] : c;
        };
        size_t r = 0;
        while (r < rows.size()) {
            const int g = gl[r], G = 1 << g, rmax = 32 >> g;
            size_t r2 = r;
            while (r2 < rows.size() && gl[r2] == g && (int)(r2 - r) < rmax) ++r2;
            const int nr = (int)(r2 - r);
            Slice sl;
            sl.q0 = (int32_t)P->dst.size();
            sl.glog_nrows = g | (nr << 8);
            sl.ebase = (int32_t)P->col.size();
            sl.trips = (rows[r].len + G - 1) / G;   // rows are sorted: the first is the longest
            const size_t e0 = P->col.size();
            P->col.resize(e0 + (size_t)sl.trips * 32, 0);
            P->val.resize(e0 + (size_t)sl.trips * 32, 0.0);
            for (int rl = 0; rl < nr; ++rl) {
                const RowRef& rr = rows[r + rl];
                const BlockPlan& bp = plan[rr.block];
                const int32_t r0 = starts[rr.block], r1 = starts[rr.block + 1], w = r1 - r0;
                const int32_t ybase = ybase_of(bp);
                const int32_t i = r0 + rr.k;
                if (rec) {
                    rec->rows.push_back(rr);
                    rec->q.push_back((int32_t)P->dst.size());
                    rec->base.push_back((int32_t)(e0 + ((size_t)rl << g)));
                    rec->g.push_back((uint8_t)g);
                }
                int e = 0;   // entry counter of this row
                auto put = [&](int32_t c, double v) {
                    const size_t pos = e0 + ((size_t)(e >> g) << 5) + ((size_t)rl << g) + (size_t)(e & (G - 1));
                    P->col[pos] = c;
                    P->val[pos] = v;
                    ++e;
                };
                if (rr.kind == 0) {
                    double dg = 1.0;
                    for (int32_t p = T.rp[i]; p < T.rp[i + 1]; ++p) {
                        const int32_t c = T.ci[p];
                        if (c < r0 || c >= r1) {
                            put(addr(c), T.va[p]);
                        } else if (c == i && !T.unit) {
                            dg = T.va[p];
                        }
                    }
                    P->init.push_back(i);
                    if (w == 1) {
                        P->dst.push_back(i);
                        P->scale.push_back(1.0 / dg);
                    } else {
                        P->dst.push_back(ybase + rr.k);
                        P->scale.push_back(1.0);
                    }
                } else if (rr.kind == 1) {
                    const std::vector<double>& Xi = inv[inv_slot[rr.block]];
                    const int j0 = T.upper ? rr.k : 0, j1 = T.upper ? w : rr.k + 1;
                    for (int j = j0; j < j1; ++j) put(ybase + j, -Xi[(size_t)rr.k * w + j]);
                    P->init.push_back(-1);
                    P->dst.push_back(i);
                    P->scale.push_back(1.0);
                } else if (rr.kind == 2) {
                    // x_i = sum_j X[k,j] b_j - sum_c Pm[k,c] x_c: the right-hand sides b_j are
                    // still in the home rows of this block (nobody writes them in this sub-level)
                    const MergedBlock& mb = mblocks[inv_slot[rr.block]];
                    const int j0 = T.upper ? rr.k : 0, j1 = T.upper ? w : rr.k + 1;
                    for (int j = j0; j < j1; ++j) put(r0 + j, -mb.X[(size_t)rr.k * w + j]);
                    const size_t cu = mb.cols.size();
                    for (int32_t j = 0; j < mb.ncol[rr.k]; ++j) put(addr(mb.cols[j]), mb.Pm[(size_t)rr.k * cu + j]);
                    P->init.push_back(-1);
                    P->dst.push_back(ybase + rr.k);
                    P->scale.push_back(1.0);
                } else {
                    P->init.push_back(ybase + rr.k);     // copy the merged result home
                    P->dst.push_back(i);
                    P->scale.push_back(1.0);
                }
                P->nent_actual += e;
            }
            P->slices.push_back(sl);
            r = r2;
        }
        P->sub_ptr.push_back((int32_t)P->slices.size());
        if (rec) rec->level_ptr.push_back((int32_t)rec->rows.size());
    }
    return OCB_OK;
}

// Numbers only: the rows recorded by emit_factor, for a factor with the same index arrays.
// Mirrors the value computations of emit_factor exactly (tests compare the images byte by byte).
int refill_factor(int64_t n, const Tri& T, const std::vector<int32_t>& starts,
                  const std::vector<BlockPlan>& plan, const FactorTemplate& rec, LuProgram* P) {
    (void)plan;
    std::vector<double> D, X, Tbuf;
    std::vector<int32_t> posbuf;
    MergedBlock mb;                       // scratch, reused by every merged block
    for (size_t lv = 0; lv + 1 < rec.level_ptr.size(); ++lv) {
        const int32_t a = rec.level_ptr[lv], b = rec.level_ptr[lv + 1];
        for (int32_t x = a; x < b; ++x) {        // plain rows: entries outside the diagonal block
            const RowRef& rr = rec.rows[x];
            if (rr.kind != 0) continue;
            const int32_t r0 = starts[rr.block], r1 = starts[rr.block + 1], w = r1 - r0;
            const int32_t i = r0 + rr.k;
            const int g = rec.g[x], G = 1 << g;
            double* vbase = P->val.data() + rec.base[x];
            int e = 0;
            double dg = 1.0;
            for (int32_t p = T.rp[i]; p < T.rp[i + 1]; ++p) {
                const int32_t c = T.ci[p];
                if (c < r0 || c >= r1) {
                    vbase[((size_t)(e >> g) << 5) + (size_t)(e & (G - 1))] = T.va[p];
                    ++e;
                } else if (c == i && !T.unit) {
                    dg = T.va[p];
                }
            }
            if (!T.unit && dg == 0.0) {
                set_error("the %s factor has a zero pivot in row %d", T.upper ? "upper" : "lower", i);
                return OCB_ERR_SINGULAR;
            }
            if (w == 1) P->scale[rec.q[x]] = 1.0 / dg;
            if (e != rr.len) {
                set_error("program template does not match the factor (row %d)", i);
                return OCB_ERR_ARG;
            }
        }
        const auto ti0 = std::chrono::steady_clock::now();
        for (int32_t bi = rec.bl_level_ptr[lv]; bi < rec.bl_level_ptr[lv + 1]; ++bi) {
            const int32_t t = rec.bl_block[bi];
            const int32_t r0 = starts[t], r1 = starts[t + 1], w = r1 - r0;
            const int32_t xa = rec.bl_ptr[bi], xb = rec.bl_ptr[bi + 1];
            const int kind = rec.rows[rec.bl_x[xa]].kind;
            const double* Xm;
            if (kind == 1) {
                invert_block(T, r0, r1, &D, &X);
                Xm = X.data();
            } else {
                // multiply_block with the recorded destinations: one pass over the stored entries
                const int32_t cu_ = rec.mb_cu[bi];
                const int32_t* dstp = rec.mb_dst.data() + rec.mb_ptr[bi];
                D.assign((size_t)w * w, 0.0);
                Tbuf.assign((size_t)w * cu_, 0.0);
                for (int32_t i = r0; i < r1; ++i)
                    for (int32_t p = T.rp[i]; p < T.rp[i + 1]; ++p) {
                        const int32_t d = *dstp++;
                        if (d >= 0) Tbuf[d] = T.va[p]; else D[-1 - d] = T.va[p];
                    }
                for (int j = 0; j < w && !T.unit; ++j)
                    if (D[(size_t)j * w + j] == 0.0) {
                        set_error("the %s factor has a zero pivot in row %d", T.upper ? "upper" : "lower", r0 + j);
                        return OCB_ERR_SINGULAR;
                    }
                mb.X.assign((size_t)w * w, 0.0);
                tri_inverse(D.data(), mb.X.data(), w, T.upper, T.unit);
                mb.Pm.assign((size_t)w * cu_, 0.0);
                tri_times_dense(mb.X.data(), Tbuf.data(), mb.Pm.data(), w, cu_, T.upper);
                mb.ncol.assign(rec.mb_ncol.begin() + rec.mb_ncol_ptr[bi], rec.mb_ncol.begin() + rec.mb_ncol_ptr[bi + 1]);
                Xm = mb.X.data();
            }
            const size_t cu = kind == 2 ? (size_t)rec.mb_cu[bi] : 0;
            for (int32_t xi = xa; xi < xb; ++xi) {
                const int32_t x = rec.bl_x[xi];
                const RowRef& rr = rec.rows[x];
                const int g = rec.g[x], G = 1 << g;
                double* vbase = P->val.data() + rec.base[x];
                int e = 0;
                const int j0 = T.upper ? rr.k : 0, j1 = T.upper ? w : rr.k + 1;
                const double* xr = Xm + (size_t)rr.k * w;
                for (int j = j0; j < j1; ++j, ++e) vbase[((size_t)(e >> g) << 5) + (size_t)(e & (G - 1))] = -xr[j];
                if (kind == 2) {
                    const double* pr = mb.Pm.data() + (size_t)rr.k * cu;
                    const int32_t nc = mb.ncol[rr.k];
                    for (int32_t j = 0; j < nc; ++j, ++e) vbase[((size_t)(e >> g) << 5) + (size_t)(e & (G - 1))] = pr[j];
                }
                if (e != rr.len) {
                    set_error("program template does not match the factor (row %d)", r0 + rr.k);
                    return OCB_ERR_ARG;
                }
            }
        }
        g_inv_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ti0).count();
    }
    return OCB_OK;
}

// Structure of a program by the index arrays it was built from.  The factors of one run share
// a handful of structures (same ordering, same pivots: every shift and every time step of the
// cavity problem yields the SAME index arrays), so most builds only recompute numbers.
struct ProgramTemplate {
    int64_t n = 0;
    int max_lanes = 0, wcap = 0;
    bool transposed = false, merge = false;
    MergeRule rule;
    std::vector<int32_t> Lrp, Lci, Urp, Uci;        // the key (raw, as handed in)
    std::vector<int32_t> Uci_sorted, Uperm;         // sorted rows of U and where each entry came from
    std::vector<int32_t> starts;
    std::vector<BlockPlan> planL, planU;
    FactorTemplate recL, recU;
    LuProgram P;                                    // everything but the numbers (val is empty)
    int64_t nent = 0;
    uint64_t stamp = 0;
    size_t bytes() const {
        return sizeof(int32_t) * (Lrp.size() + Lci.size() + Urp.size() + Uci.size() + Uci_sorted.size() +
                                  Uperm.size() + P.col.size() + 3 * P.dst.size() + 4 * recL.rows.size() +
                                  4 * recU.rows.size()) + sizeof(double) * P.scale.size();
    }
};
constexpr size_t TEMPLATE_BYTES_MAX = 256u << 20;   // all cached structures of one process

std::mutex g_tmpl_mutex;
std::vector<std::unique_ptr<ProgramTemplate>> g_templates;
uint64_t g_tmpl_clock = 0;

bool same_ints(const std::vector<int32_t>& a, const int32_t* b, size_t count) {
    return a.size() == count && (count == 0 || memcmp(a.data(), b, count * sizeof(int32_t)) == 0);
}

MergeRule merge_rule(bool merge) {
    MergeRule rule;
    if (merge) {
        const char* ew = getenv("OCB_MERGE_W");
        const char* eg = getenv("OCB_MERGE_GROWTH");
        rule.max_w = ew ? atoi(ew) : 64;
        rule.growth = eg ? atof(eg) : 1.3;
    }
    return rule;
}

}  // namespace

int64_t g_template_hits = 0;

int build_lu_program(int64_t n, const int32_t* Lrp, const int32_t* Lci, const double* Lva,
                     const int32_t* Urp, const int32_t* Uci, const double* Uva, int max_lanes,
                     bool transposed, bool merge, LuProgram* P, int wmax_cap) {
    const int wcap = wmax_cap > 0 ? std::min(wmax_cap, wmax()) : wmax();
    auto reset = [&] {
        *P = LuProgram();
        P->n = n;
        P->sub_ptr.push_back(0);
    };
    if (n == 0) {
        reset();
        return OCB_OK;
    }
    if ((int64_t)Lrp[n] + Urp[n] + 2 * n >= (int64_t)INT32_MAX / 2) {
        reset();
        set_error("factor too large for int32 program indices");
        return OCB_ERR_ARG;
    }
    const bool timing = getenv("OCB_TIMING") != nullptr;
    auto tnow = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    const MergeRule rule = merge_rule(merge);
    const bool use_templates = getenv("OCB_NO_TEMPLATE") == nullptr;
    if (use_templates) {
        // a structure seen before?  then only the numbers are recomputed
        const auto tt0 = tnow();
        std::unique_lock<std::mutex> lock(g_tmpl_mutex);
        ProgramTemplate* hit = nullptr;
        for (auto& t : g_templates)
            if (t->n == n && t->max_lanes == max_lanes && t->wcap == wcap && t->transposed == transposed &&
                t->merge == merge &&
                t->rule.max_w == rule.max_w && t->rule.growth == rule.growth &&
                same_ints(t->Lrp, Lrp, (size_t)n + 1) && same_ints(t->Urp, Urp, (size_t)n + 1) &&
                same_ints(t->Lci, Lci, (size_t)Lrp[n]) && same_ints(t->Uci, Uci, (size_t)Urp[n])) {
                hit = t.get();
                break;
            }
        if (hit) {
            const auto tt1 = tnow();
            hit->stamp = ++g_tmpl_clock;
            // structure; the numbers follow.  A caller that hands in the program of its previous
            // call (ocb_lu_pack_host keeps one per thread) and hits the same structure again keeps
            // it: every number is rewritten below and the padding is still zero - no 25 MB copy
            if (!(P->structure_id != 0 && P->structure_id == hit->P.structure_id && P->n == n &&
                  (int64_t)P->val.size() == hit->nent && P->nent() == hit->nent)) {
                *P = hit->P;
                P->val.assign((size_t)hit->nent, 0.0);
            }
            const auto tt2 = tnow();
            std::vector<double> Uva_sorted;
            const int32_t* Ucs = Uci;
            const double* Uvs = Uva;
            if (!hit->Uperm.empty()) {
                Uva_sorted.resize(hit->Uperm.size());
                for (size_t j = 0; j < hit->Uperm.size(); ++j) Uva_sorted[j] = Uva[hit->Uperm[j]];
                Ucs = hit->Uci_sorted.data();
                Uvs = Uva_sorted.data();
            }
            const Tri TL{Lrp, Lci, Lva, false, !transposed}, TU{Urp, Ucs, Uvs, true, transposed};
            int rc = refill_factor(n, TL, hit->starts, hit->planL, hit->recL, P);
            if (rc == OCB_OK) rc = refill_factor(n, TU, hit->starts, hit->planU, hit->recU, P);
            ++g_template_hits;
            if (timing) {
                fprintf(stderr, "lu_program: structure template hit, numbers refilled in %.1f ms (lookup %.1f, copy %.1f, block inverses %.1f ms)\n",
                        ms(tt0, tnow()), ms(tt0, tt1), ms(tt1, tt2), g_inv_ms);
                g_inv_ms = 0.0;
            }
            return rc;
        }
    }
    reset();
    std::unique_ptr<ProgramTemplate> tmpl;
    if (use_templates) {
        tmpl.reset(new ProgramTemplate());
        tmpl->n = n;
        tmpl->max_lanes = max_lanes;
        tmpl->wcap = wcap;
        tmpl->transposed = transposed;
        tmpl->merge = merge;
        tmpl->rule = rule;
        tmpl->Lrp.assign(Lrp, Lrp + n + 1);
        tmpl->Urp.assign(Urp, Urp + n + 1);
        tmpl->Lci.assign(Lci, Lci + Lrp[n]);
        tmpl->Uci.assign(Uci, Uci + Urp[n]);
    }
    // The supernode search compares sorted rows of U.  SuperLU hands its columns out in
    // supernodal order, not sorted, but consistently: inside one of ITS supernodes the index
    // list of a column is the previous column's list without the first entry, so the sorting
    // permutation carries over (a comparison sort runs only where that chain breaks; sorting
    // every row in the caller costs 11 ms for the N=25 cavity factor, this 1.5 ms).
    const auto tsort0 = std::chrono::steady_clock::now();
    std::vector<int32_t> Uci_sorted;
    std::vector<double> Uva_sorted;
    bool sorted = true;
    for (int64_t i = 0; i < n && sorted; ++i)
        for (int32_t p = Urp[i] + 1; p < Urp[i + 1]; ++p)
            if (Uci[p - 1] >= Uci[p]) { sorted = false; break; }
    if (!sorted) {
        Uci_sorted.resize(Urp[n]);
        Uva_sorted.resize(Urp[n]);
        if (tmpl) tmpl->Uperm.reserve(Urp[n]);
        std::vector<int32_t> perm;    // raw positions of the current row in sorted order
        for (int64_t i = 0; i < n; ++i) {
            const int32_t a0 = Urp[i], len = Urp[i + 1] - a0;
            const int32_t* raw = Uci + a0;
            bool chain = false;
            if (i > 0 && len > 0 && (int32_t)perm.size() == len + 1 && perm[0] == 0)
                chain = memcmp(Uci + Urp[i - 1] + 1, raw, (size_t)len * sizeof(int32_t)) == 0;
            if (chain) {
                for (int32_t j = 0; j < len; ++j) perm[j] = perm[j + 1] - 1;
                perm.resize(len);
            } else {
                perm.resize(len);
                std::iota(perm.begin(), perm.end(), 0);
                bool row_sorted = true;
                for (int32_t j = 1; j < len && row_sorted; ++j) row_sorted = raw[j - 1] < raw[j];
                if (!row_sorted)
                    std::sort(perm.begin(), perm.end(), [raw](int32_t x, int32_t y) { return raw[x] < raw[y]; });
            }
            for (int32_t j = 0; j < len; ++j) {
                Uci_sorted[a0 + j] = raw[perm[j]];
                Uva_sorted[a0 + j] = Uva[a0 + perm[j]];
            }
            if (tmpl)
                for (int32_t j = 0; j < len; ++j) tmpl->Uperm.push_back(a0 + perm[j]);
        }
        if (tmpl) tmpl->Uci_sorted = Uci_sorted;
        Uci = Uci_sorted.data();
        Uva = Uva_sorted.data();
    }
    for (int64_t i = 0; i < n; ++i) {   // U rows: distinct column indices, diagonal first
        for (int32_t p = Urp[i] + 1; p < Urp[i + 1]; ++p)
            if (Uci[p - 1] >= Uci[p]) {
                set_error("U has a duplicate column index (row %lld)", (long long)i);
                return OCB_ERR_ARG;
            }
        if (Urp[i + 1] <= Urp[i] || Uci[Urp[i]] > i) {
            set_error("U has a zero pivot in row %lld", (long long)i);
            return OCB_ERR_SINGULAR;
        }
        if (Uci[Urp[i]] < i) {
            set_error("U has an entry below the diagonal (row %lld col %d)", (long long)i, Uci[Urp[i]]);
            return OCB_ERR_ARG;
        }
    }
    const auto t0 = tnow();
    if (timing) fprintf(stderr, "lu_program: sort+check %.1f ms\n", ms(tsort0, t0));
    std::vector<int32_t> starts;
    find_supernodes(n, Urp, Uci, wcap, &starts);
    P->nsuper = (int32_t)starts.size() - 1;
    for (size_t t = 0; t + 1 < starts.size(); ++t) P->max_w = std::max(P->max_w, starts[t + 1] - starts[t]);
    // layout 0: unit lower L, upper U with the pivots (P A Q = L U);  transposed layout: the
    // lower factor carries the pivots and the upper one is unit (A = (L U)^T = U^T L^T)
    const Tri TL{Lrp, Lci, Lva, false, !transposed}, TU{Urp, Uci, Uva, true, transposed};
    std::vector<BlockPlan> planL, planU;
    int64_t ymax = 0;
    int rc = plan_factor(n, TL, starts, rule, &planL, &P->nsub_L, &ymax);
    if (rc == OCB_OK) rc = plan_factor(n, TU, starts, rule, &planU, &P->nsub_U, &ymax);
    if (rc != OCB_OK) return rc;
    const auto t1 = tnow();
    P->ymax = ymax;
    P->n_ext = n + 2 * ymax;
    const int64_t cap = ((int64_t)Lrp[n] + Urp[n]) * 3 / 2 + 64 * n;   // SELL padding included
    P->col.reserve(cap);
    P->val.reserve(cap);
    for (int64_t i = 0; i < n; ++i) {   // stored entries, unit diagonals not counted
        for (int32_t p = Lrp[i]; p < Lrp[i + 1]; ++p) P->nnzL += (Lci[p] != i) || transposed;
        for (int32_t p = Urp[i]; p < Urp[i + 1]; ++p) P->nnzU += (Uci[p] != i) || !transposed;
    }
    const auto t2 = tnow();
    rc = emit_factor(n, TL, starts, planL, P->nsub_L, ymax, max_lanes, P, tmpl ? &tmpl->recL : nullptr);
    const auto t3 = tnow();
    P->nsub_L = (int32_t)P->nsub();
    if (rc == OCB_OK)
        rc = emit_factor(n, TU, starts, planU, P->nsub_U, ymax, max_lanes, P, tmpl ? &tmpl->recU : nullptr);
    P->nsub_U = (int32_t)P->nsub() - P->nsub_L;
    if (timing) {
        fprintf(stderr, "lu_program: supernodes+plan %.1f ms, reserve %.1f ms, emit L %.1f ms, emit U %.1f ms (block inverses %.1f ms)\n",
                ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, tnow()), g_inv_ms);
        g_inv_ms = 0.0;
    }
    if (rc == OCB_OK && tmpl) {
        tmpl->recL.index_blocks((int)starts.size() - 1, TL, starts);
        tmpl->recU.index_blocks((int)starts.size() - 1, TU, starts);
        tmpl->starts = starts;
        tmpl->planL = planL;
        tmpl->planU = planU;
        std::unique_lock<std::mutex> lock(g_tmpl_mutex);
        P->structure_id = ++g_tmpl_clock;
        tmpl->P = *P;
        tmpl->nent = P->nent();
        std::vector<double>().swap(tmpl->P.val);       // padding is zero, the rest is refilled
        tmpl->stamp = ++g_tmpl_clock;
        size_t total = tmpl->bytes();
        for (auto& t : g_templates) total += t->bytes();
        // keep the four most recently used structures, and no more than 256 MB of them
        while (!g_templates.empty() && (g_templates.size() >= 4 || total > TEMPLATE_BYTES_MAX)) {
            size_t oldest = 0;
            for (size_t j = 1; j < g_templates.size(); ++j)
                if (g_templates[j]->stamp < g_templates[oldest]->stamp) oldest = j;
            total -= g_templates[oldest]->bytes();
            g_templates.erase(g_templates.begin() + oldest);
        }
        if (total <= TEMPLATE_BYTES_MAX) g_templates.push_back(std::move(tmpl));
    }
    return rc;
}

void build_panels(const LuProgram& P, double max_pad, PanelProgram* out) {
    PanelProgram& Q = *out;
    Q = PanelProgram();
    Q.n = P.n;
    Q.n_ext = P.n_ext;
    Q.sub_ptr.push_back(0);
    struct Row { int32_t q, dst, init; std::vector<std::pair<int32_t, double>> e; };
    std::vector<Row> rows;
    std::vector<int32_t> order, ucols;
    std::vector<Panel> lvl;
    std::vector<double> lscale;
    std::vector<std::vector<int32_t>> lcols;
    std::vector<std::vector<double>> lvals;
    for (int64_t sb = 0; sb < P.nsub(); ++sb) {
        rows.clear();
        for (int32_t s = P.sub_ptr[sb]; s < P.sub_ptr[sb + 1]; ++s) {
            const Slice& sl = P.slices[s];
            const int g = sl.glog_nrows & 255, nr = sl.glog_nrows >> 8, G = 1 << g;
            for (int r = 0; r < nr; ++r) {
                Row rw;
                rw.q = sl.q0 + r;
                rw.dst = P.dst[rw.q];
                rw.init = P.init[rw.q];
                for (int u = 0; u < sl.trips; ++u) {
                    const size_t base = (size_t)sl.ebase + ((size_t)u << 5) + ((size_t)r << g);
                    for (int l = 0; l < G; ++l)
                        if (P.val[base + l] != 0.0) rw.e.emplace_back(P.col[base + l], P.val[base + l]);
                }
                std::sort(rw.e.begin(), rw.e.end());
                Q.entries_actual += (int64_t)rw.e.size();
                rows.push_back(std::move(rw));
            }
        }
        order.resize(rows.size());
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return rows[a].dst < rows[b].dst; });
        lvl.clear(); lscale.clear(); lcols.clear(); lvals.clear();
        size_t i = 0;
        while (i < order.size()) {
            // grow a panel from row i: consecutive destinations and initial rows, bounded padding
            size_t j = i + 1;
            ucols.clear();
            for (auto& e : rows[order[i]].e) ucols.push_back(e.first);
            int64_t sum = (int64_t)rows[order[i]].e.size();
            while (j < order.size() && (int)(j - i) < PANEL_ROWS) {
                const Row& a = rows[order[j - 1]];
                const Row& b = rows[order[j]];
                if (b.dst != a.dst + 1) break;
                if (!((a.init < 0 && b.init < 0) || (a.init >= 0 && b.init == a.init + 1))) break;
                std::vector<int32_t> merged;
                merged.reserve(ucols.size() + b.e.size());
                size_t x = 0, y = 0;
                while (x < ucols.size() || y < b.e.size()) {
                    if (y >= b.e.size() || (x < ucols.size() && ucols[x] < b.e[y].first)) merged.push_back(ucols[x++]);
                    else if (x >= ucols.size() || b.e[y].first < ucols[x]) merged.push_back(b.e[y++].first);
                    else { merged.push_back(ucols[x]); ++x; ++y; }
                }
                const int64_t cnt = (int64_t)(j - i) + 1;
                const int64_t s2 = sum + (int64_t)b.e.size();
                if ((double)merged.size() * (double)cnt > max_pad * (double)s2 + 8.0) break;
                ucols.swap(merged);
                sum = s2;
                ++j;
            }
            Panel pn;
            pn.cbase = 0;
            pn.ncol = (int32_t)ucols.size();
            pn.dst0 = rows[order[i]].dst;
            pn.init0 = rows[order[i]].init;
            pn.nrows = (int32_t)(j - i);
            pn.pad[0] = pn.pad[1] = pn.pad[2] = 0;
            std::vector<double> pv((size_t)ucols.size() * PANEL_ROWS, 0.0);
            for (size_t r = i; r < j; ++r) {
                const Row& rw = rows[order[r]];
                size_t x = 0;
                for (auto& e : rw.e) {
                    while (ucols[x] < e.first) ++x;
                    pv[x * PANEL_ROWS + (r - i)] = e.second;
                }
                lscale.push_back(P.scale[rw.q]);
            }
            for (size_t r = j - i; r < (size_t)PANEL_ROWS; ++r) lscale.push_back(0.0);
            lvl.push_back(pn);
            lcols.push_back(ucols);
            lvals.push_back(std::move(pv));
            i = j;
        }
        // longest panels first (load balance: the executor deals panels to CTAs in order)
        std::vector<int32_t> po(lvl.size());
        std::iota(po.begin(), po.end(), 0);
        std::stable_sort(po.begin(), po.end(), [&](int32_t a, int32_t b) { return lvl[a].ncol > lvl[b].ncol; });
        for (int32_t pi : po) {
            Panel pn = lvl[pi];
            pn.cbase = (int32_t)Q.pcol.size();
            Q.pcol.insert(Q.pcol.end(), lcols[pi].begin(), lcols[pi].end());
            Q.pval.insert(Q.pval.end(), lvals[pi].begin(), lvals[pi].end());
            while (pn.ncol & 3) {   // the executor reads 4 entries per step: pad with 0 * xe[0]
                Q.pcol.push_back(0);
                Q.pval.insert(Q.pval.end(), PANEL_ROWS, 0.0);
                ++pn.ncol;
            }
            Q.scale.insert(Q.scale.end(), lscale.begin() + (size_t)pi * PANEL_ROWS,
                           lscale.begin() + (size_t)(pi + 1) * PANEL_ROWS);
            Q.panels.push_back(pn);
        }
        Q.sub_ptr.push_back((int32_t)Q.panels.size());
    }
}

void execute_panels_host(const PanelProgram& Q, const int32_t* perm_r, const int32_t* perm_c,
                         const double* b, double* x) {
    std::vector<double> xe((size_t)Q.n_ext + 1, 0.0), out;
    for (int64_t i = 0; i < Q.n; ++i) xe[perm_r[i]] = b[i];
    for (int64_t sb = 0; sb < Q.nsub(); ++sb) {
        out.clear();
        for (int32_t pi = Q.sub_ptr[sb]; pi < Q.sub_ptr[sb + 1]; ++pi) {
            const Panel& pn = Q.panels[pi];
            double acc[PANEL_ROWS] = {0};
            for (int32_t p = 0; p < pn.ncol; ++p) {
                const double xv = xe[Q.pcol[pn.cbase + p]];
                for (int r = 0; r < PANEL_ROWS; ++r) acc[r] = fma(Q.pval[((size_t)pn.cbase + p) * PANEL_ROWS + r], xv, acc[r]);
            }
            for (int r = 0; r < pn.nrows; ++r) {
                const double ini = pn.init0 >= 0 ? xe[pn.init0 + r] : 0.0;
                out.push_back((ini - acc[r]) * Q.scale[(size_t)pi * PANEL_ROWS + r]);
            }
        }
        size_t o = 0;
        for (int32_t pi = Q.sub_ptr[sb]; pi < Q.sub_ptr[sb + 1]; ++pi) {
            const Panel& pn = Q.panels[pi];
            for (int r = 0; r < pn.nrows; ++r) xe[pn.dst0 + r] = out[o++];
        }
    }
    for (int64_t j = 0; j < Q.n; ++j) x[j] = xe[perm_c[j]];
}

void execute_program_host(const LuProgram& P, const int32_t* perm_r, const int32_t* perm_c,
                          const double* b, double* x) {
    std::vector<double> xe((size_t)P.n_ext + 1, 0.0), out;
    for (int64_t i = 0; i < P.n; ++i) xe[perm_r[i]] = b[i];
    for (int64_t sb = 0; sb < P.nsub(); ++sb) {
        // rows of one sub-level are independent, but a row may overwrite its own init slot:
        // compute the whole sub-level, then store (the kernels do the same behind a barrier)
        out.clear();
        for (int32_t s = P.sub_ptr[sb]; s < P.sub_ptr[sb + 1]; ++s) {
            const Slice& sl = P.slices[s];
            const int g = sl.glog_nrows & 255, nr = sl.glog_nrows >> 8, G = 1 << g;
            // the 32 lanes of the slice side by side, as the warp does it (padding entries are
            // val = 0, col = 0): a fixed-width inner loop instead of one short loop per row
            double lane[32];
            for (int l = 0; l < 32; ++l) lane[l] = 0.0;
            const double* v = P.val.data() + sl.ebase;
            const int32_t* c = P.col.data() + sl.ebase;
            for (int u = 0; u < sl.trips; ++u, v += 32, c += 32)
                for (int l = 0; l < 32; ++l) lane[l] = fma(v[l], xe[c[l]], lane[l]);
            for (int r = 0; r < nr; ++r) {
                const int32_t q = sl.q0 + r;
                double acc = 0.0;
                for (int l = 0; l < G; ++l) acc += lane[(r << g) + l];
                const double ini = P.init[q] >= 0 ? xe[P.init[q]] : 0.0;
                out.push_back((ini - acc) * P.scale[q]);
            }
        }
        size_t o = 0;
        for (int32_t s = P.sub_ptr[sb]; s < P.sub_ptr[sb + 1]; ++s) {
            const Slice& sl = P.slices[s];
            const int nr = sl.glog_nrows >> 8;
            for (int r = 0; r < nr; ++r) xe[P.dst[sl.q0 + r]] = out[o++];
        }
    }
    for (int64_t j = 0; j < P.n; ++j) x[j] = xe[perm_c[j]];
}

}  // namespace ocb

// ---------------------------------------------------------------------------------
// host-only C ABI: lets the CPU test-suite execute the program without a GPU
// ---------------------------------------------------------------------------------
struct ocb_lu_program {
    ocb::LuProgram P;
};

extern "C" {

int ocb_lu_program_create(ocb_lu_program** out, int64_t n, const int32_t* h_L_rowptr,
                          const int32_t* h_L_colidx, const double* h_L_vals, const int32_t* h_U_rowptr,
                          const int32_t* h_U_colidx, const double* h_U_vals, int64_t flags) {
    OCB_ARG(out && n >= 0 && h_L_rowptr && h_U_rowptr, "lu_program_create");
    ocb_lu_program* h = new ocb_lu_program();
    const int rc = ocb::build_lu_program(n, h_L_rowptr, h_L_colidx, h_L_vals, h_U_rowptr, h_U_colidx,
                                         h_U_vals, 512, (flags & 2) != 0, (flags & 4) != 0, &h->P,
                                         (flags & 8) ? 32 : 0);
    if (rc != OCB_OK) {
        delete h;
        return rc;
    }
    *out = h;
    return OCB_OK;
}

int64_t ocb_lu_program_template_hits(void) { return ocb::g_template_hits; }

int ocb_lu_program_solve_host(const ocb_lu_program* prog, const int32_t* h_perm_r, const int32_t* h_perm_c,
                              const double* h_b, double* h_x, int64_t mode, double max_pad,
                              int64_t* h_stats4) {
    OCB_ARG(prog && h_perm_r && h_perm_c && h_b && h_x && (mode == 0 || mode == 1), "lu_program_solve_host");
    if (mode == 0) {
        ocb::execute_program_host(prog->P, h_perm_r, h_perm_c, h_b, h_x);
        return OCB_OK;
    }
    ocb::PanelProgram Q;
    ocb::build_panels(prog->P, max_pad > 1.0 ? max_pad : 1.6, &Q);
    ocb::execute_panels_host(Q, h_perm_r, h_perm_c, h_b, h_x);
    if (h_stats4) {
        h_stats4[0] = (int64_t)Q.panels.size();
        h_stats4[1] = (int64_t)Q.pcol.size() * ocb::PANEL_ROWS;   // stored values (zero padding included)
        h_stats4[2] = Q.entries_actual;
        h_stats4[3] = Q.nsub();
    }
    return OCB_OK;
}

int ocb_lu_program_destroy(ocb_lu_program* prog) {
    delete prog;
    return OCB_OK;
}

int ocb_lu_program_info(const ocb_lu_program* prog, int64_t* info12) {
    OCB_ARG(prog && info12, "lu_program_info");
    const ocb::LuProgram& P = prog->P;
    info12[0] = P.n;
    info12[1] = P.n_ext;
    info12[2] = P.ymax;
    info12[3] = P.nsub_L;
    info12[4] = P.nsub_U;
    info12[5] = P.nsuper;
    info12[6] = P.max_w;
    info12[7] = (int64_t)P.slices.size();
    info12[8] = P.nrows();
    info12[9] = P.nent();
    info12[10] = P.nnzL;
    info12[11] = P.nnzU;
    return OCB_OK;
}

int ocb_lu_program_export(const ocb_lu_program* prog, int32_t* h_sub_ptr, int32_t* h_slice4, int32_t* h_dst,
                          int32_t* h_init, double* h_scale, int32_t* h_col, double* h_val) {
    OCB_ARG(prog && h_sub_ptr && h_slice4 && h_dst && h_init && h_scale && h_col && h_val, "lu_program_export");
    const ocb::LuProgram& P = prog->P;
    memcpy(h_sub_ptr, P.sub_ptr.data(), P.sub_ptr.size() * sizeof(int32_t));
    memcpy(h_slice4, P.slices.data(), P.slices.size() * sizeof(ocb::Slice));
    memcpy(h_dst, P.dst.data(), P.dst.size() * sizeof(int32_t));
    memcpy(h_init, P.init.data(), P.init.size() * sizeof(int32_t));
    memcpy(h_scale, P.scale.data(), P.scale.size() * sizeof(double));
    memcpy(h_col, P.col.data(), P.col.size() * sizeof(int32_t));
    memcpy(h_val, P.val.data(), P.val.size() * sizeof(double));
    return OCB_OK;
}
}
