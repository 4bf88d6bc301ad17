// MRTM river routing, "skewed" warp kernel (xanthos/routing/mrtm.py:16-82), fp64, -fmad=false, results
// BIT-IDENTICAL to the reference's scipy-CSR formulation (same argument as mrtm.cu: every row of UM = UP - I is
// accumulated term by term in ascending column order).
//
// Why a second kernel.  In mrtm_warp_kernel (mrtm.cu) a lane holds one cell and all lanes of a warp work on the
// SAME sub-step t: the flow exchange, the row sum, the clamp decision (mrtm.py:54) and - when a flow changed - the
// repeated balance with F' (mrtm.py:56-69) form one dependent chain per sub-step, and the warps that hold a cell
// with dt V / L > 1 (it empties at every other sub-step) run that chain twice.  The run ends with the slowest
// such chain.
//
// Here the cells of a warp are SKEWED in time: water only moves downstream, so cell j at sub-step t needs nothing
// but the flows (F, F') of its tributaries at sub-step t.  A cell at depth D below its piece's outlet works `lag`
// = Dw - D sub-steps behind the warp's leaves: in loop iteration n it does sub-step n - lag, and every tributary
// (lag - 1) finished that sub-step one iteration earlier.  Consequences:
//   * both the trial flow F and the final flow F' of every tributary are known when a cell starts its sub-step,
//     so the balance for the decision (with F, mrtm.py:51) and the balance for the update (with F', mrtm.py:68)
//     are two independent sums: no vote, no repeated pass, no branch - every warp executes the same straight code
//     whether its cells clamp or not;
//   * a lane holds K cells (slots): K independent chains per warp (ILP), a quarter of the warps, half the
//     instructions per cell; slot 0 takes the cells with two or more tributaries (<= 4 before and <= 4 after the
//     cell's own column), the other slots the cells with at most one (their row is order-free);
//   * flows are exchanged through a double-buffered table in shared memory (one STS.128 per slot, one LDS.128 per
//     row term; the reader always reads the parity written one iteration earlier);
//   * cut edges between warps are time series in global memory (L2) indexed by the ABSOLUTE sub-step: the producer
//     exports its outlet's (F, F') one iteration after they are computed, the consumer stages them into shared
//     memory 32 sub-steps at a time with cp.async and feeds them to the table through "ghost" entries; progress
//     counters in sub-steps give hand-over and back-pressure.  No block- or grid-wide barrier in the time loop.
//   * month boundaries (new lateral inflow, monthly mean, ChStorage snapshot) are per-cell events: a cell crosses
//     the boundary `lag` iterations after the warp's leaves; the event code runs only in those iterations.
#include "mrtm_plan.cuh"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

namespace xan {

constexpr int SK_XG = 16;        // ghost entries (incoming cut edges) per warp
constexpr int SK_XO = 16;        // exported cells (outgoing cut edges) per warp
#ifndef XAN_SKEW_LAGM
#define XAN_SKEW_LAGM 1
#endif
// A cell works SK_LAGM iterations behind its tributaries.  1: the terms of iteration n + 1 are loaded after the stores of
// iteration n.  2: they were stored an iteration earlier and are loaded while iteration n is computed (no load / store
// latency between iterations, but twice the lags: longer month windows, three table buffers) - measured slower with
// 4 cells per lane (57.7 against 49.5 ms) and equal with 2 (44.5 ms), see DESIGN.md section 4.
constexpr int SK_LAGM = XAN_SKEW_LAGM;
constexpr int SK_DMAX = 62 / SK_LAGM - 1;   // largest piece height; lags reach SK_LAGM (SK_DMAX + 1) in warps with ghost entries
#ifndef XAN_SKEW_CH
#define XAN_SKEW_CH 64
#endif
constexpr int SK_CH = XAN_SKEW_CH;   // sub-steps of a cut-edge series staged per hand-over (power of two, >= 32)
constexpr int SK_W = 4 * SK_CH;  // staged entries kept per ghost (a chunk is read for at most 2 SK_CH iterations)
constexpr int SK_NB = 4;         // before / after terms of a wide row

struct SkewPlan {
    int K = 0, nw = 0, n_edges = 0, n_levels = 0, G = 0, O = 0, Dmax = 0, n_pieces = 0;
    std::vector<int> cell;         // [nw][K*32] cell index or -1
    std::vector<int> lag;          // [nw][K*32]
    std::vector<int> src;          // [nw][32][2*SK_NB + K - 1] table entry of every row term (ZERO entry = padding)
    std::vector<int> ghost_edge;   // [nw][SK_XG] edge or -1
    std::vector<int> ghost_lag;    // [nw][SK_XG]
    std::vector<int> exp_edge;     // [nw][SK_XO] edge or -1
    std::vector<int> exp_place;    // [nw][SK_XO] table entry of the exported cell
    std::vector<int> Dw;           // [nw] largest lag of the warp
    std::vector<int> edge_prod, edge_cons, edge_cell;
    // device copies
    int *d_cell = nullptr, *d_lag = nullptr, *d_src = nullptr, *d_ghost_edge = nullptr, *d_ghost_lag = nullptr,
        *d_exp_edge = nullptr, *d_exp_place = nullptr, *d_Dw = nullptr, *d_edge_prod = nullptr, *d_edge_cons = nullptr,
        *d_progress = nullptr;
    bool on_device = false;
    // step tables of the last calendar routed (device, kept with the plan: a per-call host -> device copy would queue
    // behind the forcing uploads of the ensemble runner on the copy engine and hold up the kernel)
    std::vector<int> cal_key;
    double cal_dt = 0.0;
    int *d_cal_int = nullptr;
    double *d_cal_secs = nullptr;
    int nsrc() const { return 2 * SK_NB + K - 1; }
    int zero_entry() const { return K * 32 + SK_XG; }
};

void skew_plan_destroy(SkewPlan *sp) {
    if (!sp) return;
    if (sp->on_device) {
        cudaFree(sp->d_cell); cudaFree(sp->d_lag); cudaFree(sp->d_src); cudaFree(sp->d_ghost_edge);
        cudaFree(sp->d_ghost_lag); cudaFree(sp->d_exp_edge); cudaFree(sp->d_exp_place); cudaFree(sp->d_Dw);
        cudaFree(sp->d_edge_prod); cudaFree(sp->d_edge_cons); cudaFree(sp->d_progress);
    }
    if (sp->d_cal_int) cudaFree(sp->d_cal_int);
    if (sp->d_cal_secs) cudaFree(sp->d_cal_secs);
    delete sp;
}

// =============================================================================================
// host: pieces, warps, places, lags
// =============================================================================================
static SkewPlan *build_skew(const xan_mrtm_plan *pl, int K) {
    const int n = pl->ncell;
    if (!pl->is_forest || n <= 0) return nullptr;
    const int CAP = 32 * K;
    auto nup = [&](int v) { return pl->upid[(size_t)v * 9 + 8]; };
    auto child = [&](int v, int s) { return pl->upid[(size_t)v * 9 + s] - 1; };
    // a wide row needs slot 0; its terms must fit SK_NB before and SK_NB after the own column
    for (int v = 0; v < n; ++v)
        if (nup(v) >= 2) {
            int before = 0;
            for (int s = 0; s < nup(v); ++s) before += child(v, s) < v ? 1 : 0;
            if (before > SK_NB || nup(v) - before > SK_NB) return nullptr;
        }
    std::vector<int> indeg(n), order;
    order.reserve(n);
    for (int i = 0; i < n; ++i) indeg[i] = nup(i);
    for (int i = 0; i < n; ++i)
        if (indeg[i] == 0) order.push_back(i);
    for (size_t h = 0; h < order.size(); ++h) {
        const int r = pl->down[order[h]];
        if (r >= 0 && --indeg[r] == 0) order.push_back(r);
    }
    if ((int)order.size() != n) return nullptr;

    // ---- pieces, bottom-up: size <= CAP, wide cells <= 32, incoming cut edges <= SK_XG, height <= SK_DMAX --------
    std::vector<int> sz(n, 0), wd(n, 0), gh(n, 0), ht(n, 0);
    std::vector<char> cut(n, 0);
    for (int h = 0; h < n; ++h) {
        const int v = order[h];
        const int k = nup(v);
        int ch[8];
        bool merged[8];
        for (int s = 0; s < k; ++s) {
            ch[s] = child(v, s);
            merged[s] = true;
        }
        for (;;) {
            int s_sz = 1, s_wd = (k >= 2) ? 1 : 0, s_gh = 0, s_ht = 0;
            for (int s = 0; s < k; ++s) {
                if (merged[s]) {
                    s_sz += sz[ch[s]];
                    s_wd += wd[ch[s]];
                    s_gh += gh[ch[s]];
                    s_ht = std::max(s_ht, ht[ch[s]] + 1);
                } else {
                    s_gh += 1;
                    s_ht = std::max(s_ht, 1);
                }
            }
            int viol = 0;   // 1 size, 2 wide, 3 ghosts, 4 height
            if (s_sz > CAP) viol = 1;
            else if (s_wd > 32) viol = 2;
            else if (s_gh > SK_XG) viol = 3;
            else if (s_ht > SK_DMAX) viol = 4;
            if (!viol) {
                sz[v] = s_sz; wd[v] = s_wd; gh[v] = s_gh; ht[v] = s_ht;
                break;
            }
            int best = -1, bestval = -1;
            for (int s = 0; s < k; ++s) {
                if (!merged[s]) continue;
                const int c = ch[s];
                const int val = viol == 1 ? sz[c] : viol == 2 ? wd[c] : viol == 3 ? gh[c] : ht[c];
                if (val > bestval || (val == bestval && best >= 0 && c < ch[best])) {
                    bestval = val;
                    best = s;
                }
            }
            if (best < 0) return nullptr;   // cannot happen: a bare cell satisfies every bound
            merged[best] = false;
            cut[ch[best]] = 1;
        }
    }
    std::vector<int> piece(n, -1), p_sz, p_wd, p_gh, p_ht, p_root;
    for (int h = n - 1; h >= 0; --h) {
        const int v = order[h];
        if (pl->down[v] < 0 || cut[v]) {
            piece[v] = (int)p_sz.size();
            p_sz.push_back(sz[v]); p_wd.push_back(wd[v]); p_gh.push_back(gh[v]); p_ht.push_back(ht[v]);
            p_root.push_back(v);
        } else {
            piece[v] = piece[pl->down[v]];
        }
    }
    const int np = (int)p_sz.size();
    std::vector<int> piece_down(np, -1), piece_level(np, 0);
    for (int h = 0; h < n; ++h) {   // leaves first
        const int v = order[h];
        if (cut[v]) {
            const int pu = piece[v], pd = piece[pl->down[v]];
            piece_down[pu] = pd;
            piece_level[pd] = std::max(piece_level[pd], piece_level[pu] + 1);
        }
    }
    int n_levels = 1;
    for (int p = 0; p < np; ++p) n_levels = std::max(n_levels, piece_level[p] + 1);

    // ---- warps: pieces that feed the same downstream piece may share a warp (one consumer per warp: no cycle through
    // the back-pressure); consumers without a receiver open a warp of their own; free pieces (whole trees) fill gaps ----
    struct Bin { int sz = 0, wd = 0, gh = 0, out = 0; };
    std::vector<Bin> bins;
    std::vector<int> piece_warp(np, -1), ids(np);
    std::iota(ids.begin(), ids.end(), 0);
    std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return p_sz[a] > p_sz[b]; });
    auto fits = [&](const Bin &b, int p, int out) {
        return b.sz + p_sz[p] <= CAP && b.wd + p_wd[p] <= 32 && b.gh + p_gh[p] <= SK_XG && b.out + out <= SK_XO;
    };
    auto put = [&](int w, int p, int out) {
        bins[w].sz += p_sz[p]; bins[w].wd += p_wd[p]; bins[w].gh += p_gh[p]; bins[w].out += out;
        piece_warp[p] = w;
    };
    {
        std::vector<std::vector<int>> open(np);   // per consumer piece: warps opened for its tributary pieces
        for (int p : ids) {
            const bool linked = piece_down[p] >= 0 || p_gh[p] > 0;
            if (!linked) continue;
            if (piece_down[p] < 0) {
                bins.emplace_back();
                put((int)bins.size() - 1, p, 0);
                continue;
            }
            int w = -1;
            for (int cand : open[piece_down[p]])
                if (fits(bins[cand], p, 1)) {
                    w = cand;
                    break;
                }
            if (w < 0) {
                bins.emplace_back();
                w = (int)bins.size() - 1;
                open[piece_down[p]].push_back(w);
            }
            put(w, p, 1);
        }
    }
    {
        std::vector<std::vector<int>> by_room(CAP + 1);
        for (size_t w = 0; w < bins.size(); ++w) by_room[CAP - bins[w].sz].push_back((int)w);
        for (int p : ids) {
            if (piece_down[p] >= 0 || p_gh[p] > 0) continue;
            int w = -1;
            for (int room = p_sz[p]; room <= CAP && w < 0; ++room) {
                auto &lst = by_room[room];
                for (size_t q = lst.size(); q-- > 0;)
                    if (fits(bins[lst[q]], p, 0)) {
                        w = lst[q];
                        lst.erase(lst.begin() + q);
                        break;
                    }
            }
            if (w < 0) {
                bins.emplace_back();
                w = (int)bins.size() - 1;
            }
            put(w, p, 0);
            by_room[CAP - bins[w].sz].push_back(w);
        }
    }
    const int nw = (int)bins.size();

    auto *sp = new SkewPlan();
    sp->K = K;
    sp->nw = nw;
    sp->n_levels = n_levels;
    sp->n_pieces = np;
    const int NS = sp->nsrc(), ZERO = sp->zero_entry();
    sp->cell.assign((size_t)nw * CAP, -1);
    sp->lag.assign((size_t)nw * CAP, 0);
    sp->src.assign((size_t)nw * 32 * NS, ZERO);
    sp->ghost_edge.assign((size_t)nw * SK_XG, -1);
    sp->ghost_lag.assign((size_t)nw * SK_XG, 0);
    sp->exp_edge.assign((size_t)nw * SK_XO, -1);
    sp->exp_place.assign((size_t)nw * SK_XO, ZERO);
    sp->Dw.assign(nw, 0);

    // ---- places: wide cells in slot 0, the others in slots 1 .. K-1 in depth-first post-order (a cell's tributary then
    // sits in the previous lane: neighbouring lanes read neighbouring table entries), overflow into slot 0 --------------
    std::vector<int> post;
    post.reserve(n);
    {
        std::vector<std::pair<int, int>> stack;
        for (int r = 0; r < n; ++r) {
            if (pl->down[r] >= 0) continue;
            stack.emplace_back(r, 0);
            while (!stack.empty()) {
                auto &top = stack.back();
                const int v = top.first;
                if (top.second < nup(v)) {
                    const int c = child(v, top.second);
                    ++top.second;
                    stack.emplace_back(c, 0);
                } else {
                    post.push_back(v);
                    stack.pop_back();
                }
            }
        }
    }
    std::vector<int> cell_warp(n), place(n, -1), next_wide(nw, 0), next_c(nw, 32);
    for (int v = 0; v < n; ++v) cell_warp[v] = piece_warp[piece[v]];
    for (int v : post)
        if (nup(v) >= 2) place[v] = next_wide[cell_warp[v]]++;
    for (int v : post)
        if (nup(v) < 2) {
            const int w = cell_warp[v];
            if (next_c[w] < CAP) place[v] = next_c[w]++;
            else place[v] = next_wide[w]++;
        }
    for (int w = 0; w < nw; ++w)
        if (next_wide[w] > 32) {
            delete sp;
            return nullptr;
        }
    // ---- depths (0 = outlet of a piece), cut edges, ghosts, lags --------------------------------------------------------
    std::vector<int> depth(n, 0), ghost_slot(n, -1), n_ghost(nw, 0), n_exp(nw, 0);
    for (int h = n - 1; h >= 0; --h) {
        const int v = order[h];
        depth[v] = (pl->down[v] < 0 || cut[v]) ? 0 : depth[pl->down[v]] + 1;
        sp->Dw[cell_warp[v]] = std::max(sp->Dw[cell_warp[v]], depth[v]);
    }
    for (int v = 0; v < n; ++v)
        if (cut[v]) {
            const int r = pl->down[v], wp = cell_warp[v], wc = cell_warp[r];
            if (wp == wc || n_ghost[wc] >= SK_XG || n_exp[wp] >= SK_XO) {
                delete sp;
                return nullptr;
            }
            const int e = (int)sp->edge_prod.size();
            sp->edge_prod.push_back(wp);
            sp->edge_cons.push_back(wc);
            sp->edge_cell.push_back(v);
            ghost_slot[v] = n_ghost[wc]++;
            sp->ghost_edge[(size_t)wc * SK_XG + ghost_slot[v]] = e;
            sp->Dw[wc] = std::max(sp->Dw[wc], depth[r] + 1);
            const int o = n_exp[wp]++;
            sp->exp_edge[(size_t)wp * SK_XO + o] = e;
            sp->exp_place[(size_t)wp * SK_XO + o] = place[v];
        }
    sp->n_edges = (int)sp->edge_prod.size();
    for (int w = 0; w < nw; ++w) {
        // ghost entries are loaded one iteration ahead of their use (software pipelining of the import): with every
        // lag of the warp one larger a ghost never has lag 0, i.e. never reads a series entry of a chunk still in flight
        if (n_ghost[w] > 0) sp->Dw[w] += 1;
        sp->G = std::max(sp->G, n_ghost[w]);
        sp->O = std::max(sp->O, n_exp[w]);
        sp->Dmax = std::max(sp->Dmax, SK_LAGM * sp->Dw[w]);
        if (sp->Dw[w] > SK_DMAX + 1) {
            delete sp;
            return nullptr;
        }
        sp->Dw[w] *= SK_LAGM;   // from here on: the largest lag of the warp
    }
    for (int v = 0; v < n; ++v) {
        const int w = cell_warp[v];
        sp->cell[(size_t)w * CAP + place[v]] = v;
        sp->lag[(size_t)w * CAP + place[v]] = sp->Dw[w] - SK_LAGM * depth[v];
        if (cut[v]) {
            const int r = pl->down[v], wc = cell_warp[r];
            sp->ghost_lag[(size_t)wc * SK_XG + ghost_slot[v]] = sp->Dw[wc] - SK_LAGM * (depth[r] + 1);
        }
    }
    // ---- row terms ----------------------------------------------------------------------------------------------------------
    for (int v = 0; v < n; ++v) {
        const int w = cell_warp[v], lane = place[v] & 31, slot = place[v] >> 5;
        int *row = &sp->src[((size_t)w * 32 + lane) * NS];
        auto entry = [&](int j) { return cell_warp[j] == w ? place[j] : CAP + ghost_slot[j]; };
        const int beg = pl->row_ptr[v], cnt = pl->row_ptr[v + 1] - beg;
        if (slot == 0) {
            int before = 0;
            for (int s = 0; s < cnt; ++s) before += pl->col[beg + s] < v ? 1 : 0;
            int ib = SK_NB - before, ia = SK_NB;   // before terms right-aligned in 0..3, after terms in 4..7
            for (int s = 0; s < cnt; ++s) {
                const int j = pl->col[beg + s];
                if (j == v) continue;
                if (j < v) row[ib++] = entry(j);
                else row[ia++] = entry(j);
            }
        } else {
            for (int s = 0; s < cnt; ++s) {
                const int j = pl->col[beg + s];
                if (j != v) row[2 * SK_NB + slot - 1] = entry(j);
            }
        }
    }
    return sp;
}

// Two plans per flow graph, built on first use: one for launches of a single member (K = 2 cells per lane measured
// fastest: two warps per SM sub-partition) and one for launches of several members (K = 4: fewer instructions and
// shared-memory wavefronts per cell; the other members' warps cover the latencies).  XANTHOS_MRTM_SKEW_K / _KM override.
static SkewPlan *get_skew(xan_mrtm_plan *pl, int nm = 1) {
    if (nm <= 1) {
        if (!pl->skew_tried) {
            pl->skew_tried = true;
            const char *ek = getenv("XANTHOS_MRTM_SKEW_K");
            pl->skew = build_skew(pl, ek ? std::max(1, std::min(6, atoi(ek))) : 2);
        }
        return pl->skew;
    }
    if (!pl->skew_multi_tried) {
        pl->skew_multi_tried = true;
        const char *ek = getenv("XANTHOS_MRTM_SKEW_KM");
        pl->skew_multi = build_skew(pl, ek ? std::max(1, std::min(6, atoi(ek))) : 4);
    }
    return pl->skew_multi;
}


// =============================================================================================
// device
// =============================================================================================
constexpr int SK_WINDOW_DEFAULT = 1;   // pacing window in months (0: off; raised to the smallest deadlock-free value at launch)
constexpr int SK_NM_MAX = 4;     // members per launch (co-resident blocks per SM)

struct SkewArgs {
    const int *cell, *lag, *src, *ghost_edge, *ghost_lag, *exp_edge, *exp_place, *Dw, *edge_prod, *edge_cons;
    int *progress;               // [nw] sub-steps handed over / consumed (see the block prologue)
    double2 *ring;               // [n_edges][RL] (F, F') by absolute sub-step
    const double *runoff, *chs_prev, *flow_dist, *velocity, *area;
    const int *step_start;       // [M + 1] first absolute sub-step of every routing step (spin-up months, then months)
    const int *step_nt;          // [M]
    const int *step_month;       // [M] runoff month read by the step
    const double *step_secs;     // [M]
    double *chs, *avg, *instream;
    int nw, M, T, spinup, ld, RL, G, O, sleep_ns;   // G / O: most ghost entries / exports of a warp
    double dt;
    long long *dbg;              // optional [nw][5]: cycles total, prologue wait, events, SM sub-partition, slow iterations
    // further members of a multi-member launch (member m = blocks [m, m + 1) * blocks_per_member of the grid): the same
    // plan, calendar and static arrays, their own series, rings and counters
    struct Member {
        const double *runoff, *chs_prev;
        double *chs, *avg, *instream;
        double2 *ring;
        int *progress, *done;
    } more[SK_NM_MAX - 1];
    int *done;                   // [M + 2] per member: warps that have crossed boundary b (pacing window), or null
    int window;                  // a warp crosses boundary b only after every warp has crossed b - window (0: no pacing)
    int blocks_per_member, rotate;   // rotate: shift of the block -> warp-set map per member
};

__device__ __forceinline__ int sk_ld_relaxed_pred(const int *p, int pred, int dflt) {
    int v = dflt;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.relaxed.gpu.global.s32 %0, [%1];\n\t}"
                 : "+r"(v)
                 : "l"(p), "r"(pred)
                 : "memory");
    return v;
}
__device__ __forceinline__ void sk_st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double2 sk_lds(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ double sk_lds1(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sk_sts(unsigned addr, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void sk_sts1(unsigned addr, double x) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(x) : "memory");
}
__device__ __forceinline__ void sk_sts1_pred(unsigned addr, double x, int pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.shared.f64 [%0], %1;\n\t}" ::"r"(addr), "d"(x),
                 "r"(pred)
                 : "memory");
}
__device__ __forceinline__ void sk_sts_pred(unsigned addr, double x, double y, int pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\t@p st.shared.v2.f64 [%0], {%1, %2};\n\t}" ::"r"(addr),
                 "d"(x), "d"(y), "r"(pred)
                 : "memory");
}

template <int K>
struct SkewLane {
    double S[K], ti[K], erl[K], fav[K];
    double erln[K], pend[K], qn[K], ar[K];   // lateral inflow of the next step, monthly sum waiting for its division,
                                             // prefetched runoff of the step after the next, cell area
    int cell[K], lag[K];
    // shared-memory byte ADDRESSES in table buffer 0 (F half): the terms of the wide row, the single tributary of the cells
    // in slots 1 .. K-1, this lane's own entry of slot 0.  A buffer offset is added at the use; with lag 1 it is a compile-
    // time constant, so every table access is register + immediate.
    unsigned sb[2 * SK_NB];
    unsigned su[K > 1 ? K - 1 : 1];
    unsigned mp;
};

// Row terms of one iteration: trial flows (x) and, when needed, final flows (y) of the tributaries.
template <int K>
struct SkewTerms {
    double tx[2 * SK_NB], ty[2 * SK_NB], ux[K > 1 ? K - 1 : 1], uy[K > 1 ? K - 1 : 1];
};

// Loads the row terms of an iteration from table buffer LDB (shared-memory byte address).  A buffer holds the trial
// flows F of all entries (8 B each) followed, FPOFF bytes further, by the final flows F'.  `slow`: some flow in LDB
// has F' != F (a clamp, mrtm.py:54-60); otherwise the F' half is neither read nor used.
template <int K>
__device__ __forceinline__ void skew_load(SkewTerms<K> &R, const SkewLane<K> &L, const unsigned LDB, const unsigned FPOFF,
                                          const bool slow) {   // LDB: byte offset of the buffer within the table
#pragma unroll
    for (int j = 0; j < 2 * SK_NB; ++j) R.tx[j] = sk_lds1(L.sb[j] + LDB);
#pragma unroll
    for (int s = 1; s < K; ++s) R.ux[s - 1] = sk_lds1(L.su[s - 1] + LDB);
    if (slow) {
#pragma unroll
        for (int j = 0; j < 2 * SK_NB; ++j) R.ty[j] = sk_lds1(L.sb[j] + (LDB + FPOFF));
#pragma unroll
        for (int s = 1; s < K; ++s) R.uy[s - 1] = sk_lds1(L.su[s - 1] + (LDB + FPOFF));
    }
}

// The balances of the K slots of one iteration, from terms loaded earlier; stores (F, F') of every cell into table
// buffer WR.  `slow`: the terms carry F' != F somewhere: the balance for the update (mrtm.py:68) is then a second sum
// over the F'; otherwise it EQUALS the balance for the decision (mrtm.py:51) bit for bit and is not formed.
// xdiff: this lane wrote a ghost entry with F' != F into WR.  Returns (warp-uniform) whether any flow in WR has
// F' != F.  The only values carried from one iteration to the next are the storages S (and the monthly sums).
template <int K>
__device__ __forceinline__ bool skew_compute_store(SkewLane<K> &L, const SkewTerms<K> &R, const unsigned WR,
                                                   const unsigned FPOFF, const int lane, const double dt,
                                                   const double dtinv, const bool slow, const bool xdiff) {
    double F[K], d[K], ddt[K], Sn[K];
    // ---- slot 0: up to SK_NB tributaries before and SK_NB after the cell's own column (mrtm.py:51 in CSR order) ----
    {
        F[0] = L.S[0] * L.ti[0];                                                    // mrtm.py:50
        double acc = R.tx[0];
#pragma unroll
        for (int j = 1; j < SK_NB; ++j) acc = acc + R.tx[j];
        acc = acc - F[0];
#pragma unroll
        for (int j = SK_NB; j < 2 * SK_NB; ++j) acc = acc + R.tx[j];
        d[0] = acc + L.erl[0];                                                      // balance with the trial flows
    }
    // ---- slots 1 .. K-1: at most one tributary; a two-term row is order-free ---------------------------------------
#pragma unroll
    for (int s = 1; s < K; ++s) {
        F[s] = L.S[s] * L.ti[s];
        d[s] = (R.ux[s - 1] - F[s]) + L.erl[s];
    }
#pragma unroll
    for (int s = 0; s < K; ++s) {
        ddt[s] = d[s] * dt;
        Sn[s] = L.S[s] + ddt[s];                                                    // :76
    }
    if (slow) {                                                                     // balance with the final flows, :68-69
        double acc = R.ty[0];
#pragma unroll
        for (int j = 1; j < SK_NB; ++j) acc = acc + R.ty[j];
        acc = acc - F[0];
#pragma unroll
        for (int j = SK_NB; j < 2 * SK_NB; ++j) acc = acc + R.ty[j];
        Sn[0] = L.S[0] + (acc + L.erl[0]) * dt;
#pragma unroll
        for (int s = 1; s < K; ++s) Sn[s] = L.S[s] + ((R.uy[s - 1] - F[s]) + L.erl[s]) * dt;
    }
    double Fp[K];
    bool any_clamp = false;
#pragma unroll
    for (int s = 0; s < K; ++s) {
        const bool clamp = ddt[s] < (-L.S[s]);                                      // :54
        const double Fc = (d[s] + F[s]) + L.S[s] * dtinv;                           // :60
        Fp[s] = clamp ? Fc : F[s];
        L.S[s] = clamp ? 0.0 : Sn[s];                                               // :63
        L.fav[s] = L.fav[s] + Fp[s];                                                // :78
        any_clamp = any_clamp || clamp;
    }
    const bool changed = __any_sync(0xffffffffu, any_clamp || xdiff);
#pragma unroll
    for (int s = 0; s < K; ++s) {
        sk_sts1(L.mp + (WR + (unsigned)s * 256u), F[s]);               // WR: byte offset of the buffer within the table
        sk_sts1(L.mp + (WR + FPOFF + (unsigned)s * 256u), Fp[s]);
    }
    return changed;
}

// Lag 1, the form that measured fastest: loads, balances and stores of one iteration in one piece, the balances slot by
// slot (42.7 ms with 2 cells per lane against 47.4 ms for the load-ahead form above, which exists for lag 2).
// RDO / WRO: byte offsets of the table buffers read / written (compile-time constants at the call sites).
// LAZY: the F' half is written only when it differs somewhere (launches of several members, see below).
template <int K, bool SLOW, bool LAZY>
__device__ __forceinline__ bool skew_iter(SkewLane<K> &L, const unsigned RDO, const unsigned WRO, const unsigned FPOFF,
                                          const double dt, const double dtinv, const bool xdiff) {
    double tx[2 * SK_NB], ty[2 * SK_NB], ux[K > 1 ? K - 1 : 1], uy[K > 1 ? K - 1 : 1];
#pragma unroll
    for (int j = 0; j < 2 * SK_NB; ++j) tx[j] = sk_lds1(L.sb[j] + RDO);
#pragma unroll
    for (int s = 1; s < K; ++s) ux[s - 1] = sk_lds1(L.su[s - 1] + RDO);
    if (SLOW) {
#pragma unroll
        for (int j = 0; j < 2 * SK_NB; ++j) ty[j] = sk_lds1(L.sb[j] + (RDO + FPOFF));
#pragma unroll
        for (int s = 1; s < K; ++s) uy[s - 1] = sk_lds1(L.su[s - 1] + (RDO + FPOFF));
    }
    double Fo[K], Fpo[K];
    bool any_clamp = false;
    {   // slot 0: up to SK_NB tributaries before and SK_NB after the cell's own column (mrtm.py:51 in CSR order)
        const double F = L.S[0] * L.ti[0];                                          // mrtm.py:50
        double d = tx[0], d2 = SLOW ? ty[0] : 0.0;
#pragma unroll
        for (int j = 1; j < SK_NB; ++j) {
            d = d + tx[j];
            if (SLOW) d2 = d2 + ty[j];
        }
        d = d - F;
        if (SLOW) d2 = d2 - F;
#pragma unroll
        for (int j = SK_NB; j < 2 * SK_NB; ++j) {
            d = d + tx[j];
            if (SLOW) d2 = d2 + ty[j];
        }
        d = d + L.erl[0];                                                           // balance with the trial flows
        if (SLOW) d2 = d2 + L.erl[0];                                               // balance with the final flows, :68
        const double ddt = d * dt;
        const bool clamp = ddt < (-L.S[0]);                                         // :54
        const double Sn = L.S[0] + (SLOW ? d2 * dt : ddt);                          // :69 / :76
        const double Fc = (d + F) + L.S[0] * dtinv;                                 // :60
        const double Fp = clamp ? Fc : F;
        L.S[0] = clamp ? 0.0 : Sn;                                                  // :63
        L.fav[0] = L.fav[0] + Fp;                                                   // :78
        Fo[0] = F;
        Fpo[0] = Fp;
        any_clamp = clamp;
    }
#pragma unroll
    for (int s = 1; s < K; ++s) {   // at most one tributary; a two-term row is order-free
        const double F = L.S[s] * L.ti[s];
        const double d = (ux[s - 1] - F) + L.erl[s];
        const double ddt = d * dt;
        const bool clamp = ddt < (-L.S[s]);
        double Sn;
        if (SLOW) {
            const double d2 = (uy[s - 1] - F) + L.erl[s];
            Sn = L.S[s] + d2 * dt;
        } else {
            Sn = L.S[s] + ddt;
        }
        const double Fc = (d + F) + L.S[s] * dtinv;
        const double Fp = clamp ? Fc : F;
        L.S[s] = clamp ? 0.0 : Sn;
        L.fav[s] = L.fav[s] + Fp;
        Fo[s] = F;
        Fpo[s] = Fp;
        any_clamp = any_clamp || clamp;
    }
    const bool changed = __any_sync(0xffffffffu, any_clamp || xdiff);   // before the stores: its latency hides behind them
#pragma unroll
    for (int s = 0; s < K; ++s) sk_sts1(L.mp + (WRO + (unsigned)s * 256u), Fo[s]);
    // LAZY: the F' half is written only when some flow of this iteration has F' != F: nobody reads it otherwise (the
    // next iteration takes the fast path, an export lane copies F, the final instream flow is read from the F half) - a
    // quarter of the shared-memory stores of a fast iteration.  Measured: 2 % faster with two members per launch
    // (32.1 -> 31.5 ms per member), 3 % slower with one (the stores wait for the vote), hence the switch.
    if (!LAZY || changed) {
#pragma unroll
        for (int s = 0; s < K; ++s) sk_sts1(L.mp + (WRO + FPOFF + (unsigned)s * 256u), Fpo[s]);
    }
    return changed;
}

// Shared memory of a block: per warp a fixed part - flow table [SK_LAGM + 1 buffers][F half | F' half] (NE entries each),
// export series [O][SK_CH] (F, F'), one 16-byte dump slot per lane - followed by the staged series [SK_W] (F, F') of the
// ghost entries of all its warps, packed (a warp has 0 .. SK_XG of them; the block's total is what is allocated).
template <int K>
struct SkewSmem {
    static constexpr int CAP = 32 * K, NE = CAP + SK_XG + 2;     // + all-zero entry, + 1 keeps the F' half 16-byte aligned
    static constexpr unsigned FPOFF = NE * 8u, PSTRIDE = NE * 16u;
    // ... + one 16-byte dump slot per lane (target of the import / export stores of lanes that have neither)
    static __host__ __device__ int bytes_fixed(int O) { return (SK_LAGM + 1) * NE * 16 + O * SK_CH * 16 + 32 * 16; }
};

template <int K, bool LINKED, bool LAZY>
__device__ __forceinline__ void skew_run(const SkewArgs &a, const int w, const int lane, const unsigned EX0,
                                         const unsigned ST0 /* staged series of this warp's ghost entries */) {
    using SM = SkewSmem<K>;
    constexpr int CAP = SM::CAP, NS = 2 * SK_NB + K - 1;
    constexpr unsigned FPOFF = SM::FPOFF, PSTRIDE = SM::PSTRIDE;
    const unsigned full = 0xffffffffu;
    const unsigned XS0 = EX0 + (SK_LAGM + 1) * PSTRIDE;             // export series
    const bool dbg = a.dbg != nullptr;
    const long long cyc0 = dbg ? clock64() : 0;
    long long cyc_wait = 0, cyc_evt = 0;
    int n_slow = 0;
    const int Dw = a.Dw[w], M = a.M, T = a.T, RL = a.RL;
    const double dt = a.dt, dtinv = 1. / a.dt;                                      // mrtm.py:43
    SkewLane<K> L;
#pragma unroll
    for (int s = 0; s < K; ++s) {
        const int c = a.cell[(size_t)w * CAP + s * 32 + lane];
        L.cell[s] = c;
        L.lag[s] = a.lag[(size_t)w * CAP + s * 32 + lane];
        L.S[s] = 0.0; L.ti[s] = 0.0; L.erl[s] = 0.0; L.fav[s] = 0.0; L.qn[s] = 0.0; L.ar[s] = 0.0;
        L.erln[s] = 0.0; L.pend[s] = 0.0;
        if (c >= 0) {
            L.ti[s] = a.velocity[c] / a.flow_dist[c];                               // mrtm.py:42
            L.ar[s] = a.area[c];
            L.erln[s] = ((a.runoff[(size_t)a.step_month[0] * a.ld + c] * L.ar[s]) * (1e6 / 1e3)) / a.step_secs[0];   // mrtm.py:45
            if (a.M > 1) L.qn[s] = a.runoff[(size_t)a.step_month[1] * a.ld + c];
        }
    }
    {
        const int *row = a.src + ((size_t)w * 32 + lane) * NS;
#pragma unroll
        for (int j = 0; j < 2 * SK_NB; ++j) L.sb[j] = EX0 + (unsigned)row[j] * 8u;
#pragma unroll
        for (int s = 1; s < K; ++s) L.su[s - 1] = EX0 + (unsigned)row[2 * SK_NB + s - 1] * 8u;
        L.mp = EX0 + (unsigned)lane * 8u;
    }
    // ---- ghost imports (lanes 0 .. 15) and exports (lanes 16 .. 31) ------------------------------------------------
    const int xi = lane & 15;
    const bool imp_lane = lane < SK_XG;
    int xedge = -1;
    if (LINKED) xedge = imp_lane ? a.ghost_edge[(size_t)w * SK_XG + xi] : a.exp_edge[(size_t)w * SK_XO + xi];
    const bool imp = imp_lane && xedge >= 0, expo = !imp_lane && xedge >= 0;
    const int glag = imp ? a.ghost_lag[(size_t)w * SK_XG + xi] : 0;
    const unsigned eplace8 = EX0 + (unsigned)(expo ? a.exp_place[(size_t)w * SK_XO + xi] : CAP + SK_XG) * 8u;
    const unsigned ximp0 = EX0 + (unsigned)(CAP + xi) * 8u;   // this import lane's ghost entry in buffer 0
    double2 *const xring = a.ring + (size_t)(xedge >= 0 ? xedge : 0) * RL;
    const int peer = imp ? a.edge_prod[xedge] : (expo ? a.edge_cons[xedge] : 0);
    const unsigned ghost_mask = __ballot_sync(full, imp), exp_mask = __ballot_sync(full, expo);
    const unsigned stg_lane = ST0 + (unsigned)xi * (SK_W * 16u);                    // this import lane's staged series
    const unsigned xs_lane = XS0 + (unsigned)xi * (SK_CH * 16u);                    // this export lane's series
    const unsigned dump_lane = XS0 + (unsigned)a.O * (SK_CH * 16u) + (unsigned)lane * 16u;
    // second half of an (F, F') pair: 8 bytes further in a staged series, FPOFF bytes further in the flow table
    const unsigned xdelta = imp ? 8u : FPOFF;

    // Hand-over test.  The counters of the peers are remembered: a producer that runs far ahead (leaves are not
    // throttled by anything but the ring) or a consumer that is close behind satisfies many prologues with one read,
    // and a read is an L2 round trip plus a fence.
    int seen_p = 0, seen_c = 0;
    auto wait_peers = [&](int need_p, int need_c) {
        const int poll_p = imp ? 1 : 0, poll_c = (expo && need_c > 0) ? 1 : 0;
        if (__all_sync(full, (!poll_p || seen_p >= need_p) && (!poll_c || seen_c >= need_c))) return;
        const int *pp = a.progress + peer;
        for (;;) {
            const int vp = sk_ld_relaxed_pred(pp, poll_p, need_p), vc = sk_ld_relaxed_pred(pp, poll_c, need_c);
            if (__all_sync(full, vp >= need_p && vc >= need_c)) {
                if (poll_p) seen_p = vp;
                if (poll_c) seen_c = vc;
                break;
            }
            __nanosleep(a.sleep_ns);
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");   // acquire side of the hand-over
        __syncwarp();
    };
    auto stage_chunk = [&](int tau0) {             // ring entries tau0 .. tau0 + SK_CH - 1 of every ghost -> shared memory
        unsigned gm = ghost_mask;
        while (gm) {
            const int g = __ffs(gm) - 1;
            gm &= gm - 1;
            const double2 *rp = reinterpret_cast<const double2 *>(__shfl_sync(full, (unsigned long long)xring, g));
#pragma unroll
            for (int h = 0; h < SK_CH / 32; ++h) {
                const int tau = tau0 + h * 32 + lane;
                if (tau < T) {
                    const unsigned dst = ST0 + ((unsigned)g * SK_W + (unsigned)(tau & (SK_W - 1))) * 16u;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(rp + (tau & (RL - 1))) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // exports of the SK_CH iterations before n0 (sub-steps n0 - SK_CH - 1 - Dw ...), shared memory -> ring, coalesced
    auto flush_exports = [&](int n0) {
        unsigned em = exp_mask;
        while (em) {
            const int l = __ffs(em) - 1;
            em &= em - 1;
            double2 *rp = reinterpret_cast<double2 *>(__shfl_sync(full, (unsigned long long)xring, l));
            const unsigned src = XS0 + (unsigned)(l & 15) * (SK_CH * 16u);
#pragma unroll
            for (int h = 0; h < SK_CH / 32; ++h) {
                const int i = h * 32 + lane, tau = n0 - SK_CH + i - 1 - Dw;
                if (tau >= 0 && tau < T) rp[tau & (RL - 1)] = sk_lds(src + (unsigned)i * 16u);
            }
        }
    };
    if (LINKED && ghost_mask) {
        wait_peers(min(T, SK_CH), 0);
        stage_chunk(0);
    }

    // Boundary b = start of routing step b (b = M: end of the run).  A cell crosses it `lag` iterations after the window
    // opens.  Everything the event code needs is in registers before the window opens (a global load or a division
    // inside the window stalls the warp): at the crossing a cell only stores its storage, parks its monthly sum and
    // switches to the lateral inflow prepared for the new step; the divisions (monthly mean, mrtm.py:80; lateral
    // inflow of the step after, :45) are done for all slots together when the window closes.
    int b = 0, evt = a.step_start[0];
    double ev_nt_prev = 1.0, ev_secs_next = a.step_secs[M > 1 ? 1 : 0];
    size_t ev_next2_off = (size_t)a.step_month[M > 2 ? 2 : 0] * a.ld, ev_out_off = 0;
    const bool paced = a.window > 0;
    bool ch1 = false, ch2 = false;   // some flow written one / two iterations ago has F' != F
    auto events = [&](const int n, const unsigned LAST) {   // LAST: byte offset of the table buffer written one iteration ago
        const long long c0 = dbg ? clock64() : 0;
        const int k = n - evt;
        const bool st = b > 0 && b - 1 >= a.spinup;
#pragma unroll
        for (int s = 0; s < K; ++s) {
            const int c = L.cell[s];
            if (c >= 0 && L.lag[s] == k) {
                if (st && a.chs) {
                    if (paced) a.chs[ev_out_off + c] = L.S[s];
                    else stg_stream(a.chs + ev_out_off + c, L.S[s]);
                }
                if (b == M && a.instream)   // the final flow F' of the last sub-step (= F when that iteration was a fast one)
                    a.instream[c] = sk_lds1(L.mp + (LAST + ((!LAZY || SK_LAGM == 2 || ch1) ? FPOFF : 0u) + (unsigned)s * 256u));
                L.pend[s] = L.fav[s];
                L.fav[s] = 0.0;
                L.erl[s] = L.erln[s];
                if (b == 0) L.S[s] = a.chs_prev ? a.chs_prev[c] : 0.0;
            }
        }
        if (k == Dw) {   // every cell has crossed: window closed
#pragma unroll
            for (int s = 0; s < K; ++s) {
                const int c = L.cell[s];
                if (c >= 0) {
                    if (st && a.avg) {                                                              // mrtm.py:80
                        if (paced) a.avg[ev_out_off + c] = L.pend[s] / ev_nt_prev;
                        else stg_stream(a.avg + ev_out_off + c, L.pend[s] / ev_nt_prev);
                    }
                    if (b + 1 < M) L.erln[s] = ((L.qn[s] * L.ar[s]) * (1e6 / 1e3)) / ev_secs_next;  // mrtm.py:45
                    if (b + 2 < M) L.qn[s] = a.runoff[ev_next2_off + c];                            // used a month later
                }
            }
            if (paced) {
                // Pacing: independent river trees are not coupled by any ring, and the warps of small trees run up to twice
                // as fast as the slowest chain.  Cells that share a 32-byte sector of a month row then touch it months
                // apart - each sector of the runoff is fetched from DRAM up to four times and every output sector
                // written back half-filled (3.6 GB per launch instead of 1.55).  A warp crosses boundary b only after ALL
                // warps have crossed b - window: the month rows in flight stay in L2.  The slowest warp never waits here.
                if (lane == 0) atomicAdd(a.done + b, 1);
                const int need = b - a.window;
                if (need >= 0) {
                    for (;;) {
                        const int v = sk_ld_relaxed_pred(a.done + need, 1, 0);
                        if (__all_sync(full, v >= a.nw)) break;
                        __nanosleep(2000);
                    }
                }
            }
            ++b;   // constants of the next window, fetched a month ahead of their use
            evt = (b <= M) ? a.step_start[b] : 0x7fffffff;
            if (b <= M) ev_nt_prev = (double)a.step_nt[b - 1];
            if (b + 1 < M) ev_secs_next = a.step_secs[b + 1];
            if (b + 2 < M) ev_next2_off = (size_t)a.step_month[b + 2] * a.ld;
            ev_out_off = (size_t)max(0, b - 1 - a.spinup) * a.ld;
        }
        __syncwarp();
        if (dbg) cyc_evt += clock64() - c0;
    };

    double xvx = 0.0, xvy = 0.0;     // ghost import / export value of the coming iteration
    // Table buffers: iteration n writes bw = buffer n mod 3; a cell reads what its tributaries (SK_LAGM = 2 iterations
    // ahead in their own time) wrote two iterations ago.  The terms of iteration n + 1 come from bl (written in n - 1)
    // and are loaded while iteration n is being computed: no load, store or vote latency is on the path from one
    // iteration to the next - only the storages S in registers.
    unsigned bw = 0, bl = SK_LAGM * PSTRIDE, bt = PSTRIDE;   // lag 2 only: byte offsets of the buffers written now / one / two iterations ago
    SkewTerms<K> TA, TB;
#pragma unroll
    for (int j = 0; j < 2 * SK_NB; ++j) TA.tx[j] = TA.ty[j] = TB.tx[j] = TB.ty[j] = 0.0;
#pragma unroll
    for (int s = 1; s < K; ++s) TA.ux[s - 1] = TA.uy[s - 1] = TB.ux[s - 1] = TB.uy[s - 1] = 0.0;
    // One iteration: computes with the terms in CUR, loads the terms of the next iteration into NXT; WRO is the byte
    // offset of the buffer this iteration writes (lag 1: a compile-time constant, 0 or PSTRIDE).  The import and export
    // stores are unconditional: every lane has a destination (ghost entry halves / export series slot / dump slot) -
    // predicated stores were turned into branches by ptxas.
    auto iteration = [&](const int n, const SkewTerms<K> &CUR, SkewTerms<K> &NXT, const unsigned WRO) {
        bool xdiff = false;
        if (LINKED) {   // the value was loaded at the end of the previous iteration
            const unsigned x0 = imp_lane ? ximp0 + WRO
                                         : (expo ? xs_lane + ((unsigned)(n & (SK_CH - 1)) << 4) : dump_lane);
            sk_sts1(x0, xvx);
            sk_sts1(x0 + (imp_lane ? FPOFF : 8u), xvy);
            xdiff = imp && (__double_as_longlong(xvx) != __double_as_longlong(xvy));
        }
        bool ch0;
        if (SK_LAGM == 2) {
            skew_load<K>(NXT, L, bl, FPOFF, ch1);
            ch0 = skew_compute_store<K>(L, CUR, bw, FPOFF, lane, dt, dtinv, ch2, xdiff);
            if (ch2) ++n_slow;
        } else {
            const unsigned RDO = WRO ^ PSTRIDE;     // the other of the two buffers
            if (ch1) {
                ch0 = skew_iter<K, true, LAZY>(L, RDO, WRO, FPOFF, dt, dtinv, xdiff);
                ++n_slow;
            } else {
                ch0 = skew_iter<K, false, LAZY>(L, RDO, WRO, FPOFF, dt, dtinv, xdiff);
            }
        }
        if (LINKED) {   // for iteration n + 1: staged entry n + 1 - lag (lag >= 1: its chunk has landed), or the cell just stored
            const unsigned xaddr = imp ? stg_lane + ((unsigned)((n + 1 - glag) & (SK_W - 1)) << 4)
                                       : eplace8 + (SK_LAGM == 2 ? bw : WRO);
            xvx = sk_lds1(xaddr);
            xvy = sk_lds1(xaddr + xdelta);
            if (LAZY && SK_LAGM == 1 && !imp && !ch0) xvy = xvx;   // exported cell of a fast iteration: its F' half was not written
        }
        ch2 = ch1;
        ch1 = ch0;
        if (SK_LAGM == 2) {
            const unsigned t = bt;
            bt = bl;
            bl = bw;
            bw = t;
        }
    };
    // buffer written by even / odd iterations (lag 1); lag 2 rotates three buffers at run time
    constexpr unsigned WE = 0u, WO = PSTRIDE;

    const int nlast = T + Dw;
    for (int n0 = 0; n0 <= nlast; n0 += SK_CH) {
        if (LINKED) {
            const long long c0 = dbg ? clock64() : 0;
            asm volatile("cp.async.wait_group 0;" ::: "memory");                    // chunk n0 / SK_CH has landed
            __syncwarp();
            // hand-over: after the flush every export up to sub-step n0 - 2 - Dw is in the ring; every ring entry below
            // n0 + SK_CH of the edges this warp consumes is in shared memory
            wait_peers(min(T, n0 + 2 * SK_CH), n0 - RL);
            if (exp_mask) flush_exports(n0);
            __syncwarp();   // every lane's ring reads and writes are ordered before lane 0's release
            if (lane == 0) sk_st_release(a.progress + w, max(0, min(T, n0 - 1 - Dw)));
            if (ghost_mask) stage_chunk(n0 + SK_CH);
            if (dbg) cyc_wait += clock64() - c0;
        }
        int i = 0;
        while (i < SK_CH) {
            // pairs of iterations before the next month window opens run without any event test
            const int lim = min(SK_CH, max(0, evt - n0)) & ~1;
            if (i < lim) {
#pragma unroll 1
                for (; i < lim; i += 2) {
                    iteration(n0 + i, TA, TB, WE);
                    iteration(n0 + i + 1, TB, TA, WO);
                }
            } else {             // a month boundary is being crossed by some cells (Dw + 1 iterations per month)
                const int n = n0 + i;
                if (n >= evt) events(n, SK_LAGM == 2 ? bl : WO);
                iteration(n, TA, TB, WE);
                if (n + 1 >= evt) events(n + 1, SK_LAGM == 2 ? bl : WE);
                iteration(n + 1, TB, TA, WO);
                i += 2;
            }
        }
    }
    if (LINKED) {
        if (exp_mask) flush_exports(((nlast / SK_CH) + 1) * SK_CH);
        __syncwarp();
        if (lane == 0) sk_st_release(a.progress + w, T);
    }
    if (dbg && lane == 0) {
        unsigned smid, wid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        a.dbg[5 * w] = clock64() - cyc0;
        a.dbg[5 * w + 1] = cyc_wait;
        a.dbg[5 * w + 2] = cyc_evt;
        a.dbg[5 * w + 3] = smid * 4 + (wid & 3);
        a.dbg[5 * w + 4] = n_slow;
    }
}

// NM = members per launch.  NM = 2: two blocks per SM (launch bound), block b serves member b / blocks_per_member: the
// warps of the second member fill the issue slots the first one leaves idle (one or two in-order warps per SM
// sub-partition wait 60 - 70 % of the cycles on fixed-latency dependencies).
template <int K, bool LAZY>
__device__ __forceinline__ void skew_block(const SkewArgs &a, const int bx, const int blocks) {
    extern __shared__ __align__(16) unsigned char sk_smem[];
    __shared__ int s_ghosts[17];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int w = wib * blocks + bx;                // consecutive plan warps on different SMs
    const bool live = w < a.nw;
    bool has_ghost = false, linked = false;
    if (live && lane < SK_XG) {
        has_ghost = a.ghost_edge[(size_t)w * SK_XG + lane] >= 0;
        linked = has_ghost || a.exp_edge[(size_t)w * SK_XO + lane] >= 0;
    }
    const int ng = __popc(__ballot_sync(0xffffffffu, has_ghost));
    if (lane == 0) s_ghosts[wib] = ng;
    __syncthreads();                                // the only block-level barrier: before the time loop
    if (!live) return;
    int g0 = 0;
    for (int k = 0; k < wib; ++k) g0 += s_ghosts[k];
    const int fixed = SkewSmem<K>::bytes_fixed(a.O);
    const unsigned EX0 = (unsigned)__cvta_generic_to_shared(sk_smem + (size_t)wib * fixed);
    const unsigned ST0 = (unsigned)__cvta_generic_to_shared(sk_smem + (size_t)wpb * fixed) + (unsigned)g0 * (SK_W * 16u);
    for (int i = lane; i < fixed / 16; i += 32) sk_sts(EX0 + (unsigned)i * 16u, 0.0, 0.0);
    for (int i = lane; i < ng * SK_W; i += 32) sk_sts(ST0 + (unsigned)i * 16u, 0.0, 0.0);
    __syncwarp();
    if (__any_sync(0xffffffffu, linked)) skew_run<K, true, LAZY>(a, w, lane, EX0, ST0);
    else skew_run<K, false, LAZY>(a, w, lane, EX0, ST0);
}

// NM = members per launch = blocks per SM (launch bound).  The warps of the other members fill the issue slots one
// member leaves idle: one or two in-order warps per SM sub-partition wait 60 - 70 % of the cycles on fixed-latency
// dependencies.  MAXT = threads per block the variant is compiled for.
template <int K, int NM, int MAXT>
__global__ void __launch_bounds__(MAXT, NM) mrtm_skew_kernel(const SkewArgs a0) {
    if constexpr (NM == 1) {
        skew_block<K, false>(a0, blockIdx.x, gridDim.x);
    } else {
        // one copy of the code for all members (two copies thrash the instruction cache: measured 6 x the requests)
        const int bpm = a0.blocks_per_member, m = (int)blockIdx.x / bpm;
        SkewArgs a = a0;
        if (m > 0) {
            const SkewArgs::Member &x = a0.more[m - 1];
            a.runoff = x.runoff; a.chs_prev = x.chs_prev; a.chs = x.chs; a.avg = x.avg; a.instream = x.instream;
            a.ring = x.ring; a.progress = x.progress; a.done = x.done; a.dbg = nullptr;
        }
        // member m takes the plan's warp sets in rotated order: block i of every member tends to land on the same SM, and
        // the same plan warp of two members is equally heavy (the clamping cells are a property of the network) - the
        // rotation pairs a heavy warp with some other warp instead of its twin
        int bx = (int)blockIdx.x - m * bpm + m * a0.rotate;
        if (bx >= bpm) bx -= bpm;
        skew_block<K, true>(a, bx, bpm);
    }
}

template <typename V>
static bool sk_upload(const std::vector<V> &h, V **d) {
    const size_t bytes = sizeof(V) * std::max<size_t>(h.size(), 1);
    if (cudaMalloc((void **)d, bytes) != cudaSuccess) return false;
    if (!h.empty() && cudaMemcpy(*d, h.data(), sizeof(V) * h.size(), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    return true;
}

static int skew_geometry(const SkewPlan *sp, int K, int sms, int *blocks, int *wpb) {
    *blocks = std::min(sms, sp->nw);
    *wpb = std::max(4, ((ceil_div(sp->nw, *blocks) + 3) / 4) * 4);
    return *wpb > (K == 1 ? 16 : 8) ? XAN_E_INVALID : XAN_OK;
}

template <int K, int NM, int MAXT>
static int launch_skew(SkewPlan *sp, SkewArgs &a, int sms, cudaStream_t s) {
    auto kernel = mrtm_skew_kernel<K, NM, MAXT>;
    const int fixed = SkewSmem<K>::bytes_fixed(sp->O);
    int blocks = 0, wpb = 0;
    if (skew_geometry(sp, K, sms, &blocks, &wpb) != XAN_OK || wpb * 32 > MAXT) return XAN_E_INVALID;
    // ghost entries of the warps of a block (warp w runs as warp w / blocks of block w % blocks)
    int max_g = 0;
    for (int b = 0; b < blocks; ++b) {
        int g = 0;
        for (int w = b; w < sp->nw; w += blocks)
            for (int k = 0; k < SK_XG; ++k) g += sp->ghost_edge[(size_t)w * SK_XG + k] >= 0 ? 1 : 0;
        max_g = std::max(max_g, g);
    }
    const size_t smem = (size_t)fixed * wpb + (size_t)max_g * SK_W * 16;
    if (smem > 220 * 1024) return XAN_E_INVALID;
    XAN_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    int per_sm = 0;
    XAN_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, wpb * 32, smem));
    // every block of every member must be resident (the cut-edge pipeline needs every warp alive)
    if (per_sm < 1 || blocks * wpb < sp->nw || (long long)per_sm * sms < (long long)NM * blocks) return XAN_E_INVALID;
    a.blocks_per_member = blocks;
    {
        const char *er = getenv("XANTHOS_MRTM_SKEW_ROTATE");
        a.rotate = er ? std::max(0, atoi(er)) % blocks : blocks / NM;
    }
    void *kargs[] = {(void *)&a};
    // cooperative launch = all blocks co-resident; no grid.sync is used
    XAN_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)kernel, dim3(NM * blocks), dim3(wpb * 32), kargs, smem, s));
    return XAN_OK;
}

// Variants compiled: one member for every K (the register budget of a whole SM); several members for K = 2 (256- and
// 128-thread blocks) and K = 4 (128-thread blocks), where NM blocks fit the register file without (much) spilling.
template <int K>
static int launch_skew_nm(SkewPlan *sp, SkewArgs &a, int nm, int sms, cudaStream_t s) {
    if (nm == 1) return launch_skew<K, 1, (K == 1 ? 512 : 256)>(sp, a, sms, s);
    int blocks = 0, wpb = 0;
    if (skew_geometry(sp, K, sms, &blocks, &wpb) != XAN_OK) return XAN_E_INVALID;
    if constexpr (K == 2) {
        if (wpb == 8 && nm == 2) return launch_skew<2, 2, 256>(sp, a, sms, s);
        if (wpb == 4 && nm == 2) return launch_skew<2, 2, 128>(sp, a, sms, s);
        if (wpb == 4 && nm == 3) return launch_skew<2, 3, 128>(sp, a, sms, s);
        if (wpb == 4 && nm == 4) return launch_skew<2, 4, 128>(sp, a, sms, s);
    }
    if constexpr (K == 4) {
        if (wpb == 4 && nm == 2) return launch_skew<4, 2, 128>(sp, a, sms, s);
        if (wpb == 4 && nm == 3) return launch_skew<4, 3, 128>(sp, a, sms, s);
    }
    if constexpr (K == 3 || K == 5 || K == 6) {
        if (wpb == 4 && nm == 2) return launch_skew<K, 2, 128>(sp, a, sms, s);
    }
    return XAN_E_INVALID;
}

// Effective pacing window in months.  A consumer follows its producer by up to 3 hand-over chunks + the producer's
// largest lag, level after level: the leaves must be allowed that far ahead of the outlets, plus what the rings hold, or
// the pacing and the ring back-pressure wait for each other.  (tests/test_host.py replays the hand-over protocol of the
// kernel on the plan tables with this window - it completes - and with a smaller one - it deadlocks.)
static int skew_window(const SkewPlan *sp, int requested, int nt_min, int RL) {
    if (requested <= 0) return 0;
    const int depth_steps = sp->n_levels * (3 * SK_CH + sp->Dmax + 1) + RL;
    return std::max(requested, ceil_div(depth_steps, std::max(nt_min, 1)) + 2);
}

// Routes nm (1 .. SK_NM_MAX) members with one launch of the skew kernel.  Returns XAN_E_INVALID (without an error
// message) when the plan or the calendar does not allow it (nm > 1: when nm blocks per SM do not fit or the variant is
// not compiled), so that the caller can fall back to fewer members per launch / the warp-dataflow kernel.
int route_skew(xan_mrtm_plan *pl, int nm, const double *const *d_runoff, const double *d_flow_dist,
               const double *d_velocity, const double *d_area, const double *const *d_chs_prev, const int *h_ndays,
               int nmonths, int spinup_months, int ld, double dt, double *const *d_chs, double *const *d_avg,
               double *const *d_instream, int sms, cudaStream_t s) {
    if (nm < 1 || nm > SK_NM_MAX) return XAN_E_INVALID;
    SkewPlan *sp = get_skew(pl, nm);
    if (!sp) return XAN_E_INVALID;
    const int M = spinup_months + nmonths;
    std::vector<int> start(M + 1, 0), nts(M), month(M);
    std::vector<double> secs(M);
    long long total = 0;
    int nt_min = 1 << 30;
    for (int b = 0; b < M; ++b) {
        const int m = b < spinup_months ? b : b - spinup_months;
        month[b] = m;
        nts[b] = (int)((double)h_ndays[m] * 24 * 3600 / dt);                       // mrtm.py:36
        secs[b] = (double)(h_ndays[m] * 24 * 3600);
        nt_min = std::min(nt_min, nts[b]);
        total += nts[b];
        if (total > 0x3fffffff) return XAN_E_INVALID;
        start[b + 1] = (int)total;
    }
    // the event windows of consecutive month boundaries (max lag + 1 iterations) must not overlap
    if (nt_min <= sp->Dmax + 1) return XAN_E_INVALID;
    if (!sp->on_device) {
        std::vector<int> zeros(sp->nw, 0);
        const bool ok = sk_upload(sp->cell, &sp->d_cell) && sk_upload(sp->lag, &sp->d_lag) && sk_upload(sp->src, &sp->d_src) &&
                        sk_upload(sp->ghost_edge, &sp->d_ghost_edge) && sk_upload(sp->ghost_lag, &sp->d_ghost_lag) &&
                        sk_upload(sp->exp_edge, &sp->d_exp_edge) && sk_upload(sp->exp_place, &sp->d_exp_place) &&
                        sk_upload(sp->Dw, &sp->d_Dw) && sk_upload(sp->edge_prod, &sp->d_edge_prod) &&
                        sk_upload(sp->edge_cons, &sp->d_edge_cons) && sk_upload(zeros, &sp->d_progress);
        if (!ok) {
            set_error("mrtm skew plan: CUDA allocation/copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            return XAN_E_CUDA;
        }
        sp->on_device = true;
    }
    SkewArgs a;
    memset(&a, 0, sizeof(a));
    a.cell = sp->d_cell; a.lag = sp->d_lag; a.src = sp->d_src; a.ghost_edge = sp->d_ghost_edge;
    a.ghost_lag = sp->d_ghost_lag; a.exp_edge = sp->d_exp_edge; a.exp_place = sp->d_exp_place; a.Dw = sp->d_Dw;
    a.edge_prod = sp->d_edge_prod; a.edge_cons = sp->d_edge_cons; a.progress = sp->d_progress;
    a.runoff = d_runoff[0]; a.chs_prev = d_chs_prev ? d_chs_prev[0] : nullptr;
    a.flow_dist = d_flow_dist; a.velocity = d_velocity; a.area = d_area;
    a.chs = d_chs ? d_chs[0] : nullptr; a.avg = d_avg ? d_avg[0] : nullptr; a.instream = d_instream ? d_instream[0] : nullptr;
    for (int m = 1; m < nm; ++m) {
        SkewArgs::Member &x = a.more[m - 1];
        x.runoff = d_runoff[m]; x.chs_prev = d_chs_prev ? d_chs_prev[m] : nullptr;
        x.chs = d_chs ? d_chs[m] : nullptr; x.avg = d_avg ? d_avg[m] : nullptr;
        x.instream = d_instream ? d_instream[m] : nullptr;
    }
    a.nw = sp->nw; a.M = M; a.T = (int)total; a.spinup = spinup_months; a.ld = ld; a.G = sp->G; a.O = sp->O; a.dt = dt;
    const char *er = getenv("XANTHOS_MRTM_SKEW_RING"), *es = getenv("XANTHOS_MRTM_SLEEP_NS");
    int RL = er ? atoi(er) : 1024;
    if (RL < 256 || (RL & (RL - 1))) RL = 1024;
    a.RL = RL;
    a.sleep_ns = es ? std::max(0, atoi(es)) : 100;
    // step tables: [start (M+1) | nt (M) | month (M)] ints + [secs (M)] doubles, cached with the plan per calendar
    std::vector<int> key;
    key.push_back(spinup_months);
    key.insert(key.end(), h_ndays, h_ndays + nmonths);
    if (key != sp->cal_key || dt != sp->cal_dt || !sp->d_cal_int) {
        std::vector<int> hint;
        hint.insert(hint.end(), start.begin(), start.end());
        hint.insert(hint.end(), nts.begin(), nts.end());
        hint.insert(hint.end(), month.begin(), month.end());
        // a route with the previous calendar may still be running: the old tables are released in stream order
        if (sp->d_cal_int) cudaFreeAsync(sp->d_cal_int, s);
        if (sp->d_cal_secs) cudaFreeAsync(sp->d_cal_secs, s);
        sp->d_cal_int = nullptr;
        sp->d_cal_secs = nullptr;
        XAN_CUDA_CHECK(scratch_alloc(&sp->d_cal_int, sizeof(int) * hint.size(), s));
        XAN_CUDA_CHECK(scratch_alloc(&sp->d_cal_secs, sizeof(double) * M, s));
        // pageable host memory: cudaMemcpyAsync returns after the data were staged, the vectors may go out of scope
        XAN_CUDA_CHECK(cudaMemcpyAsync(sp->d_cal_int, hint.data(), sizeof(int) * hint.size(), cudaMemcpyHostToDevice, s));
        XAN_CUDA_CHECK(cudaMemcpyAsync(sp->d_cal_secs, secs.data(), sizeof(double) * M, cudaMemcpyHostToDevice, s));
        sp->cal_key = key;
        sp->cal_dt = dt;
    }
    a.step_start = sp->d_cal_int;
    a.step_nt = sp->d_cal_int + (M + 1);
    a.step_month = sp->d_cal_int + (M + 1) + M;
    a.step_secs = sp->d_cal_secs;
    // per launch (and member): concurrent routes on one plan share no mutable device state
    const size_t ring_len = (size_t)std::max(sp->n_edges, 1) * RL;
    double2 *ring = nullptr;
    XAN_CUDA_CHECK(scratch_alloc(&ring, sizeof(double2) * ring_len * nm, s));
    a.ring = ring;
    for (int m = 1; m < nm; ++m) a.more[m - 1].ring = ring + m * ring_len;
    int *progress = nullptr;
    XAN_CUDA_CHECK(scratch_alloc(&progress, sizeof(int) * sp->nw * nm, s));
    XAN_CUDA_CHECK(cudaMemsetAsync(progress, 0, sizeof(int) * sp->nw * nm, s));
    a.progress = progress;
    for (int m = 1; m < nm; ++m) a.more[m - 1].progress = progress + (size_t)m * sp->nw;
    // pacing window in months (XANTHOS_MRTM_SKEW_WINDOW, 0 = off; raised to the smallest safe value)
    const char *ew = getenv("XANTHOS_MRTM_SKEW_WINDOW");
    a.window = ew ? std::max(0, atoi(ew)) : SK_WINDOW_DEFAULT;
    a.window = skew_window(sp, a.window, nt_min, RL);
    int *done = nullptr;
    if (a.window > 0) {
        XAN_CUDA_CHECK(scratch_alloc(&done, sizeof(int) * (size_t)(M + 2) * nm, s));
        XAN_CUDA_CHECK(cudaMemsetAsync(done, 0, sizeof(int) * (size_t)(M + 2) * nm, s));
        a.done = done;
        for (int m = 1; m < nm; ++m) a.more[m - 1].done = done + (size_t)m * (M + 2);
    }
    const char *edbg = getenv("XANTHOS_MRTM_DEBUG");
    if (edbg) XAN_CUDA_CHECK(scratch_alloc(&a.dbg, sizeof(long long) * 5 * sp->nw, s));
    int rc = XAN_E_INVALID;
    switch (sp->K) {
        case 1: rc = launch_skew_nm<1>(sp, a, nm, sms, s); break;
        case 2: rc = launch_skew_nm<2>(sp, a, nm, sms, s); break;
        case 3: rc = launch_skew_nm<3>(sp, a, nm, sms, s); break;
        case 4: rc = launch_skew_nm<4>(sp, a, nm, sms, s); break;
        case 5: rc = launch_skew_nm<5>(sp, a, nm, sms, s); break;
        case 6: rc = launch_skew_nm<6>(sp, a, nm, sms, s); break;
        default: break;
    }
    if (rc == XAN_OK && a.dbg) {
        std::vector<long long> h(5 * (size_t)sp->nw);
        XAN_CUDA_CHECK(cudaMemcpyAsync(h.data(), a.dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, s));
        XAN_CUDA_CHECK(cudaStreamSynchronize(s));
        if (FILE *f = fopen(edbg, "w")) {
            for (int w = 0; w < sp->nw; ++w)
                fprintf(f, "%d %lld %lld %lld %lld %d %lld\n", w, h[5 * w], h[5 * w + 1], h[5 * w + 2], h[5 * w + 3], sp->Dw[w],
                        h[5 * w + 4]);
            fclose(f);
        }
    }
    if (a.dbg) cudaFreeAsync(a.dbg, s);
    cudaFreeAsync(ring, s);
    cudaFreeAsync(progress, s);
    if (done) cudaFreeAsync(done, s);
    return rc;
}

}  // namespace xan

extern "C" {

// Test / diagnostics access to the skew plan (sizes first, then the tables as int arrays; null pointers are skipped).
// info: K, n_warps, n_edges, n_levels, max ghosts per warp, max lag, pieces, sources per lane, zero entry, XG, XO, lag multiplier
int xan_mrtm_skew_info(xan_mrtm_plan *pl, int *info) {
    XAN_REQUIRE(pl && info, "xan_mrtm_skew_info: null pointer");
    xan::SkewPlan *sp = xan::get_skew(pl);
    if (!sp) {
        for (int i = 0; i < 12; ++i) info[i] = 0;
        return XAN_OK;
    }
    const int v[12] = {sp->K, sp->nw, sp->n_edges, sp->n_levels, sp->G, sp->Dmax, sp->n_pieces, sp->nsrc(),
                       sp->zero_entry(), xan::SK_XG, xan::SK_XO, xan::SK_LAGM};
    for (int i = 0; i < 12; ++i) info[i] = v[i];
    return XAN_OK;
}

int xan_mrtm_skew_window(xan_mrtm_plan *pl, int requested, int nt_min, int *out4) {
    XAN_REQUIRE(pl && out4, "xan_mrtm_skew_window: null pointer");
    xan::SkewPlan *sp = xan::get_skew(pl);
    out4[0] = sp ? xan::skew_window(sp, requested, nt_min, 1024) : 0;
    out4[1] = xan::SK_CH;
    out4[2] = 1024;                 // ring entries per cut edge (XANTHOS_MRTM_SKEW_RING default)
    out4[3] = xan::SK_WINDOW_DEFAULT;
    return XAN_OK;
}

int xan_mrtm_skew_tables(xan_mrtm_plan *pl, int *cell, int *lag, int *src, int *ghost_edge, int *ghost_lag,
                         int *exp_edge, int *exp_place, int *Dw, int *edge_prod, int *edge_cons) {
    XAN_REQUIRE(pl, "xan_mrtm_skew_tables: null plan");
    xan::SkewPlan *sp = xan::get_skew(pl);
    XAN_REQUIRE(sp, "xan_mrtm_skew_tables: no skew plan for this flow graph");
    auto cp = [](const std::vector<int> &v, int *dst) {
        if (dst) std::copy(v.begin(), v.end(), dst);
    };
    cp(sp->cell, cell); cp(sp->lag, lag); cp(sp->src, src); cp(sp->ghost_edge, ghost_edge);
    cp(sp->ghost_lag, ghost_lag); cp(sp->exp_edge, exp_edge); cp(sp->exp_place, exp_place); cp(sp->Dw, Dw);
    cp(sp->edge_prod, edge_prod); cp(sp->edge_cons, edge_cons);
    return XAN_OK;
}

}  // extern "C"
