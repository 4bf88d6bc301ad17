// MRTM river routing (xanthos/routing/mrtm.py, xanthos/components.py:262-296), fp64, compiled with
// -fmad=false: the routing results are BIT-IDENTICAL to the reference's scipy-CSR formulation.
//
// Host side  : integer topology (downstream / upstream / UM = UP - I), forest check, partition of
//              every river tree into "pieces" of at most 31 lanes, levels, packing of pieces into warps.
// Device side:
//   * mrtm_warp_kernel - warp-level dataflow, NO block- or grid-wide barrier anywhere in the time
//     loop.  One warp owns one or more pieces (<= 31 cells + ghost lanes, lane 31 stays empty).  Channel
//     storage S lives in a register of the owning lane for the whole run; the flows F of the warp's cells
//     never leave the register file: the reference's sparse "UM.dot(F)" is a gather of <= 9 row terms with
//     64-bit shuffles in ascending column order (um_row).  Water only moves downstream, so a cut edge
//     between two pieces is a one-way dependency: the upstream warp runs ahead and streams the (F, F')
//     series of its outlet cell into a small ring buffer in global memory (L2 resident), month by month;
//     the downstream warp follows one month behind and stages the series into shared memory with cp.async
//     (acquire/release progress counters, back-pressure through the same counters).  The depth of the
//     piece tree only adds a start-up lag of one month per level.
//   * mrtm_sched_kernel - decides, once per plan, which packed warp runs on which SM sub-partition
//     (expensive warps on SMs of their own, the others grouped by loop variant).
//   * mrtm_grid_kernel - general fallback for graphs that are not forests: cooperative launch,
//     two grid-wide syncs per sub-step, state in global memory.
#include "common.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

namespace cg = cooperative_groups;

#include "mrtm_plan.cuh"

namespace xan {

constexpr int RING_DEFAULT = 4;   // months of a cut-edge series kept in flight
constexpr bool AUTO_DEFAULT_SKEW = true;    // which forest kernel XAN_MRTM_AUTO picks (the faster one as measured)

// =============================================================================================
// host: topology
// =============================================================================================
static int grid_positions(const double *coords, int n, int nrow, int ncol, std::vector<int> &grid,
                          std::vector<int> &ilat, std::vector<int> &ilon) {
    grid.assign((size_t)nrow * ncol, 0);
    ilat.resize(n);
    ilon.resize(n);
    for (int i = 0; i < n; ++i) {
        const int id = (int)coords[i * 5 + 0];
        ilon[i] = (int)coords[i * 5 + 3] - 1;
        ilat[i] = (int)coords[i * 5 + 4] - 1;
        XAN_REQUIRE(ilat[i] >= 0 && ilat[i] < nrow && ilon[i] >= 0 && ilon[i] < ncol,
                    "mrtm: cell %d has grid position (%d, %d) outside %d x %d", i, ilat[i], ilon[i], nrow, ncol);
        XAN_REQUIRE(id == i + 1, "mrtm: coords[:,0] must be 1..ncell in order (row %d holds id %d)", i, id);
    }
    for (int i = 0; i < n; ++i) grid[(size_t)ilat[i] * ncol + ilon[i]] = i + 1;
    return XAN_OK;
}

// downstream (mrtm.py:85-120, make_flowdirgrid :233-258)
static int host_downstream(const double *coords, const double *flow_dir, int n, int nrow, int ncol, int64_t *dsid) {
    std::vector<int> grid, ilat, ilon;
    const int rc = grid_positions(coords, n, nrow, ncol, grid, ilat, ilon);
    if (rc != XAN_OK) return rc;
    const int rt = 1 + 2 + 128, lt = 8 + 16 + 32, up = 32 + 64 + 128, dn = 2 + 4 + 8;
    for (int i = 0; i < n; ++i) {
        const double f = flow_dir[i];
        const int fd = (f == -9999.) ? 0 : (int)f;
        int dlat = 0, dlon = 0;
        if (dn & fd) dlat = -1;
        if (up & fd) dlat = 1;
        if (rt & fd) dlon = 1;
        if (lt & fd) dlon = -1;
        int tlat = ilat[i] + dlat, tlon = ilon[i] + dlon;
        if (tlon < 0 || tlon > ncol - 1) tlon = (((tlon + 1) % ncol) + ncol) % ncol;   // :101-102
        if (tlat < 0 || tlat > nrow - 1) {                                             // :104-106
            tlat = ilat[i];
            tlon = ilon[i];
        }
        int d = grid[(size_t)tlat * ncol + tlon];
        if (d == 0 || d == i + 1) d = -1;                                              // :116-118
        dsid[i] = d;
    }
    return XAN_OK;
}

// upstream (mrtm.py:123-191): stable partition of the 8 neighbours, inflowing first
static int host_upstream(const double *coords, const int64_t *dsid, int n, int nrow, int ncol, int64_t *upid) {
    std::vector<int> grid, ilat, ilon;
    const int rc = grid_positions(coords, n, nrow, ncol, grid, ilat, ilon);
    if (rc != XAN_OK) return rc;
    const int rowoff[8] = {-1, -1, -1, 0, 0, 1, 1, 1}, coloff[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
    for (int i = 0; i < n; ++i) {
        int nb[8];
        bool isup[8];
        for (int k = 0; k < 8; ++k) {
            const int r = ilat[i] + rowoff[k], c = ilon[i] + coloff[k];
            nb[k] = (r >= 0 && c >= 0 && r <= nrow - 1 && c <= ncol - 1) ? grid[(size_t)r * ncol + c] : 0;
            isup[k] = (nb[k] != 0) && (dsid[nb[k] - 1] == i + 1);
        }
        int w = 0;
        for (int k = 0; k < 8; ++k)
            if (isup[k]) upid[(size_t)i * 9 + w++] = nb[k];
        upid[(size_t)i * 9 + 8] = w;
        for (int k = 0; k < 8; ++k)
            if (!isup[k]) upid[(size_t)i * 9 + w++] = nb[k];
    }
    return XAN_OK;
}

// rows of UM = UP - I (mrtm.py:194-230), columns ascending like the canonical CSR of scipy
static int build_rows(xan_mrtm_plan *pl, const int64_t *upid) {
    const int n = pl->ncell;
    pl->upid.assign((size_t)n * 9, 0);
    for (int i = 0; i < n; ++i) {
        const int64_t k = upid[(size_t)i * 9 + 8];
        XAN_REQUIRE(k >= 0 && k <= 8, "mrtm: upid[%d, 8] = %lld is not a neighbour count", i, (long long)k);
        for (int s = 0; s < 9; ++s) pl->upid[(size_t)i * 9 + s] = (int)upid[(size_t)i * 9 + s];
        for (int s = 0; s < k; ++s) {
            const int64_t j = upid[(size_t)i * 9 + s];
            XAN_REQUIRE(j >= 1 && j <= n && j != i + 1, "mrtm: upid[%d, %d] = %lld is not a valid upstream id", i, s,
                        (long long)j);
        }
    }
    pl->row_ptr.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) pl->row_ptr[i + 1] = pl->row_ptr[i] + pl->upid[(size_t)i * 9 + 8] + 1;
    pl->col.assign(pl->row_ptr[n], 0);
    pl->sgn.assign(pl->row_ptr[n], 0);
    pl->down.assign(n, -1);
    pl->multi_receiver = false;
    for (int i = 0; i < n; ++i) {
        const int k = pl->upid[(size_t)i * 9 + 8];
        int ent[9];
        for (int s = 0; s < k; ++s) {
            ent[s] = pl->upid[(size_t)i * 9 + s] - 1;
            if (pl->down[ent[s]] >= 0) pl->multi_receiver = true;   // not produced by `upstream`; grid kernel only
            pl->down[ent[s]] = i;
        }
        ent[k] = i;
        std::sort(ent, ent + k + 1);
        for (int s = 0; s <= k; ++s) {
            pl->col[pl->row_ptr[i] + s] = ent[s];
            pl->sgn[pl->row_ptr[i] + s] = (ent[s] == i) ? -1 : 1;
        }
    }
    return XAN_OK;
}

// =============================================================================================
// host: partition of the forest into warp-sized pieces
// =============================================================================================
struct Packing {
    std::vector<int> lane_cell, lane_gedge, lane_oedge, edge_prod, edge_cons, edge_cell;
    std::vector<uint2> lane_src;
    std::vector<unsigned> lane_meta;
    int n_warps = 0, n_edges = 0, n_levels = 0, G = 0;
};

static bool build_packing(xan_mrtm_plan *pl, int lanes, Packing &pk) {
    const int n = pl->ncell;
    // Kahn order, leaves first; a cycle leaves cells unvisited -> not a forest
    std::vector<int> indeg(n), order;
    order.reserve(n);
    for (int i = 0; i < n; ++i) indeg[i] = pl->upid[(size_t)i * 9 + 8];
    for (int i = 0; i < n; ++i)
        if (indeg[i] == 0) order.push_back(i);
    for (size_t h = 0; h < order.size(); ++h) {
        const int r = pl->down[order[h]];
        if (r >= 0 && --indeg[r] == 0) order.push_back(r);
    }
    pl->is_forest = ((int)order.size() == n) && !pl->multi_receiver;
    if (!pl->is_forest) return false;

    {   // components (for reporting)
        std::vector<int> csize(n, 0), root(n);
        for (int h = n - 1; h >= 0; --h) {
            const int v = order[h];
            root[v] = (pl->down[v] < 0) ? v : root[pl->down[v]];
            csize[root[v]]++;
        }
        pl->n_components = 0;
        pl->max_component = 0;
        for (int v = 0; v < n; ++v)
            if (pl->down[v] < 0) {
                pl->n_components++;
                pl->max_component = std::max(pl->max_component, csize[v]);
            }
    }

    // Bottom-up: a piece occupies (cells + ghost lanes) <= `lanes`; every cut child costs its
    // receiver one ghost lane.  While a sub-tree is too big the heaviest child is cut off.
    std::vector<int> rs(n, 0), rg(n, 0);
    std::vector<char> cut(n, 0);
    for (int h = 0; h < n; ++h) {
        const int v = order[h];
        const int k = pl->upid[(size_t)v * 9 + 8];
        int ch[8], sz = 1, gh = 0;
        for (int s = 0; s < k; ++s) {
            ch[s] = pl->upid[(size_t)v * 9 + s] - 1;
            sz += rs[ch[s]];
            gh += rg[ch[s]];
        }
        if (sz + gh > lanes) {
            std::sort(ch, ch + k, [&](int a, int b) {
                const int wa = rs[a] + rg[a], wb = rs[b] + rg[b];
                return wa != wb ? wa > wb : a < b;
            });
            for (int s = 0; s < k && sz + gh > lanes; ++s) {
                if (rs[ch[s]] + rg[ch[s]] <= 1) break;   // cutting a bare leaf gains nothing
                cut[ch[s]] = 1;
                sz -= rs[ch[s]];
                gh -= rg[ch[s]];
                gh += 1;
            }
            if (sz + gh > lanes) return false;
        }
        rs[v] = sz;
        rg[v] = gh;
    }
    // pieces: roots are outlets and cut cells
    std::vector<int> piece(n, -1), piece_lanes, piece_level;
    for (int h = n - 1; h >= 0; --h) {
        const int v = order[h];
        if (pl->down[v] < 0 || cut[v]) {
            piece[v] = (int)piece_lanes.size();
            piece_lanes.push_back(rs[v] + rg[v]);
        } else {
            piece[v] = piece[pl->down[v]];
        }
    }
    const int np = (int)piece_lanes.size();
    piece_level.assign(np, 0);
    std::vector<char> linked(np, 0);
    for (int h = 0; h < n; ++h) {   // leaves first: a piece's incoming edges are final before its own
        const int v = order[h];
        if (cut[v]) {
            const int pu = piece[v], pd = piece[pl->down[v]];
            piece_level[pd] = std::max(piece_level[pd], piece_level[pu] + 1);
            linked[pu] = linked[pd] = 1;
        }
    }
    int n_levels = 1;
    for (int p = 0; p < np; ++p) n_levels = std::max(n_levels, piece_level[p] + 1);

    // First-fit decreasing into warps.  Linked pieces only share a warp with pieces of the same
    // level, which keeps the warp dependency graph acyclic; free pieces (whole small trees) fill gaps.
    std::vector<int> ids(np);
    std::iota(ids.begin(), ids.end(), 0);
    std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return piece_lanes[a] > piece_lanes[b]; });
    std::vector<int> w_fill, w_level, piece_warp(np, -1);
    {
        // Linked pieces share a warp only when they feed the SAME downstream piece: a warp then has a
        // single consumer, so back-pressure from a slow consumer can never stall data that another
        // (faster) consumer is waiting for.  (Mixed consumers create a cycle through the ring
        // back-pressure that throttles whole river systems to ring_depth / tree_depth.)
        std::vector<int> piece_down(np, -1);
        for (int v = 0; v < n; ++v)
            if (cut[v]) piece_down[piece[v]] = piece[pl->down[v]];
        std::vector<int> last_open(np + 1, -1);   // per downstream piece (index np = "no consumer"): last warp opened
        for (int p : ids) {
            if (!linked[p]) continue;
            const int key = piece_down[p] < 0 ? np : piece_down[p];
            int w = last_open[key];
            if (piece_down[p] < 0 || w < 0 || w_fill[w] + piece_lanes[p] > lanes) {
                w = (int)w_fill.size();
                w_fill.push_back(0);
                w_level.push_back(piece_level[p]);
                if (piece_down[p] >= 0) last_open[key] = w;
            }
            w_fill[w] += piece_lanes[p];
            w_level[w] = std::max(w_level[w], piece_level[p]);   // siblings may sit on different levels
            piece_warp[p] = w;
        }
    }
    {
        // free pieces: bucket the open warps by remaining room for an O(1)-ish first fit
        std::vector<std::vector<int>> by_room(lanes + 1);
        for (size_t w = 0; w < w_fill.size(); ++w) by_room[lanes - w_fill[w]].push_back((int)w);
        for (int p : ids) {
            if (linked[p]) continue;
            int w = -1;
            for (int room = piece_lanes[p]; room <= lanes && w < 0; ++room)
                if (!by_room[room].empty()) {
                    w = by_room[room].back();
                    by_room[room].pop_back();
                }
            if (w < 0) {
                w = (int)w_fill.size();
                w_fill.push_back(0);
                w_level.push_back(-1);
            }
            w_fill[w] += piece_lanes[p];
            piece_warp[p] = w;
            by_room[lanes - w_fill[w]].push_back(w);
        }
    }
    // order warps by level so that producers get the lower indices; free warps go last
    const int nw = (int)w_fill.size();
    std::vector<int> word(nw), wnew(nw);
    std::iota(word.begin(), word.end(), 0);
    std::stable_sort(word.begin(), word.end(), [&](int a, int b) {
        const int la = w_level[a] < 0 ? n_levels : w_level[a], lb = w_level[b] < 0 ? n_levels : w_level[b];
        return la < lb;
    });
    for (int q = 0; q < nw; ++q) wnew[word[q]] = q;

    // lanes: cells first, ghost lanes behind them.  Cells take their lanes in depth-first post-order
    // (every cell right after the sub-tree of its last tributary): along a river reach the upstream
    // neighbour then sits in the previous lane, so the 64-bit gathers of a half-warp hit 16 different
    // bank pairs instead of random ones.
    std::vector<int> post;
    post.reserve(n);
    {
        std::vector<std::pair<int, int>> stack;   // (cell, next child slot)
        for (int r = 0; r < n; ++r) {
            if (pl->down[r] >= 0) continue;
            stack.emplace_back(r, 0);
            while (!stack.empty()) {
                auto &top = stack.back();
                const int v = top.first;
                if (top.second < pl->upid[(size_t)v * 9 + 8]) {
                    const int c = pl->upid[(size_t)v * 9 + top.second] - 1;
                    ++top.second;
                    stack.emplace_back(c, 0);
                } else {
                    post.push_back(v);
                    stack.pop_back();
                }
            }
        }
    }
    std::vector<int> cell_warp(n), cell_lane(n), next_lane(nw, 0);
    for (int v : post) {
        const int w = wnew[piece_warp[piece[v]]];
        cell_warp[v] = w;
        cell_lane[v] = next_lane[w]++;
    }
    pk.n_warps = nw;
    pk.n_levels = n_levels;
    pk.lane_cell.assign((size_t)nw * 32, -1);
    pk.lane_gedge.assign((size_t)nw * 32, -1);
    pk.lane_oedge.assign((size_t)nw * 32, -1);
    pk.lane_src.assign((size_t)nw * 32, make_uint2(0, 0));
    pk.lane_meta.assign((size_t)nw * 32, 0);
    std::vector<int> ghost_lane(n, -1), n_ghost(nw, 0);
    for (int v = 0; v < n; ++v) {
        const int r = pl->down[v];
        if (r >= 0 && cell_warp[r] != cell_warp[v]) {
            const int wp = cell_warp[v], wc = cell_warp[r];
            if (wp > wc) return false;   // producers must precede consumers
            const int e = (int)pk.edge_prod.size();
            pk.edge_prod.push_back(wp);
            pk.edge_cons.push_back(wc);
            pk.edge_cell.push_back(v);
            if (next_lane[wc] >= 32) return false;
            const int gl = next_lane[wc]++;
            ghost_lane[v] = gl;
            pk.lane_gedge[(size_t)wc * 32 + gl] = e;
            pk.lane_meta[(size_t)wc * 32 + gl] = (unsigned)n_ghost[wc]++ << 8;
            pk.lane_oedge[(size_t)wp * 32 + cell_lane[v]] = e;
        }
    }
    pk.n_edges = (int)pk.edge_prod.size();
    for (int w = 0; w < nw; ++w) pk.G = std::max(pk.G, n_ghost[w]);
    for (int v = 0; v < n; ++v) {
        const int w = cell_warp[v];
        const size_t g = (size_t)w * 32 + cell_lane[v];
        pk.lane_cell[g] = v;
        // row of UM in column order, 7 bits per term: 0..31 = +F of that lane, 32 = zero padding,
        // 33 + l = -F of lane l (the cell's own term)
        unsigned long long bits = 0;
        const int beg = pl->row_ptr[v], cnt = pl->row_ptr[v + 1] - beg;
        for (int s = 0; s < 9; ++s) {
            unsigned idx = 32;
            if (s < cnt) {
                const int j = pl->col[beg + s];
                idx = (j == v) ? 33u + (unsigned)cell_lane[v]
                               : (unsigned)((cell_warp[j] == w) ? cell_lane[j] : ghost_lane[j]);
            }
            bits |= (unsigned long long)idx << (7 * s);
        }
        pk.lane_src[g] = make_uint2((unsigned)(bits & 0xffffffffu), (unsigned)(bits >> 32));
        pk.lane_meta[g] = (unsigned)cnt;
    }
    // padding rows (ghost and empty lanes): all nine terms read the zero slot
    {
        unsigned long long zero_row = 0;
        for (int s = 0; s < 9; ++s) zero_row |= 32ull << (7 * s);
        for (size_t g = 0; g < pk.lane_cell.size(); ++g)
            if (pk.lane_cell[g] < 0) pk.lane_src[g] = make_uint2((unsigned)(zero_row & 0xffffffffu), (unsigned)(zero_row >> 32));
    }
    return true;
}

// =============================================================================================
// device helpers
// =============================================================================================
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// =============================================================================================
// warp kernel
// =============================================================================================
constexpr int NM_MAX = 2;      // ensemble members advanced together by one warp (independent chains = ILP)

struct WarpArgs {
    const int *lane_cell, *lane_gedge, *lane_oedge;
    const uint2 *lane_src;
    const unsigned *lane_meta;
    const int *edge_prod, *edge_cons;
    int *progress;
    double2 *ring_buf;                 // [n_edges][ring][ntmax][NM] (F, F')
    const double *runoff[NM_MAX];      // per member [M][ld]
    const double *chs_prev[NM_MAX];    // per member [ncell] or null
    const double *flow_dist, *velocity, *area;
    const int *ndays;                  // [M] device
    double *chs[NM_MAX], *avg[NM_MAX], *instream[NM_MAX];
    int n_warps, G, ntmax, nmonths, spinup, ld, ring, sleep_ns, sb;
    double dt;
    const int *sched;        // [grid warps] packed warp run by each grid warp (mrtm_sched_kernel); null = identity
    long long *dbg;          // optional [n_warps][6]: cycles total, hand-over wait, staging wait, sub-step loops, redo count, SM id
};

template <int NM>
struct LaneState {
    double S[NM], erl[NM], Favg[NM], F[NM];
    double tauinv;
    uint2 src;
    bool is_cell;
};

// ---------------------------------------------------------------------------------------------
// Row i of UM times F in ascending column order (scipy csr_matvec order) without leaving the
// register file.  A row is NT source lanes; the cell's own term reads its own lane and flips the
// sign bit (an integer XOR on the high word, exact), padding terms read lane 31, which the packing
// keeps empty (F == 0.0 for the whole run; x + 0.0 == x for every x that is not -0.0).  Per sub-step
// and member: 2 NT SHFL.IDX, NT LOP, NT + 5 fp64 instructions, one vote - no shared-memory traffic,
// no bank conflicts, no __syncwarp.
// (Splitting the row at the own term into "before" and "after" lists saves the XOR and some rounds,
// but needs 25 x 2 loop instantiations instead of 8 x 4; with 16 warps of an SM in different loops the
// 32 KB instruction cache thrashes - measured 78 ms against 53 ms.)
// ---------------------------------------------------------------------------------------------
template <int NT>
struct RowSrc {
    int lane[NT];
    unsigned sgn[NT];
};

template <int NT>
__device__ __forceinline__ RowSrc<NT> decode_row(uint2 src) {
    const unsigned long long bits = ((unsigned long long)src.y << 32) | src.x;
    RowSrc<NT> r;
#pragma unroll
    for (int s = 0; s < NT; ++s) {
        const int idx = (int)((bits >> (7 * s)) & 0x7f);   // 0..31 lane, 32 zero, 33 + l = -F of lane l
        r.lane[s] = (idx < 32) ? idx : ((idx == 32) ? 31 : idx - 33);
        r.sgn[s] = (idx > 32) ? 0x80000000u : 0u;
    }
    return r;
}

template <int NT>
__device__ __forceinline__ double um_row(double F, const RowSrc<NT> &r) {
    const unsigned full = 0xffffffffu;
    const int lo = __double2loint(F), hi = __double2hiint(F);
    double v[NT];
#pragma unroll
    for (int s = 0; s < NT; ++s) {   // all shuffles in flight together
        const int l = __shfl_sync(full, lo, r.lane[s]);
        const int h = __shfl_sync(full, hi, r.lane[s]) ^ (int)r.sgn[s];
        v[s] = __hiloint2double(h, l);
    }
    double d = v[0];
#pragma unroll
    for (int s = 1; s < NT; ++s) d = d + v[s];
    return d;
}

// Per-lane predicated memory operations as opaque PTX.  Written as C++ `if (is_ghost) ...` the compiler
// unswitches the sub-step loop on the (loop-invariant, per-lane) condition; ghost and cell lanes then
// run in different copies of the loop and every shuffle becomes a WARPSYNC.COLLECTIVE rendezvous of
// a diverged warp.  With predication the warp stays converged.
__device__ __forceinline__ void lds_v2_pred(double &x, double &y, const double2 *p, int pred) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\t@p ld.shared.v2.f64 {%0, %1}, [%2];\n\t}"
                 : "+d"(x), "+d"(y)
                 : "r"(addr), "r"(pred));
}
__device__ __forceinline__ void stg_v2_pred(double2 *p, double x, double y, int pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\t@p st.global.v2.f64 [%0], {%1, %2};\n\t}" ::"l"(p),
                 "d"(x), "d"(y), "r"(pred)
                 : "memory");
}
__device__ __forceinline__ double bit_select(double a, double b, long long mask) {   // mask ? a : b, bitwise
    return __longlong_as_double((__double_as_longlong(a) & mask) | (__double_as_longlong(b) & ~mask));
}
// relaxed load if pred != 0, else `dflt`
__device__ __forceinline__ int ld_relaxed_pred(const int *p, int pred, int dflt) {
    int v = dflt;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.relaxed.gpu.global.s32 %0, [%1];\n\t}"
                 : "+r"(v)
                 : "l"(p), "r"(pred)
                 : "memory");
    return v;
}

// `len` (<= sb) sub-steps for one warp and NM members.  gsl: this lane's staged ghost series [len + 1][NM]
// of (F, F') (only read by ghost lanes; entry `len` is read but never used); out: ring position [len][NM]
// written by lanes that feed a cut edge.  GHOST: the warp has ghost lanes; OUT: it feeds a cut edge.
// Lanes that hold no cell (ghost and empty lanes) have S == 0, tauinv == 0, erl == 0 and an all-zero row,
// so they never clamp and need no `is_cell` test.
//
// The loop is rotated by one gather: an iteration starts with d = UM.F + erlateral of sub-step t already
// summed, and ends by gathering and summing the trial flows of sub-step t + 1.  The vote "did any flow of
// this warp change" (a clamp in the warp or a ghost series with F' != F) is issued before that gather and
// only consumed after it, so in the common case neither the vote nor its branch is on the loop-carried
// chain  d -> DMUL -> DADD -> DMUL -> SHFL -> LOP -> NT x DADD -> DADD.  When a flow changed the balance is
// repeated with F' (mrtm.py:56-69) and the look-ahead is redone from the corrected storage.
// The members are independent dependency chains (ILP).
template <int NT, int NM, bool GHOST, bool OUT>
__device__ __forceinline__ void run_block(LaneState<NM> &L, const double2 *gsl, bool is_ghost, double2 *out,
                                          bool lane_out, int len, double dt, double dtinv, double (&lastFp)[NM],
                                          int &n_redo) {
    const unsigned full = 0xffffffffu;
    const RowSrc<NT> row = decode_row<NT>(L.src);
    const int out_i = lane_out ? 1 : 0;
    const long long gmask = -(long long)(is_ghost ? 1 : 0);
    // Every lane reads a staged series without predication: ghost lanes their cut edge, all other lanes the
    // warp's all-zero slot (one broadcast wavefront).  (F, F') == (0, 0) never reports a change, so the values
    // need no per-lane guard and are plain SSA values from one iteration to the next (no in-out registers).
    double gFp[NM], d[NM];
    bool gchg[NM];
#pragma unroll
    for (int mm = 0; mm < NM; ++mm) {
        gFp[mm] = 0.0;
        gchg[mm] = false;
        if (GHOST) {
            const double2 g0 = gsl[mm];
            L.F[mm] = bit_select(g0.x, L.F[mm], gmask);
            gFp[mm] = g0.y;
            gchg[mm] = __double_as_longlong(g0.x) != __double_as_longlong(g0.y);
        }
        d[mm] = um_row<NT>(L.F[mm], row) + L.erl[mm];                            // mrtm.py:51
    }
    const double2 *gnext = gsl + NM;
    double2 *op = out;
#pragma unroll 2
    for (int t = 0; t < len; ++t) {
        double nF[NM], nFp[NM], Fp[NM], Sn[NM], Fn[NM], dn[NM];
        bool clamp[NM];
        bool changed = false;
#pragma unroll
        for (int mm = 0; mm < NM; ++mm) {
            nF[mm] = 0.0;
            nFp[mm] = 0.0;
            if (GHOST) {                                                         // ghost flows of sub-step t + 1
                const double2 g1 = gnext[mm];
                nF[mm] = g1.x;
                nFp[mm] = g1.y;
            }
        }
#pragma unroll
        for (int mm = 0; mm < NM; ++mm) {
            const double ddt = d[mm] * dt;
            clamp[mm] = ddt < (-L.S[mm]);                                        // mrtm.py:54
            changed = changed || clamp[mm];
            if (GHOST) changed = changed || gchg[mm];
            Sn[mm] = L.S[mm] + ddt;                                              // mrtm.py:76
            Fn[mm] = Sn[mm] * L.tauinv;                                          // mrtm.py:50 of sub-step t + 1
            if (GHOST) Fn[mm] = bit_select(nF[mm], Fn[mm], gmask);
        }
        const bool redo = __any_sync(full, changed);
#pragma unroll
        for (int mm = 0; mm < NM; ++mm) {                                        // look-ahead, assumes !redo
            dn[mm] = um_row<NT>(Fn[mm], row) + L.erl[mm];
            Fp[mm] = L.F[mm];
        }
        if (redo) {
            ++n_redo;
#pragma unroll
            for (int mm = 0; mm < NM; ++mm) {
                const double Fc = d[mm] + L.F[mm] + L.S[mm] * dtinv;             // mrtm.py:60
                Fp[mm] = clamp[mm] ? Fc : L.F[mm];
                if (GHOST) Fp[mm] = bit_select(gFp[mm], Fp[mm], gmask);
                const double S2 = L.S[mm] + (um_row<NT>(Fp[mm], row) + L.erl[mm]) * dt;   // mrtm.py:68-69
                Sn[mm] = clamp[mm] ? 0.0 : S2;                                   // mrtm.py:63
                Fn[mm] = Sn[mm] * L.tauinv;
                if (GHOST) Fn[mm] = bit_select(nF[mm], Fn[mm], gmask);
                dn[mm] = um_row<NT>(Fn[mm], row) + L.erl[mm];
            }
        }
#pragma unroll
        for (int mm = 0; mm < NM; ++mm) {
            if (OUT) stg_v2_pred(op + mm, L.F[mm], Fp[mm], out_i);
            L.S[mm] = Sn[mm];
            L.Favg[mm] += Fp[mm];                                                // mrtm.py:78
            lastFp[mm] = Fp[mm];
            L.F[mm] = Fn[mm];
            d[mm] = dn[mm];
            if (GHOST) {
                gFp[mm] = nFp[mm];
                gchg[mm] = __double_as_longlong(nF[mm]) != __double_as_longlong(nFp[mm]);
            }
        }
        gnext += NM;
        op += NM;
    }
}

template <int NT, int NM>
__device__ __forceinline__ void run_block_nt(bool has_ghost, bool has_out, LaneState<NM> &L, const double2 *gsl,
                                             bool is_ghost, double2 *out, bool lane_out, int len, double dt,
                                             double dtinv, double (&lastFp)[NM], int &n_redo) {
    if (has_ghost && has_out)
        run_block<NT, NM, true, true>(L, gsl, is_ghost, out, lane_out, len, dt, dtinv, lastFp, n_redo);
    else if (has_ghost)
        run_block<NT, NM, true, false>(L, gsl, is_ghost, out, lane_out, len, dt, dtinv, lastFp, n_redo);
    else if (has_out)
        run_block<NT, NM, false, true>(L, gsl, is_ghost, out, lane_out, len, dt, dtinv, lastFp, n_redo);
    else
        run_block<NT, NM, false, false>(L, gsl, is_ghost, out, lane_out, len, dt, dtinv, lastFp, n_redo);
}


// ---------------------------------------------------------------------------------------------
// Placement-aware work assignment.  The run ends when the slowest chain of packed warps ends, and the
// slow ones are known before the launch: a warp that holds a cell with dt V / L > 1 (or a ghost lane fed
// by such a cell) repeats the balance with F' at every other sub-step (that cell empties, refills,
// empties ...) and needs ~1.7x the instructions of the others.  With packed warp w bound to grid warp
// w, up to 7 of them met on one SM and 3 in one SM sub-partition, where they share one issue port.
// A block is 16 warps = one SM (1 block per SM by registers), warp i of a block runs in sub-partition
// i & 3, so the table assign[grid warp] -> packed warp decides who shares an issue port:
//   * the expensive warps go two per sub-partition and sub-partition by sub-partition onto whole SMs (8 per SM,
//     ~40 SMs here) with no other warp beside them: what slows them is the SM-wide shuffle (LSU) pipe and the issue
//     ports they would share with 15-18 cheap warps (XANTHOS_MRTM_EXP_PER_SP / XANTHOS_MRTM_PACK_SM change this);
//   * the others are sorted by loop variant (row length NT, ghost lanes, cut-edge output) and fill the remaining
//     sub-partitions in runs of consecutive entries, so that the warps of a sub-partition execute the same copy of
//     the sub-step loop (6 KB L0 instruction cache); the spare slots of the grid stay beside the expensive warps.
// The arithmetic of a packed warp does not depend on where it runs: results are bit-identical.
// One block, thread = packed warp; ranks by counting (n_warps <= 148 x 16 because all warps are co-resident).
// ---------------------------------------------------------------------------------------------
__global__ void mrtm_sched_kernel(const int *lane_cell, const int *lane_gedge, const int *lane_oedge,
                                  const unsigned *lane_meta, const int *edge_cell, const double *flow_dist,
                                  const double *velocity, double dt, int n_warps, int n_blocks, int wpb, int group,
                                  int exp_per_sp, int pack_sm, int *assign) {
    extern __shared__ int s_key[];   // [n_warps]
    for (int i = threadIdx.x; i < n_blocks * wpb; i += blockDim.x) assign[i] = -1;
    for (int w = threadIdx.x; w < n_warps; w += blockDim.x) {
        bool expensive = false, ghost = false, out = false;
        int nt = 1;
        for (int l = 0; l < 32; ++l) {
            const size_t g = (size_t)w * 32 + l;
            int c = lane_cell[g];
            const int e = lane_gedge[g];
            if (c >= 0) nt = max(nt, (int)(lane_meta[g] & 0xf));
            if (e >= 0) {
                c = edge_cell[e];
                ghost = true;
            }
            out = out || lane_oedge[g] >= 0;
            if (c >= 0) expensive = expensive || (velocity[c] / flow_dist[c]) * dt > 1.0;
        }
        const int variant = (group & 1) ? (nt * 4 + (ghost ? 2 : 0) + (out ? 1 : 0)) : 0;
        s_key[w] = (expensive ? 0 : 1 << 16) | variant;
    }
    __syncthreads();
    const int S = wpb / 4, B = n_blocks;
    for (int w = threadIdx.x; w < n_warps; w += blockDim.x) {
        const int key = s_key[w];
        int r = 0, nE = 0;
        for (int v = 0; v < n_warps; ++v) {
            const int kv = s_key[v];
            r += (kv < key || (kv == key && v < w)) ? 1 : 0;
            nE += (kv < (1 << 16)) ? 1 : 0;
        }
        // EP expensive warps share a sub-partition (slots 0 .. EP - 1); they get as few companions as the grid allows
        // (the spare slots of the grid stay beside them); sub-partition k is (block k % B, quarter k / B): spread over SMs
        const int EP = max(1, min(exp_per_sp, S));
        const int nEx = min(nE, EP * 4 * B);              // expensive warps placed as such (the rest counts as cheap)
        const int nEsp = (nEx + EP - 1) / EP;             // sub-partitions they occupy
        // slot sequence: A = the EP expensive slots of the first nEsp sub-partitions, B = c companion slots beside
        // them, C = the other sub-partitions; the warp of rank r (expensive first, then by loop variant) takes slot r
        const int nA = nEsp * EP;
        int c = S - EP;
        while (c > 0 && nA + nEsp * (c - 1) + (4 * B - nEsp) * S >= n_warps) --c;
        int k, pos;
        if (r < nA) {
            k = r / EP; pos = r % EP;
        } else {
            const int j = r - nA, cap1 = c * nEsp;
            if (j < cap1) {
                k = j / c; pos = EP + j % c;
            } else {
                const int jj = j - cap1;
                k = nEsp + jj / S; pos = jj % S;
            }
        }
        // sub-partition k -> (block, quarter): spread over the SMs (quarter-major), or SM by SM (pack_sm)
        const int q = pack_sm ? k % 4 : k / B, blk = pack_sm ? k / 4 : k % B;
        assign[blk * wpb + q + 4 * pos] = w;
    }
}

template <int NM, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) mrtm_warp_kernel(const WarpArgs a) {
    extern __shared__ double smem[];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    int w = blockIdx.x * wpb + wib;
    if (a.sched) w = a.sched[w];
    if (w < 0 || w >= a.n_warps) return;   // no block-level barrier is used below
    const int SB = a.sb;                                               // sub-steps of ghost series staged at a time
    const int SBP = SB + 1;                                            // + 1: the look-ahead reads one entry past a block
    const int GS = a.G + 1;                                            // + 1: an all-zero series for the lanes without ghost
    const int per_warp = 2 * GS * SBP * NM * 2;                        // doubles of shared memory per warp
    double2 *gs = reinterpret_cast<double2 *>(smem + (size_t)wib * per_warp);   // [2][G + 1][SB + 1][NM]
    for (int i = lane; i < 2 * GS * SBP * NM; i += 32) gs[i] = make_double2(0.0, 0.0);
    __syncwarp();

    const size_t g = (size_t)w * 32 + lane;
    const int cell = a.lane_cell[g], gedge = a.lane_gedge[g], oedge = a.lane_oedge[g];
    const unsigned meta = a.lane_meta[g];
    LaneState<NM> L;
    L.src = a.lane_src[g];
    L.is_cell = cell >= 0;
    L.tauinv = 0.0;
    double area = 0.0, qn[NM], lastFp[NM];
    if (L.is_cell) {
        L.tauinv = a.velocity[cell] / a.flow_dist[cell];                            // mrtm.py:42
        area = a.area[cell];
    }
#pragma unroll
    for (int mm = 0; mm < NM; ++mm) {
        L.S[mm] = 0.0; L.erl[mm] = 0.0; L.Favg[mm] = 0.0; L.F[mm] = 0.0;
        qn[mm] = 0.0;
        lastFp[mm] = 0.0;
        if (L.is_cell) {
            L.S[mm] = a.chs_prev[mm] ? a.chs_prev[mm][cell] : 0.0;
            qn[mm] = a.runoff[mm][cell];   // month 0 of the first pass
        }
    }
    const int gslot = (meta >> 8) & 0xff;
    const int ntmax_row = __reduce_max_sync(full, L.is_cell ? (int)(meta & 0xf) : 1);   // longest row of the warp
    const bool is_ghost = gedge >= 0, lane_out = oedge >= 0;
    const int prod = (gedge >= 0) ? a.edge_prod[gedge] : -1;
    const int cons = (oedge >= 0) ? a.edge_cons[oedge] : -1;
    const unsigned ghost_mask = __ballot_sync(full, gedge >= 0);
    const bool has_out = __any_sync(full, oedge >= 0);
    const bool linked = has_out || ghost_mask != 0;
    const double dt = a.dt, dtinv = 1. / a.dt;                                      // mrtm.py:43
    const int nsteps = a.spinup + a.nmonths;
    const bool dbg = a.dbg != nullptr;   // optional per-warp cycle accounting (XANTHOS_MRTM_DEBUG=<file>)
    long long cyc_wait = 0, cyc_stage = 0, cyc_loop = 0;
    int n_redo = 0;
    const long long cyc_all0 = dbg ? clock64() : 0;

    for (int step = 0; step < nsteps; ++step) {
        const bool store = step >= a.spinup;
        const int m = store ? step - a.spinup : step;
        const int nday = a.ndays[m];
        const int nt = (int)((double)nday * 24 * 3600 / dt);                        // mrtm.py:36
        const double secs = (double)(nday * 24 * 3600);
        const int mnext = (step + 1 < nsteps) ? ((step + 1 >= a.spinup) ? step + 1 - a.spinup : step + 1) : m;
        const int slot = step % a.ring;
#pragma unroll
        for (int mm = 0; mm < NM; ++mm) {
            L.erl[mm] = (qn[mm] * area) * (1e6 / 1e3) / secs;                       // mrtm.py:45
            L.Favg[mm] = 0.0;
            L.F[mm] = L.S[mm] * L.tauinv;                                           // mrtm.py:50
            if (L.is_cell) qn[mm] = a.runoff[mm][(size_t)mnext * a.ld + cell];      // prefetch next month
        }
        if (linked) {
            const long long c0 = dbg ? clock64() : 0;
            // Producers have published this month; consumers have released the ring slot.
            // Polled by the whole warp in lock step (predicated loads, one vote per round).  A per-lane
            // `while` here leaves the polling lane diverged from the other 31 for the rest of the month:
            // the sub-step loop then runs twice per warp and every shuffle takes the divergent path
            // (measured: 4.5x slower).
            const int want_p = step + 1, want_c = step - a.ring + 1;
            const int poll_p = (prod >= 0) ? 1 : 0, poll_c = (cons >= 0 && step >= a.ring) ? 1 : 0;
            const int *pp = a.progress + (prod >= 0 ? prod : 0), *pc = a.progress + (cons >= 0 ? cons : 0);
            for (;;) {
                const int vp = ld_relaxed_pred(pp, poll_p, want_p), vc = ld_relaxed_pred(pc, poll_c, want_c);
                if (__all_sync(full, vp >= want_p && vc >= want_c)) break;
                __nanosleep(a.sleep_ns);
            }
            __threadfence();   // acquire side of the hand-over
            __syncwarp();
            if (dbg) cyc_wait += clock64() - c0;
        }
        const double2 *gsrc = (gedge >= 0) ? a.ring_buf + ((size_t)gedge * a.ring + slot) * a.ntmax * NM : nullptr;
        double2 *obase = a.ring_buf + ((size_t)(lane_out ? oedge : 0) * a.ring + slot) * a.ntmax * NM;

        // Ghost series are staged ring (L2) -> shared memory with cp.async, one block of SB sub-steps
        // ahead of the computation (double buffer), so the L2 latency never sits on the critical path.
        auto stage = [&](int t0, int buf) {
            unsigned gm = ghost_mask;
            const int cnt = min(SB, nt - t0) * NM;
            while (gm) {
                const int src_lane = __ffs(gm) - 1;
                gm &= gm - 1;
                const double2 *src = reinterpret_cast<const double2 *>(
                    __shfl_sync(full, (unsigned long long)gsrc, src_lane));
                const int sl = __shfl_sync(full, gslot, src_lane);
                for (int i = lane; i < cnt; i += 32) {
                    const unsigned dst =
                        (unsigned)__cvta_generic_to_shared(gs + ((size_t)buf * GS + sl) * SBP * NM + i);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (size_t)t0 * NM + i)
                                 : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (ghost_mask) stage(0, 0);
        int buf = 0;
        for (int t0 = 0; t0 < nt; t0 += SB, buf ^= 1) {
            const int len = min(SB, nt - t0);
            const long long cs0 = dbg ? clock64() : 0;
            if (ghost_mask) {
                if (t0 + SB < nt) {
                    stage(t0 + SB, buf ^ 1);
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                } else {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
                __syncwarp();
            }
            const long long cs1 = dbg ? clock64() : 0;
            cyc_stage += cs1 - cs0;
            const double2 *gcur = gs + ((size_t)buf * GS + (is_ghost ? gslot : a.G)) * SBP * NM;
            double2 *out = obase + (size_t)t0 * NM;
#define XAN_RUN(NT_) \
    run_block_nt<NT_, NM>(ghost_mask != 0, has_out, L, gcur, is_ghost, out, lane_out, len, dt, dtinv, lastFp, n_redo)
            switch (ntmax_row) {
                case 1: XAN_RUN(1); break;
                case 2: XAN_RUN(2); break;
                case 3: XAN_RUN(3); break;
                case 4: XAN_RUN(4); break;
                case 5: XAN_RUN(5); break;
                case 6: XAN_RUN(6); break;
                case 7: XAN_RUN(7); break;
                default: XAN_RUN(9); break;
            }
#undef XAN_RUN
            if (dbg) cyc_loop += clock64() - cs1;
            if (ghost_mask) __syncwarp();   // everyone is done with this buffer before it is refilled
        }
        if (store && L.is_cell) {
#pragma unroll
            for (int mm = 0; mm < NM; ++mm) {
                if (a.chs[mm]) stg_stream(a.chs[mm] + (size_t)m * a.ld + cell, L.S[mm]);
                if (a.avg[mm]) stg_stream(a.avg[mm] + (size_t)m * a.ld + cell, L.Favg[mm] / nt);   // mrtm.py:80
            }
        }
        if (linked) {
            // publish: this warp has finished (and, as a consumer, has finished reading) month `step`
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release(a.progress + w, step + 1);
        }
    }
    if (L.is_cell) {
#pragma unroll
        for (int mm = 0; mm < NM; ++mm)
            if (a.instream[mm]) a.instream[mm][cell] = lastFp[mm];
    }
    if (dbg && lane == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        a.dbg[6 * w] = clock64() - cyc_all0;
        a.dbg[6 * w + 1] = cyc_wait;
        a.dbg[6 * w + 2] = cyc_stage;
        a.dbg[6 * w + 3] = cyc_loop;
        a.dbg[6 * w + 4] = n_redo;
        unsigned wid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        a.dbg[6 * w + 5] = smid * 4 + (wid & 3);   // SM sub-partition
    }
}

// =============================================================================================
// grid kernel (fallback, any graph)
// =============================================================================================
struct GridArgs {
    const int *gcol;   // [9][ncell]
    const double *runoff, *flow_dist, *velocity, *area, *chs_prev;
    const int *ndays;
    double *S, *Favg, *D, *X, *Y, *Z;   // [ncell] work arrays (X and Z swap every sub-step)
    double *chs, *avg, *instream;
    int ncell, nmonths, spinup, ld;
    double dt;
};

__device__ __forceinline__ double gather_global(const int *__restrict__ gcol, const double *F, int ncell, int c) {
    double d = 0.0;
    for (int s = 0; s < 9; ++s) {
        const int e = gcol[(size_t)s * ncell + c];
        if (e == -1) break;
        const double v = __ldcg(F + (e & 0x7fffffff));
        d = d + ((e < 0) ? -v : v);
    }
    return d;
}

// Two grid-wide syncs per sub-step: (A) trial balance + clamp -> F', (B) balance with F' -> S, next F.
__global__ void __launch_bounds__(256) mrtm_grid_kernel(const GridArgs a) {
    cg::grid_group grid = cg::this_grid();
    const int n = a.ncell, stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const double dt = a.dt, dtinv = 1. / a.dt;
    double *X = a.X, *Z = a.Z;
    for (int c = t0; c < n; c += stride) a.S[c] = a.chs_prev ? a.chs_prev[c] : 0.0;
    const int nsteps = a.spinup + a.nmonths;
    for (int step = 0; step < nsteps; ++step) {
        const bool store = step >= a.spinup;
        const int m = store ? step - a.spinup : step;
        const int nday = a.ndays[m];
        const int nt = (int)((double)nday * 24 * 3600 / dt);
        const double secs = (double)(nday * 24 * 3600);
        for (int c = t0; c < n; c += stride) {
            a.Favg[c] = 0.0;
            X[c] = a.S[c] * (a.velocity[c] / a.flow_dist[c]);
        }
        grid.sync();
        for (int t = 0; t < nt; ++t) {
            for (int c = t0; c < n; c += stride) {
                const double S = a.S[c], F = __ldcg(X + c);
                const double erl = (a.runoff[(size_t)m * a.ld + c] * a.area[c]) * (1e6 / 1e3) / secs;
                const double d = gather_global(a.gcol, X, n, c) + erl;
                const bool clamp = (d * dt) < (-S);
                a.D[c] = d;
                a.Y[c] = clamp ? (d + F + S * dtinv) : F;
            }
            grid.sync();
            for (int c = t0; c < n; c += stride) {
                const double S = a.S[c], Fp = __ldcg(a.Y + c);
                const double erl = (a.runoff[(size_t)m * a.ld + c] * a.area[c]) * (1e6 / 1e3) / secs;
                const bool clamp = (a.D[c] * dt) < (-S);
                double Sn = 0.0;
                if (!clamp) Sn = S + (gather_global(a.gcol, a.Y, n, c) + erl) * dt;
                a.S[c] = Sn;
                a.Favg[c] += Fp;
                Z[c] = Sn * (a.velocity[c] / a.flow_dist[c]);
                if (t == nt - 1 && step == nsteps - 1 && a.instream) a.instream[c] = Fp;
            }
            grid.sync();
            double *tmp = X; X = Z; Z = tmp;
        }
        if (store)
            for (int c = t0; c < n; c += stride) {
                if (a.chs) a.chs[(size_t)m * a.ld + c] = a.S[c];
                if (a.avg) a.avg[(size_t)m * a.ld + c] = a.Favg[c] / nt;
            }
    }
}

}  // namespace xan

using namespace xan;

static void free_device(xan_mrtm_plan *pl) {
    cudaFree(pl->d_gcol);
    cudaFree(pl->d_lane_cell);
    cudaFree(pl->d_lane_gedge);
    cudaFree(pl->d_lane_oedge);
    cudaFree(pl->d_lane_src);
    cudaFree(pl->d_lane_meta);
    cudaFree(pl->d_edge_prod);
    cudaFree(pl->d_edge_cons);
    cudaFree(pl->d_progress);
    cudaFree(pl->d_edge_cell);
    if (pl->d_sched) cudaFree(pl->d_sched);
}

template <typename V>
static bool upload(const std::vector<V> &h, V **d) {
    const size_t bytes = sizeof(V) * std::max<size_t>(h.size(), 1);
    if (cudaMalloc((void **)d, bytes) != cudaSuccess) return false;
    if (!h.empty() && cudaMemcpy(*d, h.data(), sizeof(V) * h.size(), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    return true;
}

static int ensure_device(xan_mrtm_plan *pl) {
    if (pl->on_device) return XAN_OK;
    const Packing &pk = *pl->packing;
    bool ok = upload(pl->h_gcol, &pl->d_gcol);
    if (ok && pl->n_warps > 0) {
        std::vector<int> zeros(pk.n_warps, 0);
        ok = upload(pk.lane_cell, &pl->d_lane_cell) && upload(pk.lane_gedge, &pl->d_lane_gedge) &&
             upload(pk.lane_oedge, &pl->d_lane_oedge) && upload(pk.lane_src, &pl->d_lane_src) &&
             upload(pk.lane_meta, &pl->d_lane_meta) && upload(pk.edge_prod, &pl->d_edge_prod) &&
             upload(pk.edge_cons, &pl->d_edge_cons) && upload(zeros, &pl->d_progress) &&
             upload(pk.edge_cell, &pl->d_edge_cell);
    }
    if (!ok) {
        set_error("mrtm plan: CUDA allocation/copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        return XAN_E_CUDA;
    }
    pl->on_device = true;
    return XAN_OK;
}

extern "C" {

int xan_mrtm_downstream(const double *h_coords, const double *h_flow_dir, int ncell, int nrow, int ncol,
                        int64_t *h_dsid) {
    XAN_REQUIRE(h_coords && h_flow_dir && h_dsid && ncell > 0 && nrow > 0 && ncol > 0, "xan_mrtm_downstream: bad arguments");
    return host_downstream(h_coords, h_flow_dir, ncell, nrow, ncol, h_dsid);
}

int xan_mrtm_upstream(const double *h_coords, const int64_t *h_dsid, int ncell, int nrow, int ncol, int64_t *h_upid) {
    XAN_REQUIRE(h_coords && h_dsid && h_upid && ncell > 0 && nrow > 0 && ncol > 0, "xan_mrtm_upstream: bad arguments");
    for (int i = 0; i < ncell; ++i)
        XAN_REQUIRE(h_dsid[i] == -1 || (h_dsid[i] >= 1 && h_dsid[i] <= ncell), "xan_mrtm_upstream: dsid[%d] = %lld", i,
                    (long long)h_dsid[i]);
    return host_upstream(h_coords, h_dsid, ncell, nrow, ncol, h_upid);
}

xan_mrtm_plan *xan_mrtm_plan_create(const int64_t *h_upid, int ncell, int block_threads, int chunk_substeps) {
    if (!h_upid || ncell <= 0) {
        set_error("xan_mrtm_plan_create: bad arguments");
        return nullptr;
    }
    auto *pl = new xan_mrtm_plan();
    pl->ncell = ncell;
    if (build_rows(pl, h_upid) != XAN_OK) {
        delete pl;
        return nullptr;
    }
    const char *env_t = getenv("XANTHOS_MRTM_THREADS"), *env_c = getenv("XANTHOS_MRTM_CHUNK"),
               *env_l = getenv("XANTHOS_MRTM_LANES");
    pl->block_threads = (block_threads > 0) ? block_threads : (env_t ? atoi(env_t) : 640);
    pl->chunk = (chunk_substeps > 0) ? chunk_substeps : (env_c ? atoi(env_c) : 64);
    // lanes a warp may occupy (tests use small values); 31 keeps lane 31 empty, which the shuffle
    // exchange uses as its constant-zero source
    const int lanes = env_l ? atoi(env_l) : 31;
    pl->lanes = lanes;
    if (pl->block_threads % 32 != 0 || pl->block_threads < 32 || pl->block_threads > 640 || pl->chunk < 1 ||
        pl->chunk > 1024 || lanes < 9 || lanes > 31) {
        set_error("xan_mrtm_plan_create: block_threads must be a multiple of 32 in 32..640, chunk_substeps in "
                  "1..1024 (XANTHOS_MRTM_LANES in 9..31)");
        delete pl;
        return nullptr;
    }
    // grid-kernel gather table
    pl->h_gcol.assign((size_t)9 * ncell, -1);
    for (int i = 0; i < ncell; ++i)
        for (int s = pl->row_ptr[i]; s < pl->row_ptr[i + 1]; ++s)
            pl->h_gcol[(size_t)(s - pl->row_ptr[i]) * ncell + i] = pl->col[s] | (pl->sgn[s] < 0 ? (int)0x80000000 : 0);

    // The plan is host-side integer work (like the reference's upstream_genmatrix); the device
    // tables are uploaded by the first xan_mrtm_route call.
    pl->packing = new Packing();
    if (build_packing(pl, lanes, *pl->packing)) {
        pl->n_warps = pl->packing->n_warps;
        pl->n_edges = pl->packing->n_edges;
        pl->n_levels = pl->packing->n_levels;
        pl->G = pl->packing->G;
    } else {
        pl->n_warps = 0;   // warp kernel unavailable (graph with cycles)
    }
    return pl;
}

void xan_mrtm_plan_destroy(xan_mrtm_plan *pl) {
    if (!pl) return;
    if (pl->on_device) free_device(pl);
    xan::skew_plan_destroy(pl->skew);
    xan::skew_plan_destroy(pl->skew_multi);
    delete pl->packing;
    delete pl;
}

int xan_mrtm_plan_packing(const xan_mrtm_plan *pl, int *h_lane_cell, int *h_edge_prod, int *h_edge_cons) {
    XAN_REQUIRE(pl && pl->packing, "xan_mrtm_plan_packing: null plan");
    const Packing &pk = *pl->packing;
    if (h_lane_cell) std::copy(pk.lane_cell.begin(), pk.lane_cell.end(), h_lane_cell);
    if (h_edge_prod) std::copy(pk.edge_prod.begin(), pk.edge_prod.end(), h_edge_prod);
    if (h_edge_cons) std::copy(pk.edge_cons.begin(), pk.edge_cons.end(), h_edge_cons);
    return XAN_OK;
}

int xan_mrtm_plan_um_nnz(const xan_mrtm_plan *pl) { return pl ? pl->row_ptr[pl->ncell] : XAN_E_INVALID; }

int xan_mrtm_plan_um(const xan_mrtm_plan *pl, int64_t *h_indptr, int64_t *h_indices, int64_t *h_data) {
    XAN_REQUIRE(pl && h_indptr && h_indices && h_data, "xan_mrtm_plan_um: null pointer");
    for (int i = 0; i <= pl->ncell; ++i) h_indptr[i] = pl->row_ptr[i];
    for (int s = 0; s < pl->row_ptr[pl->ncell]; ++s) {
        h_indices[s] = pl->col[s];
        h_data[s] = pl->sgn[s];
    }
    return XAN_OK;
}

int xan_mrtm_plan_info(const xan_mrtm_plan *pl, int *info) {
    XAN_REQUIRE(pl && info, "xan_mrtm_plan_info: null pointer");
    info[0] = pl->is_forest ? 1 : 0;
    info[1] = pl->n_components;
    info[2] = pl->max_component;
    info[3] = pl->n_warps;
    info[4] = pl->n_edges;
    info[5] = pl->n_levels;
    info[6] = pl->block_threads;
    info[7] = pl->G;
    return XAN_OK;
}

}  // extern "C"

// Launch the warp kernel for `nm` (1..NM_MAX) members.  Returns XAN_OK, or XAN_E_INVALID when the
// blocks cannot all be resident (the caller may then fall back to the grid kernel).
template <int NM, int MAXT>
static int launch_warp_t(xan_mrtm_plan *pl, WarpArgs &a, int ntmax, int sms, int threads, cudaStream_t s) {
    const int wpb = threads / 32;
    int blocks = ceil_div(pl->n_warps, wpb);
    auto kernel = mrtm_warp_kernel<NM, MAXT>;
    XAN_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int per_sm = 0;
    size_t smem = 0;
    const char *esb = getenv("XANTHOS_MRTM_SB");
    // 50 sub-steps per staging block: 5 blocks per month at the reference's 3-hour step (224 - 248 sub-steps); every
    // block costs a prologue (row decode, first gather, staging hand-shake): 45.3 ms with 32, 44.5 ms with 50
    const int sb_first = esb ? std::max(8, std::min(64, atoi(esb))) : 50;
    for (int sb : {sb_first, 32, 16, 8}) {   // shrink the staging blocks until every warp fits on the device
        if (sb > sb_first) continue;
        a.sb = sb;
        smem = sizeof(double) * (size_t)wpb * (2 * ((size_t)pl->G + 1) * (sb + 1) * NM * 2);
        if (smem > 200 * 1024) continue;
        XAN_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
        if (per_sm * sms >= blocks) break;
    }
    if (per_sm * sms < blocks) {
        // the cut-edge pipeline needs every warp resident at the same time
        set_error("mrtm warp kernel: %d blocks of %d threads (%zu B smem, %d members) cannot be co-resident "
                  "(%d per SM x %d SMs)", blocks, threads, smem, NM, per_sm, sms);
        return XAN_E_INVALID;
    }
    const char *er = getenv("XANTHOS_MRTM_RING"), *es = getenv("XANTHOS_MRTM_SLEEP_NS");
    a.ring = er ? std::max(1, atoi(er)) : RING_DEFAULT;
    a.sleep_ns = es ? std::max(0, atoi(es)) : 100;

    double2 *ring = nullptr;
    const size_t ring_elems = (size_t)std::max(pl->n_edges, 1) * a.ring * ntmax * NM;
    XAN_CUDA_CHECK(scratch_alloc(&ring, sizeof(double2) * ring_elems, s));
    // hand-over counters of THIS launch (stream-ordered scratch): two routes on the same plan from different streams
    // share no mutable device state
    int *progress = nullptr;
    XAN_CUDA_CHECK(scratch_alloc(&progress, sizeof(int) * pl->n_warps, s));
    XAN_CUDA_CHECK(cudaMemsetAsync(progress, 0, sizeof(int) * pl->n_warps, s));
    a.progress = progress;
    a.ring_buf = ring;
    a.dbg = nullptr;
    a.sched = nullptr;
    const char *esch = getenv("XANTHOS_MRTM_SCHED");   // static | spread (no variant grouping) | grouped (default)
    if (!(esch && !strcmp(esch, "static")) && wpb % 4 == 0 && pl->n_warps <= 12288) {
        // one block per SM on every SM: the slots the packed warps do not need stay empty beside the expensive warps
        if (per_sm == 1 && wpb >= 16) blocks = std::max(blocks, sms);
        // The table depends on the topology and on which cells have dt V / L > 1; it is kept with the plan and
        // rebuilt when the static arrays, dt or the launch geometry change (a stale table costs time, never results).
        const char *eep = getenv("XANTHOS_MRTM_EXP_PER_SP");
        const int exp_per_sp = eep ? std::max(1, atoi(eep)) : 2;
        const char *epk = getenv("XANTHOS_MRTM_PACK_SM");
        const int pack_sm = epk ? atoi(epk) : 1;
        const int group = ((esch && !strcmp(esch, "spread")) ? 0 : 1) | (exp_per_sp << 8) | (pack_sm << 16);
        const xan_mrtm_plan::SchedKey key{a.flow_dist, a.velocity, a.dt, blocks, wpb, group};
        const xan_mrtm_plan::SchedKey &old = pl->sched_key;
        if (!pl->d_sched || old.flow_dist != key.flow_dist || old.velocity != key.velocity || old.dt != key.dt ||
            old.blocks != key.blocks || old.wpb != key.wpb || old.group != key.group) {
            // the previous table may still be read by a kernel in flight on another stream: stream-ordered free
            if (pl->d_sched) XAN_CUDA_CHECK(cudaFreeAsync(pl->d_sched, s));
            pl->d_sched = nullptr;
            XAN_CUDA_CHECK(scratch_alloc(&pl->d_sched, sizeof(int) * (size_t)blocks * wpb, s));
            mrtm_sched_kernel<<<1, 1024, sizeof(int) * pl->n_warps, s>>>(
                a.lane_cell, a.lane_gedge, a.lane_oedge, a.lane_meta, pl->d_edge_cell, a.flow_dist, a.velocity, a.dt,
                pl->n_warps, blocks, wpb, group, exp_per_sp, pack_sm, pl->d_sched);
            XAN_CUDA_CHECK(cudaGetLastError());
            pl->sched_key = key;
        }
        a.sched = pl->d_sched;
    }
    if (getenv("XANTHOS_MRTM_DEBUG")) XAN_CUDA_CHECK(scratch_alloc(&a.dbg, sizeof(long long) * 6 * pl->n_warps, s));
    void *kargs[] = {(void *)&a};
    // Optional (XANTHOS_MRTM_L2_PERSIST=1): pin the cut-edge rings (n_edges x ring x ntmax x 16 B = 33 MB for the 0.5
    // degree world) in L2 with a persisting access-policy window.  Measured in round 2: the kernel's time does not
    // change (44.6 ms either way - it moves 38 GB/s), and its DRAM traffic above the algorithmic 0.78 GB is not ring
    // eviction but the 32-byte sector granularity of the 8-byte runoff reads / ChStorage and Avg_ChFlow writes, which
    // are scattered by the depth-first cell order of the warps (reads 2.0x, writes 1.7x, proportional to the months).
    // Setting the window on the stream around every launch also serialises the ensemble runner's copy streams
    // (end to end 75 instead of 54 ms per member).  Hence off by default.
    const char *epers = getenv("XANTHOS_MRTM_L2_PERSIST");
    bool window = epers && atoi(epers) != 0;
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof(av));
    if (window) {
        static int max_persist = -1, max_window = 0;
        static size_t persist_limit = 0;
        if (max_persist < 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
            cudaGetLastError();
        }
        const size_t ring_bytes = sizeof(double2) * ring_elems;
        // the carve-out is taken from the L2 of every other kernel (the calibration passes re-use their forcing from
        // L2): only as much as the rings need, not the device maximum
        const size_t want = std::min((size_t)std::max(max_persist, 0), ring_bytes + ring_bytes / 4);
        if (want > persist_limit && max_persist > 0) {
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) persist_limit = want;
            cudaGetLastError();
        }
        window = max_persist > 0 && max_window > 0 && persist_limit > 0;
        if (window) {
            av.accessPolicyWindow.base_ptr = ring;
            av.accessPolicyWindow.num_bytes = std::min(ring_bytes, (size_t)max_window);
            av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist_limit / (double)av.accessPolicyWindow.num_bytes);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) {
                cudaGetLastError();
                window = false;
            }
        }
    }
    // cooperative launch = all blocks co-resident (no grid.sync is used)
    XAN_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)kernel, dim3(blocks), dim3(threads), kargs, smem, s));
    if (window) {   // the window applies to kernels launched while it is set: switch it off again for the stream
        av.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
    }
    if (a.dbg) {
        std::vector<long long> h(6 * (size_t)pl->n_warps);
        XAN_CUDA_CHECK(cudaMemcpyAsync(h.data(), a.dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, s));
        XAN_CUDA_CHECK(cudaStreamSynchronize(s));
        FILE *f = fopen(getenv("XANTHOS_MRTM_DEBUG"), "w");
        if (f) {
            for (int w = 0; w < pl->n_warps; ++w)
                fprintf(f, "%d %lld %lld %lld %lld %lld %lld\n", w, h[6 * w], h[6 * w + 1], h[6 * w + 2], h[6 * w + 3],
                        h[6 * w + 4], h[6 * w + 5]);
            fclose(f);
        }
        XAN_CUDA_CHECK(cudaFreeAsync(a.dbg, s));
    }
    XAN_CUDA_CHECK(cudaFreeAsync(ring, s));
    XAN_CUDA_CHECK(cudaFreeAsync(progress, s));
    return XAN_OK;
}

template <int NM>
static int launch_warp(xan_mrtm_plan *pl, WarpArgs &a, int ntmax, int sms, cudaStream_t s) {
    // the register budget follows the block size: up to 512 threads -> 128 registers, 640 -> 96 (one member per pass only)
    if constexpr (NM == 1) {
        if (pl->block_threads > 512) return launch_warp_t<NM, 640>(pl, a, ntmax, sms, pl->block_threads, s);
    }
    return launch_warp_t<NM, 512>(pl, a, ntmax, sms, std::min(pl->block_threads, 512), s);
}

static int route_grid(xan_mrtm_plan *pl, const double *d_runoff, const double *d_flow_dist, const double *d_velocity,
                      const double *d_area, const double *d_chs_prev, const int *d_ndays, int nmonths, int spinup_months,
                      int ld, double dt, double *d_chs, double *d_avg, double *d_instream, int sms, cudaStream_t s) {
    GridArgs g;
    g.gcol = pl->d_gcol;
    g.runoff = d_runoff;
    g.flow_dist = d_flow_dist;
    g.velocity = d_velocity;
    g.area = d_area;
    g.chs_prev = d_chs_prev;
    g.ndays = d_ndays;
    g.chs = d_chs;
    g.avg = d_avg;
    g.instream = d_instream;
    g.ncell = pl->ncell;
    g.nmonths = nmonths;
    g.spinup = spinup_months;
    g.ld = ld;
    g.dt = dt;
    double *work = nullptr;
    XAN_CUDA_CHECK(scratch_alloc(&work, sizeof(double) * 6 * (size_t)pl->ncell, s));
    g.S = work;
    g.Favg = work + pl->ncell;
    g.D = work + 2 * (size_t)pl->ncell;
    g.X = work + 3 * (size_t)pl->ncell;
    g.Y = work + 4 * (size_t)pl->ncell;
    g.Z = work + 5 * (size_t)pl->ncell;
    int per_sm = 0;
    XAN_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mrtm_grid_kernel, 256, 0));
    const int blocks = std::max(1, std::min(per_sm * sms, ceil_div(pl->ncell, 256)));
    void *kargs[] = {(void *)&g};
    XAN_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)mrtm_grid_kernel, dim3(blocks), dim3(256), kargs, 0, s));
    XAN_CUDA_CHECK(cudaFreeAsync(work, s));
    return XAN_OK;
}

extern "C" {

int xan_mrtm_route_batch(xan_mrtm_plan *pl, int n_members, const double *const *h_runoff, const double *d_flow_dist,
                         const double *d_velocity, const double *d_area, const double *const *h_chs_prev,
                         const int *h_ndays, int nmonths, int spinup_months, int ld, double dt, int method,
                         double *const *h_chs, double *const *h_avg, double *const *h_instream, void *stream) {
    XAN_REQUIRE(pl && h_runoff && d_flow_dist && d_velocity && d_area && h_ndays && n_members >= 1,
                "xan_mrtm_route_batch: null pointer");
    XAN_REQUIRE(nmonths > 0 && spinup_months >= 0 && spinup_months <= nmonths && ld >= pl->ncell && dt > 0,
                "xan_mrtm_route: bad arguments nmonths=%d spinup=%d ld=%d dt=%g", nmonths, spinup_months, ld, dt);
    for (int k = 0; k < n_members; ++k) XAN_REQUIRE(h_runoff[k], "xan_mrtm_route_batch: runoff of member %d is null", k);
    cudaStream_t s = (cudaStream_t)stream;
    {
        const int rc0 = ensure_device(pl);
        if (rc0 != XAN_OK) return rc0;
    }
    int ntmax = 0;
    for (int m = 0; m < nmonths; ++m) {
        const int nt = (int)((double)h_ndays[m] * 24 * 3600 / dt);
        XAN_REQUIRE(nt >= 1, "xan_mrtm_route: month %d has no sub-step (ndays=%d, dt=%g)", m, h_ndays[m], dt);
        ntmax = std::max(ntmax, nt);
    }
    const bool tree_ok = pl->is_forest && pl->n_warps > 0;
    XAN_REQUIRE(method != XAN_MRTM_TREE || tree_ok, "xan_mrtm_route: the flow graph is not a forest; warp kernel unavailable");
    XAN_REQUIRE(method >= XAN_MRTM_AUTO && method <= XAN_MRTM_SKEW, "xan_mrtm_route: unknown method %d", method);
    bool use_tree = (method == XAN_MRTM_TREE) || ((method == XAN_MRTM_AUTO || method == XAN_MRTM_SKEW) && tree_ok);
    // AUTO: XANTHOS_MRTM_AUTO=skew|tree selects the forest kernel (A/B runs); see AUTO_DEFAULT_SKEW
    const char *env_auto = getenv("XANTHOS_MRTM_AUTO");
    const bool auto_skew = env_auto ? !strcmp(env_auto, "skew") : AUTO_DEFAULT_SKEW;
    bool use_skew = pl->is_forest && (method == XAN_MRTM_SKEW || (method == XAN_MRTM_AUTO && auto_skew));
    XAN_REQUIRE(method != XAN_MRTM_SKEW || use_skew, "xan_mrtm_route: the flow graph is not a forest; skew kernel unavailable");

    int dev = 0, sms = 0;
    XAN_CUDA_CHECK(cudaGetDevice(&dev));
    XAN_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // month lengths on the device for the warp-dataflow / grid kernels, made on first need: the skew kernel keeps
    // its own cached step tables (a small host -> device copy here would queue behind the forcing uploads of an
    // ensemble run on the copy engine)
    int *d_ndays = nullptr;
    auto need_ndays = [&]() -> int {
        if (d_ndays) return XAN_OK;
        XAN_CUDA_CHECK(scratch_alloc(&d_ndays, sizeof(int) * nmonths, s));
        XAN_CUDA_CHECK(cudaMemcpyAsync(d_ndays, h_ndays, sizeof(int) * nmonths, cudaMemcpyHostToDevice, s));
        return XAN_OK;
    };

    const char *env_nm = getenv("XANTHOS_MRTM_MEMBERS");
    const int nm_cap = std::max(1, std::min(NM_MAX, env_nm ? atoi(env_nm) : 1));   // members per warp pass (2 measured slower: 69 vs 54 ms per member)
    const char *env_snm = getenv("XANTHOS_MRTM_SKEW_MEMBERS");
    int skew_nm = std::max(1, std::min(4, env_snm ? atoi(env_snm) : 2));
    int rc = XAN_OK;
    for (int k0 = 0; k0 < n_members && rc == XAN_OK;) {
        const int nm = std::min(nm_cap, n_members - k0);
        if (use_skew) {
            // several members per launch: their blocks share the SMs (XANTHOS_MRTM_SKEW_MEMBERS=1: off)
            const int nms = std::min(skew_nm, n_members - k0);
            const int r = xan::route_skew(pl, nms, h_runoff + k0, d_flow_dist, d_velocity, d_area,
                                          h_chs_prev ? h_chs_prev + k0 : nullptr, h_ndays, nmonths, spinup_months, ld, dt,
                                          h_chs ? h_chs + k0 : nullptr, h_avg ? h_avg + k0 : nullptr,
                                          h_instream ? h_instream + k0 : nullptr, sms, s);
            if (r == XAN_E_INVALID && nms > 1) {   // that many blocks per SM do not fit / are not compiled: one fewer
                skew_nm = nms - 1;
                continue;
            }
            if (r == XAN_E_INVALID) {   // plan or calendar not supported by the skew kernel
                XAN_REQUIRE(method != XAN_MRTM_SKEW, "xan_mrtm_route: the skew kernel cannot run this plan / calendar "
                            "(row with more than 4 tributaries on one side of the diagonal, or a month shorter than "
                            "the largest lag)");
                use_skew = false;
                continue;
            }
            rc = r;
            k0 += nms;
            continue;
        }
        {
            const int rcn = need_ndays();
            if (rcn != XAN_OK) return rcn;
        }
        if (use_tree) {
            WarpArgs a;
            memset(&a, 0, sizeof(a));
            a.lane_cell = pl->d_lane_cell;
            a.lane_gedge = pl->d_lane_gedge;
            a.lane_oedge = pl->d_lane_oedge;
            a.lane_src = pl->d_lane_src;
            a.lane_meta = pl->d_lane_meta;
            a.edge_prod = pl->d_edge_prod;
            a.edge_cons = pl->d_edge_cons;
            for (int k = 0; k < nm; ++k) {
                a.runoff[k] = h_runoff[k0 + k];
                a.chs_prev[k] = h_chs_prev ? h_chs_prev[k0 + k] : nullptr;
                a.chs[k] = h_chs ? h_chs[k0 + k] : nullptr;
                a.avg[k] = h_avg ? h_avg[k0 + k] : nullptr;
                a.instream[k] = h_instream ? h_instream[k0 + k] : nullptr;
            }
            a.flow_dist = d_flow_dist;
            a.velocity = d_velocity;
            a.area = d_area;
            a.ndays = d_ndays;
            a.n_warps = pl->n_warps;
            a.G = pl->G;
            a.ntmax = ntmax;
            a.nmonths = nmonths;
            a.spinup = spinup_months;
            a.ld = ld;
            a.dt = dt;
            int r = XAN_OK;
            if (nm == 1) r = launch_warp<1>(pl, a, ntmax, sms, s);
            else r = launch_warp<2>(pl, a, ntmax, sms, s);
            if (r == XAN_E_INVALID && method == XAN_MRTM_AUTO) {
                use_tree = false;   // does not fit: grid kernel for this and the remaining members
                continue;
            }
            rc = r;
            k0 += nm;
        } else {
            rc = route_grid(pl, h_runoff[k0], d_flow_dist, d_velocity, d_area, h_chs_prev ? h_chs_prev[k0] : nullptr,
                            d_ndays, nmonths, spinup_months, ld, dt, h_chs ? h_chs[k0] : nullptr,
                            h_avg ? h_avg[k0] : nullptr, h_instream ? h_instream[k0] : nullptr, sms, s);
            k0 += 1;
        }
    }
    if (d_ndays) cudaFreeAsync(d_ndays, s);
    return rc;
}

int xan_mrtm_route(xan_mrtm_plan *pl, const double *d_runoff, const double *d_flow_dist, const double *d_velocity,
                   const double *d_area, const double *d_chs_prev, const int *h_ndays, int nmonths, int spinup_months,
                   int ld, double dt, int method, double *d_chs, double *d_avg, double *d_instream, void *stream) {
    XAN_REQUIRE(d_runoff, "xan_mrtm_route: null pointer");
    return xan_mrtm_route_batch(pl, 1, &d_runoff, d_flow_dist, d_velocity, d_area, &d_chs_prev, h_ndays, nmonths,
                                spinup_months, ld, dt, method, &d_chs, &d_avg, &d_instream, stream);
}

}  // extern "C"
