// MRTM river routing (xanthos/routing/mrtm.py, xanthos/components.py:262-296), fp64, compiled with
// -fmad=false: the routing results are BIT-IDENTICAL to the reference's scipy-CSR formulation.
//
// Host side  : integer topology (downstream / upstream / UM = UP - I), forest check, tree
//              partition into pieces of bounded size, levels, packing of pieces into thread blocks.
// Device side:
//   * mrtm_tree_kernel - one persistent thread block per group of river (sub)trees.  Channel
//     storage S lives in registers for the whole run; the flows F of the block's cells live in
//     shared memory, and the reference's sparse "UM.dot(F)" is a <= 9-entry gather from that
//     buffer in ascending column order.  A sub-step needs ONE __syncthreads_or (two only when a
//     cell of the block was clamped).  River trees larger than a block are cut into sub-trees;
//     the flow over a cut edge travels downstream-only, so the upstream block simply runs ahead
//     and hands the per-sub-step flow series of the cut cell to the downstream block one month at
//     a time through a small ring buffer in global memory (acquire/release progress counters).
//   * mrtm_grid_kernel - general fallback for graphs that are not forests: cooperative launch,
//     two grid-wide syncs per sub-step, state in global memory.
#include "common.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

namespace cg = cooperative_groups;

namespace xan {
struct Packing;
}

struct xan_mrtm_plan {
    int ncell = 0;
    // ---- topology (host) -------------------------------------------------------------------
    std::vector<int> upid;      // [ncell][9] (mrtm.py:123-191)
    bool multi_receiver = false;
    std::vector<int> row_ptr;   // CSR of UM = UP - I, columns ascending (mrtm.py:194-230)
    std::vector<int> col;
    std::vector<signed char> sgn;
    std::vector<int> down;      // routing graph: 0-based receiver of cell j, -1 if none
    bool is_forest = false;
    int n_components = 0, max_component = 0;
    // ---- grid kernel ------------------------------------------------------------------------
    int *d_gcol = nullptr;              // [9][ncell] column (bit 31 set = minus sign), -1 = empty
    // ---- tree kernel ------------------------------------------------------------------------
    int T = 0, K = 0, C = 0;            // threads per block, cells per thread, slots per block
    int n_blocks = 0, n_edges = 0, n_levels = 0, G = 0;   // G = max ghosts per block
    int *d_slot_cell = nullptr;         // [n_blocks * C] cell index or -1
    uint4 *d_slot_nbr = nullptr;        // [n_blocks * C] local F index of upstream cell 0..7 (16 bit each, column order)
    unsigned *d_slot_meta = nullptr;    // [n_blocks * C] nup | ps << 4 | local index of the receiver << 16 (0xffff none)
    int *d_ghost_down = nullptr;        // [n_ghosts] local index of the cell fed by ghost k
    int *d_slot_out = nullptr;          // [n_blocks * C] cut edge fed by this cell or -1
    int *d_ghost_ptr = nullptr;         // [n_blocks + 1]
    int *d_ghost_edge = nullptr;        // [n_ghosts] cut edge read by ghost k
    int *d_edge_prod = nullptr;         // [n_edges] producing block
    int *d_edge_cons = nullptr;         // [n_edges] consuming block
    int *d_progress = nullptr;          // [n_blocks] months completed (reset per run)
    bool on_device = false;             // device tables are uploaded lazily by the first route()
    xan::Packing *packing = nullptr;    // host copy of the tree-kernel tables
    std::vector<int> h_gcol;
};

namespace xan {

constexpr int RING = 4;                 // months of cut-edge series kept in flight

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
// host: tree partition and block packing
// =============================================================================================
struct Packing {
    std::vector<int> slot_cell, slot_out, ghost_ptr, ghost_edge, ghost_down, edge_prod, edge_cons;
    std::vector<uint4> slot_nbr;
    std::vector<unsigned> slot_meta;
    int n_blocks = 0, n_edges = 0, n_levels = 0, G = 0;
};

static bool build_packing(xan_mrtm_plan *pl, int C, int T, int fill, Packing &pk) {
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

    // components (for reporting)
    {
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

    // bottom-up residual sizes; cut the largest children while a sub-tree exceeds `fill`
    std::vector<int> res(n, 0);
    std::vector<char> cut(n, 0);
    for (int h = 0; h < n; ++h) {
        const int v = order[h];
        const int k = pl->upid[(size_t)v * 9 + 8];
        int ch[8], sz = 1;
        for (int s = 0; s < k; ++s) {
            ch[s] = pl->upid[(size_t)v * 9 + s] - 1;
            sz += res[ch[s]];
        }
        if (sz > fill) {
            std::sort(ch, ch + k, [&](int a, int b) { return res[a] != res[b] ? res[a] > res[b] : a < b; });
            for (int s = 0; s < k && sz > fill; ++s) {
                cut[ch[s]] = 1;
                sz -= res[ch[s]];
            }
        }
        res[v] = sz;
    }
    // pieces: roots are outlets and cut cells
    std::vector<int> piece(n, -1), piece_root, piece_size, piece_level;
    for (int h = n - 1; h >= 0; --h) {
        const int v = order[h];
        if (pl->down[v] < 0 || cut[v]) {
            piece[v] = (int)piece_root.size();
            piece_root.push_back(v);
            piece_size.push_back(0);
        } else {
            piece[v] = piece[pl->down[v]];
        }
        piece_size[piece[v]]++;
    }
    const int np = (int)piece_root.size();
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

    // first-fit decreasing; linked pieces only share a block with pieces of the same level, which
    // keeps the block dependency graph acyclic.  Free pieces (whole small trees) fill the gaps.
    std::vector<int> ids(np);
    std::iota(ids.begin(), ids.end(), 0);
    std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return piece_size[a] > piece_size[b]; });
    std::vector<int> blk_fill, blk_level, piece_block(np, -1);
    for (int lvl = 0; lvl < n_levels; ++lvl) {
        for (int p : ids) {
            if (!linked[p] || piece_level[p] != lvl) continue;
            int b = -1;
            for (size_t q = 0; q < blk_fill.size(); ++q)
                if (blk_level[q] == lvl && blk_fill[q] + piece_size[p] <= fill) {
                    b = (int)q;
                    break;
                }
            if (b < 0) {
                b = (int)blk_fill.size();
                blk_fill.push_back(0);
                blk_level.push_back(lvl);
            }
            blk_fill[b] += piece_size[p];
            piece_block[p] = b;
        }
    }
    size_t first_open = 0;
    for (int p : ids) {
        if (linked[p]) continue;
        int b = -1;
        while (first_open < blk_fill.size() && blk_fill[first_open] >= fill) ++first_open;
        for (size_t q = first_open; q < blk_fill.size(); ++q)
            if (blk_fill[q] + piece_size[p] <= fill) {
                b = (int)q;
                break;
            }
        if (b < 0) {
            b = (int)blk_fill.size();
            blk_fill.push_back(0);
            blk_level.push_back(-1);
        }
        blk_fill[b] += piece_size[p];
        piece_block[p] = b;
    }
    // order blocks by level so that producers get the lower block indices
    const int nbk = (int)blk_fill.size();
    std::vector<int> bord(nbk), bnew(nbk);
    std::iota(bord.begin(), bord.end(), 0);
    std::stable_sort(bord.begin(), bord.end(), [&](int a, int b) {
        const int la = blk_level[a] < 0 ? n_levels : blk_level[a], lb = blk_level[b] < 0 ? n_levels : blk_level[b];
        return la < lb;
    });
    for (int q = 0; q < nbk; ++q) bnew[bord[q]] = q;

    // per block cell lists, heavy gather rows first (uniform work inside a warp)
    std::vector<std::vector<int>> cells(nbk);
    for (int v = 0; v < n; ++v) cells[bnew[piece_block[piece[v]]]].push_back(v);
    std::vector<int> cell_block(n), cell_slot(n);
    for (int b = 0; b < nbk; ++b) {
        std::stable_sort(cells[b].begin(), cells[b].end(), [&](int x, int y) {
            return pl->upid[(size_t)x * 9 + 8] > pl->upid[(size_t)y * 9 + 8];
        });
        if ((int)cells[b].size() > C) return false;
        for (size_t s = 0; s < cells[b].size(); ++s) {
            cell_block[cells[b][s]] = b;
            cell_slot[cells[b][s]] = (int)s;
        }
    }
    // cut edges = routing edges whose ends sit in different blocks
    pk.n_blocks = nbk;
    pk.n_levels = n_levels;
    pk.slot_cell.assign((size_t)nbk * C, -1);
    pk.slot_out.assign((size_t)nbk * C, -1);
    pk.slot_nbr.assign((size_t)nbk * C, make_uint4(0, 0, 0, 0));
    pk.slot_meta.assign((size_t)nbk * C, 0xffff0000u);
    std::vector<std::vector<int>> ghosts(nbk);   // producing cells seen by block b
    std::vector<int> edge_of_cell(n, -1);
    for (int v = 0; v < n; ++v) {
        const int r = pl->down[v];
        if (r >= 0 && cell_block[r] != cell_block[v]) {
            if (cell_block[v] > cell_block[r]) return false;   // would break the producer-first order
            edge_of_cell[v] = (int)pk.edge_prod.size();
            pk.edge_prod.push_back(cell_block[v]);
            pk.edge_cons.push_back(cell_block[r]);
            ghosts[cell_block[r]].push_back(v);
        }
    }
    pk.n_edges = (int)pk.edge_prod.size();
    pk.ghost_ptr.assign(nbk + 1, 0);
    for (int b = 0; b < nbk; ++b) {
        pk.ghost_ptr[b + 1] = pk.ghost_ptr[b] + (int)ghosts[b].size();
        pk.G = std::max(pk.G, (int)ghosts[b].size());
    }
    if (C + pk.G > 32767 || pk.G > T) return false;
    pk.ghost_edge.assign(std::max(pk.ghost_ptr[nbk], 1), -1);
    pk.ghost_down.assign(std::max(pk.ghost_ptr[nbk], 1), 0);
    std::vector<int> ghost_local(n, -1);   // local F index of producing cell v inside its consumer
    for (int b = 0; b < nbk; ++b)
        for (size_t k = 0; k < ghosts[b].size(); ++k) {
            pk.ghost_edge[pk.ghost_ptr[b] + k] = edge_of_cell[ghosts[b][k]];
            pk.ghost_down[pk.ghost_ptr[b] + k] = cell_slot[pl->down[ghosts[b][k]]];
            ghost_local[ghosts[b][k]] = C + (int)k;
        }
    for (int v = 0; v < n; ++v) {
        const int b = cell_block[v];
        const size_t g = (size_t)b * C + cell_slot[v];
        pk.slot_cell[g] = v;
        pk.slot_out[g] = edge_of_cell[v];
        unsigned short e[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int beg = pl->row_ptr[v], cnt = pl->row_ptr[v + 1] - beg;
        int nup = 0, ps = 0;
        for (int s = 0; s < cnt; ++s) {
            const int j = pl->col[beg + s];
            if (j == v) {
                ps = nup;   // the -F(self) term sits after `ps` upstream terms
                continue;
            }
            e[nup++] = (unsigned short)((cell_block[j] == b) ? cell_slot[j] : ghost_local[j]);
        }
        pk.slot_nbr[g] = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
        const int r = pl->down[v];
        const unsigned dl = (r >= 0 && cell_block[r] == b) ? (unsigned)cell_slot[r] : 0xffffu;
        pk.slot_meta[g] = (unsigned)nup | ((unsigned)ps << 4) | (dl << 16);
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

// Row i of UM times F, accumulated from 0.0 in ascending column order (scipy csr_matvec order):
// `nup` upstream terms (+F_j, read from shared memory) with the -F_i term after the first `ps`.
// Cells are sorted by nup inside a block, so the nested branches are (nearly) warp-uniform and a
// headwater cell (nup == 0, about half of all cells) touches no shared memory at all.
__device__ __forceinline__ double um_row(const double *__restrict__ Fb, uint4 nb, unsigned meta, double Fself) {
    const int nup = meta & 0xf, ps = (meta >> 4) & 0xf;
    double d = 0.0;
#define XAN_TERM(s, word, shift)                    \
    if (ps == (s)) d = d - Fself;                   \
    d = d + Fb[((word) >> (shift)) & 0xffffu];
    if (nup > 0) {
        XAN_TERM(0, nb.x, 0)
        if (nup > 1) {
            XAN_TERM(1, nb.x, 16)
            if (nup > 2) {
                XAN_TERM(2, nb.y, 0)
                if (nup > 3) {
                    XAN_TERM(3, nb.y, 16)
                    if (nup > 4) {
                        XAN_TERM(4, nb.z, 0)
                        if (nup > 5) {
                            XAN_TERM(5, nb.z, 16)
                            if (nup > 6) {
                                XAN_TERM(6, nb.w, 0)
                                if (nup > 7) {
                                    XAN_TERM(7, nb.w, 16)
                                }
                            }
                        }
                    }
                }
            }
        }
    }
#undef XAN_TERM
    if (ps == nup) d = d - Fself;
    return d;
}

// =============================================================================================
// tree kernel
// =============================================================================================
struct TreeArgs {
    const int *slot_cell;
    const uint4 *slot_nbr;
    const unsigned *slot_meta;
    const int *slot_out;
    const int *ghost_ptr;
    const int *ghost_edge;
    const int *ghost_down;
    const int *edge_prod;
    const int *edge_cons;
    int *progress;
    double *ring;            // [n_edges][RING][ntmax][2]
    const double *runoff;    // [M][ld]
    const double *flow_dist, *velocity, *area, *chs_prev;
    const int *ndays;        // [M] device
    double *chs, *avg, *instream;
    int C, G, ntmax, nmonths, spinup, ld;
    double dt;
};

template <int K>
__global__ void __launch_bounds__(K >= 3 ? 256 : 512, 1) mrtm_tree_kernel(const TreeArgs a) {
    extern __shared__ double smem[];
    const int T = blockDim.x, tid = threadIdx.x, b = blockIdx.x;
    const int C = a.C, W = a.C + a.G;
    double *X = smem, *Y = smem + W, *Z = smem + 2 * W;   // F, F', next F (X and Z swap)
    double *gs = smem + 3 * W;                            // [G][ntmax][2] staged ghost series
    unsigned char *dirty = reinterpret_cast<unsigned char *>(gs + (size_t)a.G * a.ntmax * 2);   // [C]
    const int g0 = a.ghost_ptr[b], ng = a.ghost_ptr[b + 1] - g0;

    int cell[K], oedge[K];
    uint4 nb[K];
    unsigned meta[K];
    double S[K], tauinv[K], area[K], Favg[K], erl[K], qn[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const size_t g = (size_t)b * C + tid + j * T;
        cell[j] = a.slot_cell[g];
        oedge[j] = -1;
        S[j] = 0.0; tauinv[j] = 0.0; area[j] = 0.0; Favg[j] = 0.0; erl[j] = 0.0; qn[j] = 0.0;
        nb[j] = make_uint4(0, 0, 0, 0);
        meta[j] = 0xffff0000u;
        dirty[tid + j * T] = 0;
        if (cell[j] >= 0) {
            nb[j] = a.slot_nbr[g];
            meta[j] = a.slot_meta[g];
            oedge[j] = a.slot_out[g];
            tauinv[j] = a.velocity[cell[j]] / a.flow_dist[cell[j]];                 // mrtm.py:42
            area[j] = a.area[cell[j]];
            S[j] = a.chs_prev ? a.chs_prev[cell[j]] : 0.0;
            qn[j] = a.runoff[cell[j]];   // month 0 of the first pass
        }
    }
    const int my_edge = (tid < ng) ? a.ghost_edge[g0 + tid] : -1;
    const int my_prod = (my_edge >= 0) ? a.edge_prod[my_edge] : -1;
    const int my_gdown = (my_edge >= 0) ? a.ghost_down[g0 + tid] : 0;
    const double dt = a.dt, dtinv = 1. / a.dt;                                      // mrtm.py:43
    const int nsteps = a.spinup + a.nmonths;

    for (int step = 0; step < nsteps; ++step) {
        const bool store = step >= a.spinup;
        const int m = store ? step - a.spinup : step;
        const int nday = a.ndays[m];
        const int nt = (int)((double)nday * 24 * 3600 / dt);                        // mrtm.py:36
        const double secs = (double)(nday * 24 * 3600);
        const int slot = step % RING;
        // ---- wait: producers have finished this month; consumers have freed the ring slot -------
        if (my_prod >= 0)
            while (ld_acquire(a.progress + my_prod) < step + 1) __nanosleep(64);
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (oedge[j] >= 0 && step >= RING) {
                const int cb = a.edge_cons[oedge[j]];
                while (ld_acquire(a.progress + cb) < step - RING + 1) __nanosleep(64);
            }
        __syncthreads();
        // ---- stage the ghost series of this month into shared memory -----------------------------
        for (int k = 0; k < ng; ++k) {
            const double *src = a.ring + ((size_t)a.ghost_edge[g0 + k] * RING + slot) * a.ntmax * 2;
            double *dst = gs + (size_t)k * a.ntmax * 2;
            for (int i = tid; i < nt * 2; i += T) dst[i] = __ldcg(src + i);
        }
        // ---- month setup ---------------------------------------------------------------------------
        const int mnext = (step + 1 < nsteps) ? ((step + 1 >= a.spinup) ? step + 1 - a.spinup : step + 1) : m;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            erl[j] = (qn[j] * area[j]) * (1e6 / 1e3) / secs;                        // mrtm.py:45
            Favg[j] = 0.0;
            if (cell[j] >= 0) {
                qn[j] = a.runoff[(size_t)mnext * a.ld + cell[j]];                   // prefetch next month
                X[tid + j * T] = S[j] * tauinv[j];                                  // mrtm.py:50
            }
        }
        __syncthreads();   // staged series visible
        if (tid < ng) X[C + tid] = gs[(size_t)tid * a.ntmax * 2 + 0];
        __syncthreads();

        for (int t = 0; t < nt; ++t) {
            double F[K], Fp[K], Sn[K];
            bool clamp[K];
            int flag = 0;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                F[j] = S[j] * tauinv[j];
                Fp[j] = F[j];
                Sn[j] = S[j];
                clamp[j] = false;
                if (cell[j] >= 0) {
                    const double d = um_row(X, nb[j], meta[j], F[j]) + erl[j];      // mrtm.py:51
                    clamp[j] = (d * dt) < (-S[j]);                                  // mrtm.py:54
                    if (clamp[j]) {
                        Fp[j] = d + F[j] + S[j] * dtinv;                            // mrtm.py:60
                        Sn[j] = 0.0;                                                // mrtm.py:63
                        const unsigned dl = meta[j] >> 16;
                        if (dl != 0xffffu) {   // the receiver must redo its balance with F'
                            dirty[dl] = 1;
                            flag = 1;
                        }
                    } else {
                        Sn[j] = S[j] + d * dt;                                      // mrtm.py:76
                    }
                    Y[tid + j * T] = Fp[j];
                    Z[tid + j * T] = Sn[j] * tauinv[j];   // speculative flow of the next sub-step
                }
            }
            if (tid < ng) {
                const double *g = gs + ((size_t)tid * a.ntmax + t) * 2;
                Y[C + tid] = g[1];
                if (t + 1 < nt) Z[C + tid] = g[2];
                if (__double_as_longlong(g[0]) != __double_as_longlong(g[1])) {
                    dirty[my_gdown] = 1;
                    flag = 1;
                }
            }
            if (__syncthreads_or(flag)) {
                // an inflow changed in the clamp pass: the receivers redo the balance with F' (mrtm.py:66-69)
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    if (cell[j] >= 0 && dirty[tid + j * T]) {
                        dirty[tid + j * T] = 0;
                        if (!clamp[j]) {
                            const double d2 = um_row(Y, nb[j], meta[j], F[j]) + erl[j];
                            Sn[j] = S[j] + d2 * dt;
                            Z[tid + j * T] = Sn[j] * tauinv[j];
                        }
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int j = 0; j < K; ++j) {
                S[j] = Sn[j];
                Favg[j] += Fp[j];                                                   // mrtm.py:78
                if (oedge[j] >= 0) {
                    double *r = a.ring + (((size_t)oedge[j] * RING + slot) * a.ntmax + t) * 2;
                    __stcg(r, F[j]);
                    __stcg(r + 1, Fp[j]);
                }
            }
            double *tmp = X; X = Z; Z = tmp;
        }
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (cell[j] >= 0) {
                if (store) {
                    if (a.chs) stg_stream(a.chs + (size_t)m * a.ld + cell[j], S[j]);
                    if (a.avg) stg_stream(a.avg + (size_t)m * a.ld + cell[j], Favg[j] / nt);   // mrtm.py:80
                }
                // instantaneous flow after the last sub-step = last published F' (still in Y)
                if (step == nsteps - 1 && a.instream) a.instream[cell[j]] = Y[tid + j * T];
            }
        // ---- publish: this block has finished month `step` ------------------------------------------
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(a.progress + b, step + 1);
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

template <int K>
static int launch_tree(const xan_mrtm_plan *pl, TreeArgs &args, size_t smem, cudaStream_t s) {
    XAN_CUDA_CHECK(cudaFuncSetAttribute(mrtm_tree_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, dev = 0, sms = 0;
    XAN_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mrtm_tree_kernel<K>, pl->T, smem));
    XAN_CUDA_CHECK(cudaGetDevice(&dev));
    XAN_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (per_sm * sms < pl->n_blocks) {
        set_error("mrtm tree kernel: %d blocks cannot be co-resident (%d per SM x %d SMs)", pl->n_blocks, per_sm, sms);
        return XAN_E_INVALID;
    }
    void *kargs[] = {(void *)&args};
    // cooperative launch = all blocks co-resident, which the cut-edge pipeline relies on
    XAN_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)mrtm_tree_kernel<K>, dim3(pl->n_blocks), dim3(pl->T), kargs, smem, s));
    return XAN_OK;
}

}  // namespace xan

using namespace xan;

static void free_device(xan_mrtm_plan *pl) {
    cudaFree(pl->d_gcol);
    cudaFree(pl->d_slot_cell);
    cudaFree(pl->d_slot_nbr);
    cudaFree(pl->d_slot_meta);
    cudaFree(pl->d_ghost_down);
    cudaFree(pl->d_slot_out);
    cudaFree(pl->d_ghost_ptr);
    cudaFree(pl->d_ghost_edge);
    cudaFree(pl->d_edge_prod);
    cudaFree(pl->d_edge_cons);
    cudaFree(pl->d_progress);
}

template <typename V>
static bool upload(const std::vector<V> &h, V **d) {
    const size_t bytes = sizeof(V) * std::max<size_t>(h.size(), 1);
    if (cudaMalloc((void **)d, bytes) != cudaSuccess) return false;
    if (!h.empty() && cudaMemcpy(*d, h.data(), sizeof(V) * h.size(), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    return true;
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

xan_mrtm_plan *xan_mrtm_plan_create(const int64_t *h_upid, int ncell, int block_threads, int cells_per_thread) {
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
    // The plan itself is host-side integer work (like the reference's downstream/upstream); the
    // device tables are uploaded by the first xan_mrtm_route call.  Without a device the SM count
    // of a B200 is assumed for the fill target so that the packing can still be inspected.
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
        cudaGetLastError();
        sms = 148;
    }
    const char *env_t = getenv("XANTHOS_MRTM_THREADS"), *env_k = getenv("XANTHOS_MRTM_CELLS_PER_THREAD");
    pl->T = (block_threads > 0) ? block_threads : (env_t ? atoi(env_t) : 256);
    pl->K = (cells_per_thread > 0) ? cells_per_thread : (env_k ? atoi(env_k) : 2);
    if (pl->T % 32 != 0 || pl->T > 512 || pl->K > 4 || (pl->K >= 3 && pl->T > 256)) {
        set_error("xan_mrtm_plan_create: block_threads must be a multiple of 32, <= 512 (<= 256 when "
                  "cells_per_thread >= 3), and cells_per_thread <= 4");
        delete pl;
        return nullptr;
    }
    pl->C = pl->T * pl->K;
    // fill target: spread the cells over (a multiple of) the SM count, never above the capacity
    const char *env = getenv("XANTHOS_MRTM_BLOCK_CELLS");
    int fill = env ? atoi(env) : 0;
    if (fill <= 0) {
        const int waves = std::max(1, (ncell + sms * pl->C - 1) / (sms * pl->C));
        fill = (ncell + sms * waves - 1) / (sms * waves);
        fill = std::max(fill + fill / 32 + 1, 32);
    }
    fill = std::min(fill, pl->C);

    // grid-kernel gather table
    pl->h_gcol.assign((size_t)9 * ncell, -1);
    for (int i = 0; i < ncell; ++i)
        for (int s = pl->row_ptr[i]; s < pl->row_ptr[i + 1]; ++s)
            pl->h_gcol[(size_t)(s - pl->row_ptr[i]) * ncell + i] = pl->col[s] | (pl->sgn[s] < 0 ? (int)0x80000000 : 0);

    pl->packing = new Packing();
    if (build_packing(pl, pl->C, pl->T, fill, *pl->packing)) {
        pl->n_blocks = pl->packing->n_blocks;
        pl->n_edges = pl->packing->n_edges;
        pl->n_levels = pl->packing->n_levels;
        pl->G = pl->packing->G;
    } else {
        pl->n_blocks = 0;   // tree kernel unavailable (graph with cycles or packing failure)
    }
    return pl;
}

static int ensure_device(xan_mrtm_plan *pl) {
    if (pl->on_device) return XAN_OK;
    const Packing &pk = *pl->packing;
    bool ok = upload(pl->h_gcol, &pl->d_gcol);
    if (ok && pl->n_blocks > 0) {
        std::vector<int> zeros(pk.n_blocks, 0);
        ok = upload(pk.slot_cell, &pl->d_slot_cell) && upload(pk.slot_nbr, &pl->d_slot_nbr) &&
             upload(pk.slot_meta, &pl->d_slot_meta) && upload(pk.slot_out, &pl->d_slot_out) &&
             upload(pk.ghost_down, &pl->d_ghost_down) &&
             upload(pk.ghost_ptr, &pl->d_ghost_ptr) && upload(pk.ghost_edge, &pl->d_ghost_edge) &&
             upload(pk.edge_prod, &pl->d_edge_prod) && upload(pk.edge_cons, &pl->d_edge_cons) &&
             upload(zeros, &pl->d_progress);
    }
    if (!ok) {
        set_error("mrtm plan: CUDA allocation/copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        return XAN_E_CUDA;
    }
    pl->on_device = true;
    return XAN_OK;
}

void xan_mrtm_plan_destroy(xan_mrtm_plan *pl) {
    if (!pl) return;
    if (pl->on_device) free_device(pl);
    delete pl->packing;
    delete pl;
}

/* test/diagnostic export of the tree-kernel packing: slot_cell [n_blocks * T * K] (cell or -1),
 * edge_prod / edge_cons [n_cut_edges] (block indices).  Any pointer may be NULL. */
int xan_mrtm_plan_packing(const xan_mrtm_plan *pl, int *h_slot_cell, int *h_edge_prod, int *h_edge_cons) {
    XAN_REQUIRE(pl && pl->packing, "xan_mrtm_plan_packing: null plan");
    const Packing &pk = *pl->packing;
    if (h_slot_cell) std::copy(pk.slot_cell.begin(), pk.slot_cell.end(), h_slot_cell);
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
    info[3] = pl->n_blocks;
    info[4] = pl->n_edges;
    info[5] = pl->n_levels;
    info[6] = pl->T;
    info[7] = pl->K;
    return XAN_OK;
}

int xan_mrtm_route(xan_mrtm_plan *pl, const double *d_runoff, const double *d_flow_dist, const double *d_velocity,
                   const double *d_area, const double *d_chs_prev, const int *h_ndays, int nmonths, int spinup_months,
                   int ld, double dt, int method, double *d_chs, double *d_avg, double *d_instream, void *stream) {
    XAN_REQUIRE(pl && d_runoff && d_flow_dist && d_velocity && d_area && h_ndays, "xan_mrtm_route: null pointer");
    XAN_REQUIRE(nmonths > 0 && spinup_months >= 0 && spinup_months <= nmonths && ld >= pl->ncell && dt > 0,
                "xan_mrtm_route: bad arguments nmonths=%d spinup=%d ld=%d dt=%g", nmonths, spinup_months, ld, dt);
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
    const bool tree_ok = pl->is_forest && pl->n_blocks > 0;
    XAN_REQUIRE(method != XAN_MRTM_TREE || tree_ok, "xan_mrtm_route: the flow graph is not a forest; tree kernel unavailable");
    const bool use_tree = (method == XAN_MRTM_TREE) || (method == XAN_MRTM_AUTO && tree_ok);

    int *d_ndays = nullptr;
    XAN_CUDA_CHECK(cudaMallocAsync(&d_ndays, sizeof(int) * nmonths, s));
    XAN_CUDA_CHECK(cudaMemcpyAsync(d_ndays, h_ndays, sizeof(int) * nmonths, cudaMemcpyHostToDevice, s));
    int rc = XAN_OK;
    if (use_tree) {
        TreeArgs a;
        a.slot_cell = pl->d_slot_cell;
        a.slot_nbr = pl->d_slot_nbr;
        a.slot_meta = pl->d_slot_meta;
        a.ghost_down = pl->d_ghost_down;
        a.slot_out = pl->d_slot_out;
        a.ghost_ptr = pl->d_ghost_ptr;
        a.ghost_edge = pl->d_ghost_edge;
        a.edge_prod = pl->d_edge_prod;
        a.edge_cons = pl->d_edge_cons;
        a.progress = pl->d_progress;
        a.runoff = d_runoff;
        a.flow_dist = d_flow_dist;
        a.velocity = d_velocity;
        a.area = d_area;
        a.chs_prev = d_chs_prev;
        a.ndays = d_ndays;
        a.chs = d_chs;
        a.avg = d_avg;
        a.instream = d_instream;
        a.C = pl->C;
        a.G = pl->G;
        a.ntmax = ntmax;
        a.nmonths = nmonths;
        a.spinup = spinup_months;
        a.ld = ld;
        a.dt = dt;
        double *ring = nullptr;
        const size_t ring_elems = (size_t)std::max(pl->n_edges, 1) * RING * ntmax * 2;
        XAN_CUDA_CHECK(cudaMallocAsync(&ring, sizeof(double) * ring_elems, s));
        XAN_CUDA_CHECK(cudaMemsetAsync(pl->d_progress, 0, sizeof(int) * pl->n_blocks, s));
        a.ring = ring;
        const size_t smem = sizeof(double) * (3 * (size_t)(pl->C + pl->G) + (size_t)pl->G * ntmax * 2) + (size_t)pl->C + 16;
        if (smem > 227 * 1024) {
            set_error("xan_mrtm_route: tree kernel needs %zu B of shared memory (G=%d, ntmax=%d)", smem, pl->G, ntmax);
            rc = XAN_E_INVALID;
        } else {
            switch (pl->K) {
                case 1: rc = launch_tree<1>(pl, a, smem, s); break;
                case 2: rc = launch_tree<2>(pl, a, smem, s); break;
                case 3: rc = launch_tree<3>(pl, a, smem, s); break;
                default: rc = launch_tree<4>(pl, a, smem, s); break;
            }
        }
        cudaFreeAsync(ring, s);
        if (rc != XAN_OK && method == XAN_MRTM_AUTO) rc = 1;   // fall through to the grid kernel
    }
    if (!use_tree || rc == 1) {
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
        XAN_CUDA_CHECK(cudaMallocAsync(&work, sizeof(double) * 6 * (size_t)pl->ncell, s));
        g.S = work;
        g.Favg = work + pl->ncell;
        g.D = work + 2 * (size_t)pl->ncell;
        g.X = work + 3 * (size_t)pl->ncell;
        g.Y = work + 4 * (size_t)pl->ncell;
        g.Z = work + 5 * (size_t)pl->ncell;
        int per_sm = 0, dev = 0, sms = 0;
        XAN_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mrtm_grid_kernel, 256, 0));
        XAN_CUDA_CHECK(cudaGetDevice(&dev));
        XAN_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int blocks = std::max(1, std::min(per_sm * sms, ceil_div(pl->ncell, 256)));
        void *kargs[] = {(void *)&g};
        XAN_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)mrtm_grid_kernel, dim3(blocks), dim3(256), kargs, 0, s));
        XAN_CUDA_CHECK(cudaFreeAsync(work, s));
        rc = XAN_OK;
    }
    XAN_CUDA_CHECK(cudaFreeAsync(d_ndays, s));
    return rc;
}

}  // extern "C"
