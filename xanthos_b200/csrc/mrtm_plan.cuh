// The routing plan shared by the MRTM kernels (mrtm.cu: warp-dataflow and grid kernels; mrtm_skew.cu: skewed
// multi-cell-per-lane kernel).  Host-side topology is built once by xan_mrtm_plan_create (mrtm.cu).
#pragma once

#include "common.cuh"

#include <vector>

namespace xan {
struct SkewPlan;
void skew_plan_destroy(SkewPlan *sp);
}
struct xan_mrtm_plan;
namespace xan {
// mrtm_skew.cu: routes nm (1 or 2) members with one launch of the skew kernel; XAN_E_INVALID (no error message set) =
// not applicable
int route_skew(xan_mrtm_plan *pl, int nm, const double *const *d_runoff, const double *d_flow_dist,
               const double *d_velocity, const double *d_area, const double *const *d_chs_prev, const int *h_ndays,
               int nmonths, int spinup_months, int ld, double dt, double *const *d_chs, double *const *d_avg,
               double *const *d_instream, int sms, cudaStream_t s);
}

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
    std::vector<int> h_gcol;
    int *d_gcol = nullptr;              // [9][ncell] column (bit 31 set = minus sign), -1 = empty
    // ---- warp kernel ------------------------------------------------------------------------
    int block_threads = 256, chunk = 64, lanes = 31;   // lanes: lanes of a warp the packing may occupy
    int n_warps = 0, n_edges = 0, n_levels = 0, G = 0;   // G = max ghost lanes of a warp
    xan::Packing *packing = nullptr;    // host copy of the lane tables
    int *d_lane_cell = nullptr;         // [n_warps * 32] cell index or -1
    int *d_lane_gedge = nullptr;        // [n_warps * 32] cut edge READ by this (ghost) lane or -1
    int *d_lane_oedge = nullptr;        // [n_warps * 32] cut edge WRITTEN by this lane or -1
    uint2 *d_lane_src = nullptr;        // [n_warps * 32] row of UM in column order, 9 x 7 bit (see build_packing)
    unsigned *d_lane_meta = nullptr;    // [n_warps * 32] row length | ghost slot << 8
    int *d_edge_prod = nullptr;         // [n_edges] producing warp
    int *d_edge_cons = nullptr;         // [n_edges] consuming warp
    int *d_progress = nullptr;          // [n_warps] chunks completed (reset per run)
    int *d_edge_cell = nullptr;         // [n_edges] cell whose flow the edge carries
    struct SchedKey {
        const double *flow_dist, *velocity;
        double dt;
        int blocks, wpb, group;
    };
    int *d_sched = nullptr;             // [grid warps] packed warp run by each grid warp (mrtm_sched_kernel), cached
    SchedKey sched_key{};
    bool on_device = false;
    // ---- skew kernel (mrtm_skew.cu), built on first use ----------------------------------------
    xan::SkewPlan *skew = nullptr, *skew_multi = nullptr;   // single-member launches / several members per launch
    bool skew_tried = false, skew_multi_tried = false;
};

