// wfa_plan.cuh -- the routing plan shared by the sweep (wfa.cu) and the sparse routing operator (router.cu).
#pragma once

#include <vector>

#include "common.cuh"

struct wfa_plan {
    int device = 0, nrow = 0, ncol = 0, nlevels = 0, norder = 0, tail_level = 0, flags = 0;
    long long ncell = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm = nullptr;   // start, end, end of the wide-level phase
    bool timed = false;
    unsigned epoch = 0;
    DevBuf receiver, inmask, omask, indeg, level, order, level_off, contam, tailmask, done, ticket, work, work2, qs, src, poff,
        pcells, qr;
    std::vector<int32_t> h_level_off;
    // upper forest regrouped by outlet basin (see k_accum_components)
    int comp_level = 0, ncomp = 0, nupper = 0;
    bool comp_ok = false;
    DevBuf corder, cmask, cnext, comps, cup_off, cup_idx;
    // upper network regrouped by river reach (see k_reach_sweep)
    bool run_ok = false;
    int run_critical = 0;
    int nruns = 0;
    DevBuf rcell, rup_off, rup_idx, run_start, row_off, rowcell, rstream, rinfo, rdone;
    int nrows = 0;
    std::vector<int32_t> h_rlv_off;      // first reach of every reach level
    // chain segments of the levels below the cut (see k_accum_segments)
    bool seg_ok = false;
    int seg_levels = 0;                  // segments cover levels [1, seg_levels)
    long long nseg = 0, nseg_cells = 0;
    DevBuf segs;                         // SegDesc[nseg], heads in (level, cell) order
    std::vector<int32_t> h_seg_off;      // first segment of every level
};
