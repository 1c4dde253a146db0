// clane_plan: device-side schedule + scratch of one (graph shape, row range).  See clane_b200.h.
#pragma once
#include "common.cuh"

struct clane_plan {
    int32_t n = 0, d = 0, ld = 0;
    int64_t e = 0;
    int32_t row_lo = 0, row_hi = 0;
    int64_t edge_lo = 0, edge_hi = 0;   // rowptr[row_lo], rowptr[row_hi]
    int32_t hub_threshold = 1024;
    bool has_schedule = false;
    int32_t G = 8;              // rows per group
    int32_t nslab = 1;
    int32_t fuse = 0;           // L1 change fused into the sweep (d in {32, 64, 128}, whole graph)
    int32_t n_groups = 0;       // groups covering [row_lo, row_hi)
    int32_t span_edges = 128;   // edge budget of a span
    int32_t n_spans = 0, n_fix_groups = 0, n_hub_rows = 0;
    int32_t n_long_hub_rows = 0;   // the first n_long_hub_rows hub rows (degree-descending) have >= kLongBlocks blocks
    // the sweep's schedule (sweep.cuh): tasks sorted by work descending (hub segments first)
    void* d_tasks = nullptr;           // SweepTask[n_tasks]
    int32_t n_tasks = 0;
    int32_t* d_fix_groups = nullptr;   // fused mode: groups whose chunk partial is recomputed from memory
    int32_t* d_hub_rows = nullptr;     // rows of degree > hub_threshold, degree-descending
    void* d_hub_info = nullptr;        // int4 per hub row: {row, first edge, degree, first scratch block}
    int64_t hub_blocks = 0;            // 8-neighbour blocks of all hub rows
    unsigned long long chain_spin_ns = 400000;   // early chain pass time-out: ~2x the expected time of all segment tasks
    int32_t limit = 0, ntail4 = 0, nslab32b = 0;   // 16*floor(d/16); float4 pieces beyond it; 32-column slabs below it
    void* d_hubS = nullptr;            // float4[hub_blocks][ld]  {z6, z4, X, Y}
    void* d_hubT = nullptr;            // float4[hub_blocks * 8][ntail4]  raw z, sequential-regime columns
    int32_t* d_hub_cnt = nullptr;      // per hub row: segment warps done this sweep
    int32_t* d_hub_done = nullptr;     // per chain CTA: produced by the early (overlapped) chain pass
    cudaStream_t side = nullptr, side2 = nullptr;   // the early chain passes (long rows / short rows) run here,
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;   // forked from / joined to the caller's stream
    // row-partitioned run: peer Znext buffers for the two Z ping-pong buffers (entry self = own buffer)
    int32_t n_peers = 0, self_rank = 0;
    float* peers[2][16] = {};
    float* mc[2] = {nullptr, nullptr};   // multicast (NVLS) addresses of the two buffers, or null
    int32_t* d_coloff = nullptr;       // col[e] * ld, rebuilt when the caller's column array changes
    const int32_t* coloff_src = nullptr;
    // CUDA-graph cache of whole sweeps (both streams, all kernels): a propagate() call ping-pongs between
    // two argument sets, so two entries suffice; anything else falls back to direct launches.
    struct SweepGraph {
        const void* key[10] = {nullptr};
        float gamma = 0.0f;
        int log_cap = 0;
        cudaGraphExec_t exec = nullptr;
        unsigned long long last_use = 0;
    } graphs[2];
    unsigned long long graph_clock = 0;
    bool use_graphs = true;
    bool profile = false;              // record timing events around the kernels of each sweep
    cudaEvent_t ev_prof[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float* d_P0 = nullptr;      // [n_groups][32]   (fused only)
    // cascade scratch: level-1 slots and level-2 slots, sized for max(n*d x1, e*d x2)
    float* d_p1 = nullptr;
    float* d_p2 = nullptr;
    size_t p1_floats = 0, p2_floats = 0;
};

extern "C" int clane_internal_prepare_kernels(void);
