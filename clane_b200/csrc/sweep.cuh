// Jacobi sweep of Embedder.propagate (/root/reference/clane/embedder.py:84-94) for sm_100a.
//
//   z_v <- fl(x_v + fl(gamma * (w_v[1,k] @ Zcur[nbrs(v)][k,d])))        for every row with k > 0
//
// with the [1,k]x[k,d] product in oneMKL sgemm's summation order (SURVEY 7.1 step 4 /
// Appendix A.1): per output column, neighbours in ascending column id;
//   k < 8 or column >= 16*floor(d/16): sequential fma chain;
//   else per block of 8 neighbours:  a = fma(w6,z6,a); a = fma(w4,z4,a);
//        a += fma(w5,z5, w7*z7);  a += fma(w0,z0, w2*z2) + fma(w1,z1, w3*z3);
//   then the k mod 8 leftovers sequentially.
// A row's neighbours are therefore never split across lanes: parallelism is over columns
// (one float4 of columns per lane) and rows, and memory-level parallelism comes from issuing
// the (address-independent) gathers ahead of the in-order chains.
//
// The graph is static, so ALL control flow of the sweep is precomputed on the host (clane_plan_create,
// "program"): the kernel only decodes 32-bit batch descriptors.
//
//   task      : one warp per (task, 128-column slab).  A span task is a run of ordinary rows of one row
//               group (<= span_edges edges); a segment task is 128 neighbours (16 blocks) of a hub row.
//   batch     : up to 8 neighbours of one row (one 8-block, or the row's k mod 8 leftovers); a row's last
//               batch also carries the row's X piece and (fused L1) its own Zcur piece.  Every piece is
//               one 128-bit load per lane straight into registers -- lane L loads exactly the float4 it
//               will reduce -- and batches are double-buffered: the loads of batch b + 1 are in flight
//               while batch b is reduced (two 8 x float4 register buffers; no shared-memory staging, so
//               each gathered byte crosses the SM's L1 / shared-memory data path once, not twice).
//   descriptor: m | last | publish-meta-window | row | position in the (offset, w) ring.
//   hub rows  : the reference's 8-neighbour block is  a = fma(w6,z6,a); a = fma(w4,z4,a); a += X; a += Y
//               with X = fma(w5,z5,w7*z7), Y = fma(w0,z0,w2*z2) + fma(w1,z1,w3*z3) independent of the
//               running sum.  Segment tasks (all SMs, inside k_sweep_rows) gather the neighbours and park
//               {z6, z4, X, Y} per (block, column); k_hub_chain then runs the 4-op chain per block over
//               that contiguous stream (one warp per (hub row, 32 columns), cp.async ring).  Columns in
//               the sequential regime (>= 16*floor(d/16)) are parked raw and chained by one more warp.
#pragma once
#include "common.cuh"
#include "program.cuh"

namespace clane {

struct SweepParams {
    const float* X;
    const float* Zc;
    float* Zn;
    int ld, d, n;
    const int32_t* rowptr;
    const int32_t* coloff;      // col[e] * ld / 4: float4 index of the neighbour's row
    const float* w;
    float gamma;
    const SweepTask* tasks;     // sorted by work, descending (segments first)
    const int32_t* descs;
    int n_tasks;
    int row_lo;
    int G;                      // rows per group (<= 32)
    int nslab;                  // 128-column slabs per row
    int fuse;                   // 1: direct spans write one 32-lane |delta| partial per group to P0
    float* P0;
    // hub rows
    const int32_t* hub_rows;    // rows of degree > hub_threshold, degree-descending
    const int32_t* hub_blk0;    // first scratch block of each hub row
    int n_hub_rows;
    int limit;                  // 16 * floor(d / 16): columns below it use the 8-block order
    int ntail4;                 // (ld - limit) / 4: float4 pieces per row in the sequential regime (0..4)
    int nslab32b;               // 32-column slabs below `limit`
    int sld;                    // 32 * nslab32b: columns per block in hubS
    float4* hubS;               // per hub row [32-column slab][block][32] {z6, z4, X, Y}
    float2* hubW;               // [block] {w4, w6}
    float4* hubT;               // [block * 8 + i][ntail4] raw z of the sequential-regime columns
    const clane_patience* st;
};

constexpr int kRowThreads = 128;               // row kernel: 4 warps per CTA
constexpr int kRowWarps = 4;
// row kernel shared memory per warp: (offset, w) ring | 512-byte transpose scratch (fused L1)
constexpr size_t kRowWarpSmem = (size_t)(kMetaRing + 8) * sizeof(int2) + 512;
constexpr size_t kRowSmemBytes = (size_t)kRowWarps * kRowWarpSmem;

// hub chain kernel: one warp per CTA
constexpr int kChainGroup = 16;                // blocks per cp.async group
constexpr int kChainGroups = 8;                // groups in flight (8 x 16 x 512 B = 64 KB)
constexpr size_t kChainSmemBytes = (size_t)kChainGroups * kChainGroup * 32 * sizeof(float4) +
                                   (size_t)kChainGroups * kChainGroup * sizeof(float2);
constexpr int kTailGroup = 32;                 // neighbours per group of the sequential-regime chain
constexpr int kTailGroups = 16;
constexpr size_t kTailSmemBytes = (size_t)kTailGroups * kTailGroup * (4 * sizeof(float4) + sizeof(float));
static_assert(kTailSmemBytes <= kChainSmemBytes, "one dynamic shared memory size for both roles");

__device__ __forceinline__ void fma4(float wv, const float4& z, float4& acc) {
    acc.x = ffma(wv, z.x, acc.x); acc.y = ffma(wv, z.y, acc.y);
    acc.z = ffma(wv, z.z, acc.z); acc.w = ffma(wv, z.w, acc.w);
}

__device__ __forceinline__ float blocked8(float a, const float* w, float z0, float z1, float z2, float z3,
                                          float z4, float z5, float z6, float z7) {
    a = ffma(w[6], z6, a);
    a = ffma(w[4], z4, a);
    a = fadd(a, ffma(w[5], z5, fmul(w[7], z7)));
    a = fadd(a, fadd(ffma(w[0], z0, fmul(w[2], z2)), ffma(w[1], z1, fmul(w[3], z3))));
    return a;
}

__device__ __forceinline__ void blocked8x4(float4& acc, const float* ww, const float4* z) {
    acc.x = blocked8(acc.x, ww, z[0].x, z[1].x, z[2].x, z[3].x, z[4].x, z[5].x, z[6].x, z[7].x);
    acc.y = blocked8(acc.y, ww, z[0].y, z[1].y, z[2].y, z[3].y, z[4].y, z[5].y, z[6].y, z[7].y);
    acc.z = blocked8(acc.z, ww, z[0].z, z[1].z, z[2].z, z[3].z, z[4].z, z[5].z, z[6].z, z[7].z);
    acc.w = blocked8(acc.w, ww, z[0].w, z[1].w, z[2].w, z[3].w, z[4].w, z[5].w, z[6].w, z[7].w);
}

// the two sub-trees of an 8-block that do not involve the running sum, parked with z6 and z4
__device__ __forceinline__ float4 park8(const float* w, float z0, float z1, float z2, float z3, float z4,
                                        float z5, float z6, float z7) {
    const float x = ffma(w[5], z5, fmul(w[7], z7));
    const float y = fadd(ffma(w[0], z0, fmul(w[2], z2)), ffma(w[1], z1, fmul(w[3], z3)));
    return make_float4(z6, z4, x, y);
}

__device__ __forceinline__ float4 finish_row(const float4& x, const float4& acc, float gamma) {
    float4 out;
    out.x = fadd(x.x, fmul(gamma, acc.x));
    out.y = fadd(x.y, fmul(gamma, acc.y));
    out.z = fadd(x.z, fmul(gamma, acc.z));
    out.w = fadd(x.w, fmul(gamma, acc.w));
    return out;
}

__device__ __forceinline__ float4 absdiff4(const float4& a, const float4& b) {
    return make_float4(fabsf(fsub(a.x, b.x)), fabsf(fsub(a.y, b.y)), fabsf(fsub(a.z, b.z)), fabsf(fsub(a.w, b.w)));
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async16_sa(unsigned smem_addr, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8_sa(unsigned smem_addr, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4_sa(unsigned smem_addr, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Add one row's |delta| (lane L holds columns 4L..4L+3) to the chunk accumulator, in which lane m owns
// cascade lane m: the row is d/32 consecutive cascade rows, taken in order.  The float4-per-lane layout is
// transposed through 512 bytes of the warp's shared memory (one 128-bit store, d/32 32-bit loads).
__device__ __forceinline__ float chunk_add_row(float chunk_acc, const float4& dl, int nseg, int lane, float* scratch) {
    __syncwarp();                                              // every lane has read its X / own piece
    reinterpret_cast<float4*>(scratch)[lane] = dl;
    __syncwarp();
#pragma unroll
    for (int seg = 0; seg < 4; ++seg)
        if (seg < nseg) chunk_acc = fadd(chunk_acc, scratch[seg * 32 + lane]);
    return chunk_acc;
}

// neighbour row piece of this lane: 16 bytes at float4 index `off16` of the lane's column base
__device__ __forceinline__ float4 gather4(const float4* __restrict__ zb, int off16) {
    return __ldg(zb + (unsigned)off16);
}

// start the M gathers of one batch: 128-bit loads straight into registers (lane L loads exactly the float4
// of columns it will reduce)
template <int M>
__device__ __forceinline__ void load_batch(float4 (&buf)[8], const int2* __restrict__ mp, const float4* __restrict__ zb) {
#pragma unroll
    for (int i = 0; i < M; ++i) buf[i] = gather4(zb, mp[i].x);
}

template <int M>
__device__ __forceinline__ void reduce_batch(const float4 (&z)[8], const int2* __restrict__ mp, float4& acc,
                                             bool col_blocked) {
    float w[8];
#pragma unroll
    for (int i = 0; i < M; ++i) w[i] = __int_as_float(mp[i].y);
    if (M == 8 && col_blocked) {
        blocked8x4(acc, w, z);
    } else {
#pragma unroll
        for (int i = 0; i < M; ++i) fma4(w[i], z[i], acc);
    }
}

// Warm L2 with what a later task needs first: its descriptors, the head of its (offset, w) stream, its X
// rows and -- fused L1 -- its own Zcur rows.  Called by the warp that runs `kPrefetchAhead` places earlier
// in the schedule: those cold, streaming reads are otherwise serial DRAM round trips at the start of a task.
constexpr int kPrefetchAhead = 2048;
__device__ __forceinline__ void prefetch_task(const SweepParams& p, int ti, int slab, int lane) {
    if (ti >= p.n_tasks) return;
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(p.tasks + ti));
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(p.tasks + ti) + 1);
    const int desc_first = t0.x, nb = t0.y, e_first = t0.z, e_total = t0.w, r0 = t1.x, flags = t1.y;
    if (lane < 4) {   // first 128 edges of the stream
        if (lane * 32 < e_total) { prefetch_l2(p.coloff + e_first + lane * 32); prefetch_l2(p.w + e_first + lane * 32); }
    } else if (lane < 8) {
        if ((lane - 4) * 32 < nb) prefetch_l2(p.descs + desc_first + (lane - 4) * 32);
    }
    if (flags & kTaskSegment) return;
    // rows: lane -> (row lane / 4, 128-byte line lane % 4) of the slab
    const int nrows = flags & 0xff;
    const int line = lane & 3;
    if (slab * 128 + line * 32 < p.ld)
        for (int r = lane >> 2; r < nrows; r += 8) {
            const size_t off = (size_t)(r0 + r) * p.ld + slab * 128 + line * 32;
            prefetch_l2(p.X + off);
            if (flags & kTaskDirect) prefetch_l2(p.Zc + off);
        }
}

// ------------------------------------------------------------------------------------------
// row kernel: one warp per (task, 128-column slab)
// ------------------------------------------------------------------------------------------
// Per-warp state of the two streams a task reads: batch descriptors (a 32-entry window in registers, read
// with a shuffle) and (offset, w) pairs (32-edge windows published to a 128-entry shared-memory ring).
struct Streams {
    const int32_t* dp;      // descriptors of the task
    const int* offp;        // col * ld / 4 of the task's edges
    const float* wp;
    int nb, e_total;
    int dwin, dnext;        // descriptor windows: current, prefetched
    int pc;                 // (offset, w) window held in registers, not yet published
    float pw;
    int win_q;              // (offset, w) windows published so far
};

// descriptor of batch ib; afterwards the (offset, w) ring holds the batch's edges
__device__ __forceinline__ int next_desc(Streams& s, int ib, int lane, int2* meta) {
    if ((ib & 31) == 0 && ib > 0) {
        s.dwin = s.dnext;
        s.dnext = ib + 32 + lane < s.nb ? __ldg(s.dp + ib + 32 + lane) : 0;
    }
    const int id = __shfl_sync(kFull, s.dwin, ib & 31);
    if (id & kDescPub) {                               // publish the fetched window, fetch the next
        const int base = (s.win_q & 3) * 32;
        const int2 v = make_int2(s.pc, __float_as_int(s.pw));
        meta[base + lane] = v;
        if (base == 0 && lane < 8) meta[kMetaRing + lane] = v;   // mirror: a batch never wraps
        __syncwarp();
        const int off = (++s.win_q) * 32 + lane;
        if (off < s.e_total) { s.pc = __ldg(s.offp + off); s.pw = __ldg(s.wp + off); }
    }
    return id;
}

struct RowCtx {
    const float4* zb;       // Zcur + this lane's columns
    const float* xb;        // X + this lane's columns
    float* znb;             // Znext + this lane's columns
    int ld;
    int r0;
    float gamma;
    bool active, col_blocked, direct;
    int nseg, lane;
    float* scratch;         // 512 bytes of the warp's shared memory (fused L1 transpose)
};

// loads of one batch into a register buffer (+ the row's X and own Zcur pieces if it is the row's last)
__device__ __forceinline__ void issue_loads(const RowCtx& c, int id, const int2* meta, float4 (&buf)[8], float4& xs,
                                            float4& own) {
    const int2* mp = meta + ((id >> kDescMetaShift) & 127);
    switch (id & 15) {
        case 8: load_batch<8>(buf, mp, c.zb); break;
        case 7: load_batch<7>(buf, mp, c.zb); break;
        case 6: load_batch<6>(buf, mp, c.zb); break;
        case 5: load_batch<5>(buf, mp, c.zb); break;
        case 4: load_batch<4>(buf, mp, c.zb); break;
        case 3: load_batch<3>(buf, mp, c.zb); break;
        case 2: load_batch<2>(buf, mp, c.zb); break;
        default: load_batch<1>(buf, mp, c.zb); break;
    }
    if (id & kDescLast) {
        const size_t row_off = (size_t)(c.r0 + ((id >> kDescRowShift) & 31)) * c.ld;
        xs = ld_stream4(c.xb + row_off);
        if (c.direct) own = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(c.zb) + row_off));
    }
}

__device__ __forceinline__ void consume_batch(const RowCtx& c, int id, const int2* meta, const float4 (&buf)[8],
                                              const float4& xs, const float4& own, float4& acc, float& chunk_acc) {
    const int2* mp = meta + ((id >> kDescMetaShift) & 127);
    switch (id & 15) {
        case 8: reduce_batch<8>(buf, mp, acc, c.col_blocked); break;
        case 7: reduce_batch<7>(buf, mp, acc, c.col_blocked); break;
        case 6: reduce_batch<6>(buf, mp, acc, c.col_blocked); break;
        case 5: reduce_batch<5>(buf, mp, acc, c.col_blocked); break;
        case 4: reduce_batch<4>(buf, mp, acc, c.col_blocked); break;
        case 3: reduce_batch<3>(buf, mp, acc, c.col_blocked); break;
        case 2: reduce_batch<2>(buf, mp, acc, c.col_blocked); break;
        default: reduce_batch<1>(buf, mp, acc, c.col_blocked); break;
    }
    if (id & kDescLast) {
        const float4 out = finish_row(xs, acc, c.gamma);
        const size_t row_off = (size_t)(c.r0 + ((id >> kDescRowShift) & 31)) * c.ld;
        if (c.active) *reinterpret_cast<float4*>(c.znb + row_off) = out;
        if (c.direct) {
            const float4 dl = c.active ? absdiff4(out, own) : make_float4(0.f, 0.f, 0.f, 0.f);
            chunk_acc = chunk_add_row(chunk_acc, dl, c.nseg, c.lane, c.scratch);
        }
        acc = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// span task: double-buffered batches -- the loads of batch b + 1 are in flight while batch b is reduced
__device__ __forceinline__ void run_span(const SweepParams& p, const int4 t0, const int4 t1, int slab, int lane,
                                         int2* meta, float* scratch) {
    const int c0 = slab * 128 + lane * 4;
    RowCtx c;
    c.active = c0 < p.ld;
    c.col_blocked = c0 < p.limit;
    const int cc = c.active ? c0 : 0;          // idle lanes shadow lane 0 (same sectors: no extra traffic)
    c.zb = reinterpret_cast<const float4*>(p.Zc + cc);
    c.xb = p.X + cc;
    c.znb = p.Zn + cc;
    c.ld = p.ld;
    c.r0 = t1.x;
    c.gamma = p.gamma;
    c.direct = (t1.y & kTaskDirect) != 0;
    c.nseg = p.d >> 5;
    c.lane = lane;
    c.scratch = scratch;

    Streams s;
    s.dp = p.descs + t0.x; s.nb = t0.y; s.offp = p.coloff + t0.z; s.wp = p.w + t0.z; s.e_total = t0.w;
    // first windows: 32 descriptors, 32 (offset, w) pairs -- one coalesced round trip
    s.dwin = lane < s.nb ? __ldg(s.dp + lane) : 0;
    s.pc = 0; s.pw = 0.0f;
    if (lane < s.e_total) { s.pc = __ldg(s.offp + lane); s.pw = __ldg(s.wp + lane); }
    s.dnext = 32 + lane < s.nb ? __ldg(s.dp + 32 + lane) : 0;
    s.win_q = 0;
    // the rest of the streams: one L2 prefetch per 128-byte line now, so that the window loads further
    // down are L2 hits instead of DRAM round trips on the warp's critical path
    for (int i = 32 + lane * 32; i < s.e_total; i += 32 * 32) { prefetch_l2(s.offp + i); prefetch_l2(s.wp + i); }
    for (int i = 64 + lane * 32; i < s.nb; i += 32 * 32) prefetch_l2(s.dp + i);

    float4 A[8], B[8], xa, xb, oa, ob;
    xa = xb = oa = ob = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float chunk_acc = 0.0f;
    const int nb = s.nb;
    int ida = next_desc(s, 0, lane, meta), idb = 0;
    issue_loads(c, ida, meta, A, xa, oa);
    for (int cb = 0;;) {
        if (cb + 1 < nb) { idb = next_desc(s, cb + 1, lane, meta); issue_loads(c, idb, meta, B, xb, ob); }
        consume_batch(c, ida, meta, A, xa, oa, acc, chunk_acc);
        if (++cb >= nb) break;
        if (cb + 1 < nb) { ida = next_desc(s, cb + 1, lane, meta); issue_loads(c, ida, meta, A, xa, oa); }
        consume_batch(c, idb, meta, B, xb, ob, acc, chunk_acc);
        if (++cb >= nb) break;
    }
    // the span is one whole level-0 chunk: its rows were added in order, skipped rows count +0
    if (c.direct && p.fuse) p.P0[(size_t)((c.r0 - p.row_lo) / p.G) * 32 + lane] = chunk_acc;
}

// hub segment task: up to 16 full 8-blocks of one hub row; park {z6, z4, X, Y} per (block, column), or the
// raw z in the sequential regime, for k_hub_chain
__device__ __forceinline__ void run_segment(const SweepParams& p, const int4 t0, const int4 t1, int slab, int lane,
                                            int2* meta) {
    const int c0 = slab * 128 + lane * 4;
    const bool active = c0 < p.ld, col_blocked = c0 < p.limit;
    const int cc = active ? c0 : 0;
    const float4* zb = reinterpret_cast<const float4*>(p.Zc + cc);
    const int nb = t0.y, b_first = t1.x, nblk_row = t1.w;
    const size_t B0 = (size_t)t1.z;
    // lane's four columns c0..c0+3 sit in 32-column slab c0 / 32 of the row's scratch: [slab][block][32]
    float4* sdst = p.hubS + B0 * p.sld + ((size_t)(cc >> 5) * nblk_row + b_first) * 32 + (cc & 31);
    float4* tdst = p.hubT + (B0 + b_first) * 8 * p.ntail4 + (col_blocked ? 0 : (cc - p.limit) >> 2);
    float2* wdst = p.hubW + B0 + b_first;
    const int* __restrict__ offp = p.coloff + t0.z;
    const float* __restrict__ wp = p.w + t0.z;
    // a segment is at most 128 edges: the whole (offset, w) stream fits the ring
    for (int i = lane; i < t0.w; i += 32) meta[i] = make_int2(__ldg(offp + i), __float_as_int(__ldg(wp + i)));
    __syncwarp();
    float4 A[8], B[8];
    load_batch<8>(A, meta, zb);
    auto park = [&](const float4 (&z)[8], int b) {
        const int2* mp = meta + b * 8;
        float w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = __int_as_float(mp[i].y);
        if (active) {
            if (col_blocked) {
                float4* o = sdst + (size_t)b * 32;
                o[0] = park8(w, z[0].x, z[1].x, z[2].x, z[3].x, z[4].x, z[5].x, z[6].x, z[7].x);
                o[1] = park8(w, z[0].y, z[1].y, z[2].y, z[3].y, z[4].y, z[5].y, z[6].y, z[7].y);
                o[2] = park8(w, z[0].z, z[1].z, z[2].z, z[3].z, z[4].z, z[5].z, z[6].z, z[7].z);
                o[3] = park8(w, z[0].w, z[1].w, z[2].w, z[3].w, z[4].w, z[5].w, z[6].w, z[7].w);
            } else {
                float4* o = tdst + (size_t)b * 8 * p.ntail4;
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i * p.ntail4] = z[i];
            }
        }
        if (slab == 0 && lane == 0) wdst[b] = make_float2(w[4], w[6]);
    };
    for (int cb = 0;;) {
        if (cb + 1 < nb) load_batch<8>(B, meta + (cb + 1) * 8, zb);
        park(A, cb);
        if (++cb >= nb) break;
        if (cb + 1 < nb) load_batch<8>(A, meta + (cb + 1) * 8, zb);
        park(B, cb);
        if (++cb >= nb) break;
    }
}

__global__ void __launch_bounds__(kRowThreads, CLANE_ROW_OCC) k_sweep_rows(SweepParams p) {
    __shared__ __align__(16) unsigned char smem[kRowSmemBytes];
    if (p.st != nullptr && p.st->stop) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* mine = smem + (size_t)warp * kRowWarpSmem;
    int2* meta = reinterpret_cast<int2*>(mine);
    float* scratch = reinterpret_cast<float*>(meta + kMetaRing + 8);
    const int wtask = blockIdx.x * kRowWarps + warp;
    int ti = wtask, slab = 0;
    if (p.nslab > 1) { ti = wtask / p.nslab; slab = wtask - ti * p.nslab; }
    if (ti >= p.n_tasks) return;
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(p.tasks + ti));
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(p.tasks + ti) + 1);
    prefetch_task(p, ti + kPrefetchAhead, slab, lane);
    if (t1.y & kTaskSegment) run_segment(p, t0, t1, slab, lane, meta);
    else run_span(p, t0, t1, slab, lane, meta, scratch);
}

// ------------------------------------------------------------------------------------------
// hub chain: one warp per (hub row, 32 columns below `limit`), plus one warp per hub row for the
// sequential-regime columns.  Runs after k_sweep_rows (same stream).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_hub_chain(SweepParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    if (p.st != nullptr && p.st->stop) return;
    const int lane = threadIdx.x;
    const int per = p.nslab32b + (p.ntail4 > 0 ? 1 : 0);
    const int hr = blockIdx.x / per, s = blockIdx.x - hr * per;
    const int row = __ldg(p.hub_rows + hr);
    const int a = __ldg(p.rowptr + row), k = __ldg(p.rowptr + row + 1) - a;
    const int nblk = k >> 3;
    const size_t B0 = (size_t)__ldg(p.hub_blk0 + hr);
    const int nleft = k - nblk * 8;
    float acc = 0.0f;
    int col;
    bool act;
    if (s < p.nslab32b) {
        // ---- 8-block order: a = fma(w6,z6,a); a = fma(w4,z4,a); a += X; a += Y per block ----
        col = s * 32 + lane;
        act = col < p.limit;
        const int ccol = act ? col : 0;
        float4* ringS = reinterpret_cast<float4*>(smem);                                  // [groups][16][32]
        float2* wq = reinterpret_cast<float2*>(ringS + kChainGroups * kChainGroup * 32);  // [groups][16]
        const unsigned ring_sa = smem_u32(ringS) + lane * 16;
        const unsigned wq_sa = smem_u32(wq) + (lane & 15) * 8;
        const float4* src = p.hubS + B0 * p.sld + (size_t)s * nblk * 32 + lane;   // contiguous 512 B per block
        const float2* wsrc = p.hubW + B0;
        const int ngroups = (nblk + kChainGroup - 1) / kChainGroup;
        auto issue = [&](int g) {
            if (g < ngroups) {
                const int b0 = g * kChainGroup;
                const unsigned dst = ring_sa + (unsigned)(g % kChainGroups) * (kChainGroup * 512);
#pragma unroll
                for (int j = 0; j < kChainGroup; ++j)
                    if (b0 + j < nblk) cp_async16_sa(dst + j * 512, src + (size_t)(b0 + j) * 32);
                if (lane < kChainGroup && b0 + lane < nblk)
                    cp_async8_sa(wq_sa + (unsigned)(g % kChainGroups) * (kChainGroup * 8), wsrc + b0 + lane);
            }
            cp_async_commit();
        };
#pragma unroll
        for (int g = 0; g < kChainGroups - 1; ++g) issue(g);
        // the k mod 8 leftovers: gathered now, added after the blocks
        float lz[7], lw[7];
#pragma unroll
        for (int o = 0; o < 7; ++o) {
            lz[o] = 0.0f; lw[o] = 0.0f;
            if (o < nleft) {
                lw[o] = __ldg(p.w + a + nblk * 8 + o);
                lz[o] = __ldg(p.Zc + (size_t)__ldg(p.coloff + a + nblk * 8 + o) * 4 + ccol);
            }
        }
        for (int g = 0; g < ngroups; ++g) {
            issue(g + kChainGroups - 1);     // its slot held group g - 1: consumed (program order + syncwarp below)
            cp_async_wait<kChainGroups - 1>();
            __syncwarp();                    // wq of group g visible to all lanes
            const float4* rs = ringS + (size_t)(g % kChainGroups) * kChainGroup * 32 + lane;
            const float2* ws = wq + (g % kChainGroups) * kChainGroup;
            const int cnt = min(kChainGroup, nblk - g * kChainGroup);
            if (cnt == kChainGroup) {
                float4 v[kChainGroup];
                float2 wv[kChainGroup];
#pragma unroll
                for (int j = 0; j < kChainGroup; ++j) { v[j] = rs[j * 32]; wv[j] = ws[j]; }
#pragma unroll
                for (int j = 0; j < kChainGroup; ++j) {
                    acc = ffma(wv[j].y, v[j].x, acc);
                    acc = ffma(wv[j].x, v[j].y, acc);
                    acc = fadd(acc, v[j].z);
                    acc = fadd(acc, v[j].w);
                }
            } else {
                for (int j = 0; j < cnt; ++j) {
                    const float4 v = rs[j * 32];
                    const float2 wv = ws[j];
                    acc = ffma(wv.y, v.x, acc);
                    acc = ffma(wv.x, v.y, acc);
                    acc = fadd(acc, v.z);
                    acc = fadd(acc, v.w);
                }
            }
            __syncwarp();                    // all lanes done with wq of this group before it is refilled
        }
        cp_async_wait<0>();
#pragma unroll
        for (int o = 0; o < 7; ++o)
            if (o < nleft) acc = ffma(lw[o], lz[o], acc);
    } else {
        // ---- sequential regime: a = fma(w_i, z_i, a) over all neighbours ----
        const int nt = p.ntail4;
        col = p.limit + lane;
        act = col < p.ld;
        const int ccol = act ? col : p.limit;
        float* zr = reinterpret_cast<float*>(smem);                                  // [groups][32][nt * 4]
        float* wr = zr + (size_t)kTailGroups * kTailGroup * 16;                      // [groups][32]
        const int nnb = nblk * 8;
        const int ngroups = (nnb + kTailGroup - 1) / kTailGroup;
        const float4* src = p.hubT + B0 * 8 * nt;
        auto issue = [&](int g) {
            if (g < ngroups) {
                const int i0 = g * kTailGroup;
                const int cnt = min(kTailGroup, nnb - i0);
                const unsigned zdst = smem_u32(zr + (size_t)(g % kTailGroups) * kTailGroup * 16);
                for (int q = lane; q < cnt * nt; q += 32) cp_async16_sa(zdst + q * 16, src + (size_t)i0 * nt + q);
                if (lane < cnt) cp_async4_sa(smem_u32(wr + (g % kTailGroups) * kTailGroup + lane), p.w + a + i0 + lane);
            }
            cp_async_commit();
        };
#pragma unroll
        for (int g = 0; g < kTailGroups - 1; ++g) issue(g);
        float lz[7], lw[7];
#pragma unroll
        for (int o = 0; o < 7; ++o) {
            lz[o] = 0.0f; lw[o] = 0.0f;
            if (o < nleft) {
                lw[o] = __ldg(p.w + a + nnb + o);
                lz[o] = __ldg(p.Zc + (size_t)__ldg(p.coloff + a + nnb + o) * 4 + ccol);
            }
        }
        const int cl = min(lane, nt * 4 - 1);
        for (int g = 0; g < ngroups; ++g) {
            issue(g + kTailGroups - 1);
            cp_async_wait<kTailGroups - 1>();
            __syncwarp();
            const float* zs = zr + (size_t)(g % kTailGroups) * kTailGroup * 16 + cl;
            const float* ws = wr + (g % kTailGroups) * kTailGroup;
            const int cnt = min(kTailGroup, nnb - g * kTailGroup);
            for (int j = 0; j < cnt; ++j) acc = ffma(ws[j], zs[j * nt * 4], acc);
            __syncwarp();
        }
        cp_async_wait<0>();
#pragma unroll
        for (int o = 0; o < 7; ++o)
            if (o < nleft) acc = ffma(lw[o], lz[o], acc);
    }
    if (act) {
        const size_t off = (size_t)row * p.ld + col;
        p.Zn[off] = fadd(__ldg(p.X + off), fmul(p.gamma, acc));
    }
}

// Fused mode: the level-0 partial of every group that was not swept by a single warp (it holds
// a hub row or was cut into several spans), recomputed from memory.
__global__ void __launch_bounds__(256)
k_fix_chunks(const float* __restrict__ Zn, const float* __restrict__ Zc, int d, int n, int G,
             const int32_t* __restrict__ fix_groups, int n_fix_groups, float* __restrict__ P0,
             const clane_patience* __restrict__ st) {
    if (st != nullptr && st->stop) return;
    const int lane = threadIdx.x & 31;
    const int i = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= n_fix_groups) return;
    const int g = __ldg(fix_groups + i);
    const int r0 = g * G, nrows = min(G, n - r0);
    const int ncr = nrows * (d >> 5);                   // cascade rows of the chunk (ld == d here)
    const size_t base = (size_t)r0 * d + lane;
    // a chunk is at most 32 cascade rows (G*d <= 1024): fetch them all, then add in order -- one round trip
    float v[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const size_t t = base + (size_t)r * 32;
        v[r] = (r < ncr) ? fabsf(fsub(__ldcg(Zn + t), __ldg(Zc + t))) : 0.0f;
    }
    float acc = 0.0f;
#pragma unroll
    for (int r = 0; r < 32; ++r)
        if (r < ncr) acc = fadd(acc, v[r]);
    P0[(size_t)g * 32 + lane] = acc;
}

}  // namespace clane
