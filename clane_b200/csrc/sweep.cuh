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
// Scheduling unit: a GROUP of G consecutive rows.  When d is 32/64/128, G*d is exactly one
// level-0 chunk of the ATen cascade sum, so the warp (or CTA) that owns a group also produces
// that chunk's 32-lane partial of sum|Znext - Zcur| in the reference's order (fused L1).
//
//   row role : one warp per (span, 128-column slab); a span is a run of rows of one group with
//              a bounded edge count.  Per "batch" (one 8-neighbour block of one row): col/w from
//              a per-warp shared-memory ring, 8 x LDG.128 gathers per lane, in-order reduction;
//              the X row and the own Zcur row ride with the row's last batch.
//   hub role : one CTA per (row of degree > hub_threshold, 32-column slab).  The neighbour rows'
//              128-byte slab pieces are streamed through a 16-stage cp.async ring (32 neighbours
//              per stage, 60 KB in flight) by all 8 warps; warp 0 (lane = column) runs the
//              in-order chain out of shared memory.  A hub row is thus spread over d/32 SMs.
//              Row-role warps skip hub rows; in fused mode the level-0 partial of a group that
//              holds a hub row (or was cut into several spans) is recomputed by k_fix_chunks.
#pragma once
#include "common.cuh"

// build-time tuning knobs of the row kernel: ring slots per warp, resident CTAs per SM
#ifndef CLANE_RING
#define CLANE_RING 16
#endif
#ifndef CLANE_ROW_OCC
#define CLANE_ROW_OCC 6
#endif

namespace clane {

struct SweepParams {
    const float* X;
    const float* Zc;
    float* Zn;
    int ld, d, n;
    const int32_t* rowptr;
    const int32_t* coloff;      // col[e] * ld: element offset of the neighbour's row
    const float* w;
    float gamma;
    const int32_t* hub_rows;    // rows of degree > hub_threshold, degree-descending
    int n_hub_rows;
    int nslab32;                // 32-column slabs per row (hub role)
    const int32_t* span_row;    // first row of each span; spans sorted by edge count, descending
    const int32_t* span_meta;   // rows in the span | (1 << 8 if the span is a whole fused chunk)
    const int2* span_edges;     // (first edge, edge count) of the span: the first window is fetched with rowptr
    int n_spans;
    int row_lo, row_hi;         // rows covered by the plan; groups are cut from row_lo
    int G;                      // rows per group (<= 32)
    int nslab;                  // 128-column slabs per row
    int fuse;                   // 1: write one 32-lane |delta| partial per group to P0
    float* P0;
    int hub_threshold;
    const clane_patience* st;
};

constexpr int kRowThreads = 128;               // row kernel: 4 warps per CTA, 5 CTAs per SM at <= 102 registers
constexpr int kRowWarps = 4;
constexpr int kHubThreads = 416;               // hub kernel: 13 warps per CTA (chain, 4 pre-reduce, 8 copy)
constexpr int kMetaRing = 128;                 // (offset, w) pairs per warp, + 8 mirrored entries
constexpr int kMetaSlots = kMetaRing + 8 + 4;   // + 8-entry batch descriptor queue (8 ints = 4 int2)
constexpr int kHubStage = 64;                  // neighbours per ring stage
constexpr int kHubStages = 8;                  // 8 x 64 x 128 B = 64 KB
constexpr int kHubMeta = 4;                    // col / w are fetched this many stages ahead of the copies
constexpr int kHubRingFloats = kHubStages * kHubStage * 32;
// row kernel shared memory per warp: (offset, w) ring | 32 x 512-byte row-piece ring
constexpr size_t kRowWarpSmem = (size_t)kMetaSlots * sizeof(int2) + (size_t)CLANE_RING * 32 * sizeof(float4);
constexpr size_t kRowSmemBytes = (size_t)kRowWarps * kRowWarpSmem;
// hub kernel shared memory: copy ring | w ring
constexpr size_t kHubSmemUsed = (size_t)kHubRingFloats * sizeof(float) + (size_t)kHubStages * kHubStage * sizeof(float) +
                                (size_t)2 * 8 * 32 * sizeof(float4);   // copy ring | w ring | {z6,z4,X,Y} of two stages
constexpr size_t kHubSmemBytes = kHubSmemUsed;   // (asking for the whole SM to keep row CTAs away delays the hub CTAs' start: measured worse)

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

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Add one row's |delta| (lane L holds columns 4L..4L+3) to the chunk accumulator, in which lane m owns
// cascade lane m: the row is d/32 consecutive cascade rows, taken in order.  The float4-per-lane layout is
// transposed through 512 bytes of the warp's shared memory (one 128-bit store, d/32 32-bit loads).
__device__ __forceinline__ float chunk_add_row(float chunk_acc, const float4& dl, int nseg, int lane, float* scratch) {
    __syncwarp();                                              // the previous row's reads are done
    reinterpret_cast<float4*>(scratch)[lane] = dl;
    __syncwarp();
#pragma unroll
    for (int seg = 0; seg < 4; ++seg)
        if (seg < nseg) chunk_acc = fadd(chunk_acc, scratch[seg * 32 + lane]);
    return chunk_acc;
}

// ------------------------------------------------------------------------------------------
// row role: one warp per (span, 128-column slab)
// ------------------------------------------------------------------------------------------
// A span is a run of consecutive rows of one group with a bounded edge count (hub rows are
// skipped).  The warp walks the span's edge stream twice, with two cursors:
//   issue   : per "batch" (one 8-neighbour block of one row) it reads the neighbours' row
//             offsets from its (offset, w) ring and starts one 512-byte cp.async per neighbour
//             (16 bytes per lane: lane L copies exactly the float4 of columns it will reduce,
//             so no barrier is ever needed) into its private ring of row pieces; the
//             row's X piece (and own Zcur piece, fused L1) ride with the row's last batch;
//   consume : waits for the OLDEST batch only (cp.async.wait_group), reduces it in the
//             reference's order, and frees its slots.
// Up to CLANE_RING row pieces (512 B each) per warp are in flight whatever the row lengths: the gathers of
// the next rows overlap the reduction of the current one (decoupled access / execute).
constexpr int kRing = CLANE_RING;      // 512-byte row-piece slots per warp
static_assert(kRing >= 16 && (kRing & (kRing - 1)) == 0, "a batch needs up to 10 slots; slot indices are masked");
constexpr int kMaxPending = 6;         // batches in flight per warp

struct Cursor { int ri, a, k, pos; };

__device__ __forceinline__ void cp_async_wait_pending(int pending) {   // oldest of `pending` groups complete
    switch (pending) {
        case 1: cp_async_wait<0>(); break;
        case 2: cp_async_wait<1>(); break;
        case 3: cp_async_wait<2>(); break;
        case 4: cp_async_wait<3>(); break;
        case 5: cp_async_wait<4>(); break;
        case 6: cp_async_wait<5>(); break;
        case 7: cp_async_wait<6>(); break;
        default: cp_async_wait<7>(); break;
    }
}

__device__ __forceinline__ void cp_async16_sa(unsigned smem_addr, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}

// start the M gathers of one batch: slot (tail + i) of the warp's ring <- 16 bytes of neighbour i's row
template <int M>
__device__ __forceinline__ void issue_batch(unsigned ring_sa, int tail, const int2* __restrict__ mp,
                                            const float* __restrict__ zb) {
#pragma unroll
    for (int i = 0; i < M; ++i) cp_async16_sa(ring_sa + (((tail + i) & (kRing - 1)) << 9), zb + mp[i].x);
}

template <int M>
__device__ __forceinline__ void reduce_batch(const float4* __restrict__ ring, int head, const int2* __restrict__ mp,
                                             int lane, float4& acc, bool col_blocked) {
    float4 z[M];
    float w[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
        z[i] = ring[((head + i) & (kRing - 1)) * 32 + lane];
        w[i] = __int_as_float(mp[i].y);
    }
    if (M == 8 && col_blocked) {
        blocked8x4(acc, w, z);
    } else {
#pragma unroll
        for (int i = 0; i < M; ++i) fma4(w[i], z[i], acc);
    }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Warm L2 with what a span needs first (row pointers, the head of its (offset, w) stream, its X rows
// and -- fused L1 -- its own Zcur rows).  Called by the warp that sweeps a span `kPrefetchAhead`
// places earlier in the schedule: those cold, streaming reads are otherwise three serial DRAM round
// trips at the start of every span.
constexpr int kPrefetchAhead = 2048;
__device__ __forceinline__ void prefetch_span(const SweepParams& p, int64_t si, int slab, int lane) {
    if (si >= p.n_spans) return;
    const int r0 = __ldg(p.span_row + si);
    const int nrows = __ldg(p.span_meta + si) & 0xff;
    const int2 se = __ldg(p.span_edges + si);
    if (lane == 0) { prefetch_l2(p.rowptr + r0); prefetch_l2(p.rowptr + r0 + nrows); }
    if (lane < 4) {   // first 128 edges of the stream
        if (lane * 32 < se.y) { prefetch_l2(p.coloff + se.x + lane * 32); prefetch_l2(p.w + se.x + lane * 32); }
    }
    const int row_lines = (min(128, p.ld - slab * 128) * 4 + 127) >> 7;
    for (int i = lane; i < nrows * row_lines; i += 32) {
        const size_t off = (size_t)(r0 + i / row_lines) * p.ld + slab * 128 + (i % row_lines) * 32;
        prefetch_l2(p.X + off);
        if (p.fuse) prefetch_l2(p.Zc + off);
    }
}

__device__ __forceinline__ void row_span_task(const SweepParams& p, int r0, int nrows, bool direct, int slab,
                                              int lane, int2* meta, float4* ring, int e_first, int e_total) {
    const int c = slab * 128 + lane * 4;
    const bool active = c < p.ld;
    const bool col_blocked = c < (p.d / 16) * 16;
    const int cc = active ? c : 0;
    const float* zb = p.Zc + cc;
    const int nseg = p.d >> 5;
    const int extra = direct ? 2 : 1;      // ring slots a row's last batch adds: X piece (+ own Zcur piece)
    const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring) + lane * 16;   // this lane's 16 bytes of slot 0
    int* dq = reinterpret_cast<int*>(meta + kMetaRing + 8);

    // row pointers of the span: lane i holds [start, end) of row r0 + i
    int rp_a = 0, rp_b = 0;
    if (lane < nrows) { rp_a = __ldg(p.rowptr + r0 + lane); rp_b = __ldg(p.rowptr + r0 + lane + 1); }
    const int* __restrict__ offp = p.coloff + e_first;
    const float* __restrict__ wp = p.w + e_first;
    int pc = 0;
    float pw = 0.0f;
    if (lane < e_total) { pc = __ldg(offp + lane); pw = __ldg(wp + lane); }
    int win_q = 0, filled = 0;   // window held in registers / stream offset published to the meta ring
    // the rest of this span's (offset, w) stream: one L2 prefetch per 128-byte line now, so that the
    // window loads further down are L2 hits instead of DRAM round trips on the warp's critical path
    for (int i = 32 + lane * 32; i < e_total; i += 32 * 32) { prefetch_l2(offp + i); prefetch_l2(wp + i); }

    auto advance = [&](Cursor& cu) -> bool {   // move to the next batch; false when the span is exhausted
        while (cu.pos >= cu.k) {
            if (++cu.ri >= nrows) return false;
            cu.a = __shfl_sync(kFull, rp_a, cu.ri);
            cu.k = __shfl_sync(kFull, rp_b, cu.ri) - cu.a;
            cu.pos = cu.k > p.hub_threshold ? cu.k : 0;   // hub rows have their own kernel; sinks have k = 0
        }
        return true;
    };

    // The consume side does not walk the rows again: every issued batch leaves a descriptor
    //   m | last << 4 | row-in-span << 8 | meta offset << 16
    // in an 8-entry queue of the warp's shared memory (all lanes store the same word, each reads its own).
    Cursor ic;
    ic.ri = -1; ic.a = 0; ic.k = 0; ic.pos = 0;
    bool more = advance(ic);
    int head = 0, tail = 0, used = 0, pending = 0, qh = 0, qt = 0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float chunk_acc = 0.0f;

    for (;;) {
        // ---- issue: as many batches as the ring and the group budget allow ----
        while (more && pending < kMaxPending) {
            const int m = min(8, ic.k - ic.pos);
            const bool last = ic.pos + m >= ic.k;
            const int need = m + (last ? extra : 0);
            if (used + need > kRing) break;
            const int u = ic.a + ic.pos - e_first;          // stream offset of the batch
            while (filled < u + m) {                        // publish the fetched window, fetch the next
                const int base = (win_q & 3) * 32;
                const int2 v = make_int2(pc, __float_as_int(pw));
                meta[base + lane] = v;
                if (base == 0 && lane < 8) meta[128 + lane] = v;   // mirror: a batch never wraps
                __syncwarp();
                filled = (++win_q) * 32;
                const int off = filled + lane;
                if (off < e_total) { pc = __ldg(offp + off); pw = __ldg(wp + off); }
            }
            dq[qt & 7] = m | (last ? 16 : 0) | (ic.ri << 8) | ((u & 127) << 16);
            ++qt;
            if (active) {
                const int2* mp = meta + (u & 127);
                switch (m) {
                    case 8: issue_batch<8>(ring_sa, tail, mp, zb); break;
                    case 7: issue_batch<7>(ring_sa, tail, mp, zb); break;
                    case 6: issue_batch<6>(ring_sa, tail, mp, zb); break;
                    case 5: issue_batch<5>(ring_sa, tail, mp, zb); break;
                    case 4: issue_batch<4>(ring_sa, tail, mp, zb); break;
                    case 3: issue_batch<3>(ring_sa, tail, mp, zb); break;
                    case 2: issue_batch<2>(ring_sa, tail, mp, zb); break;
                    default: issue_batch<1>(ring_sa, tail, mp, zb); break;
                }
                if (last) {
                    const size_t row_off = (size_t)(r0 + ic.ri) * p.ld + cc;
                    cp_async16_sa(ring_sa + (((tail + m) & (kRing - 1)) << 9), p.X + row_off);
                    if (direct) cp_async16_sa(ring_sa + (((tail + m + 1) & (kRing - 1)) << 9), p.Zc + row_off);
                }
            }
            cp_async_commit();
            tail += need; used += need; ++pending;
            ic.pos += m;
            more = advance(ic);
        }
        if (pending == 0) break;
        // ---- consume the oldest batch ----
        cp_async_wait_pending(pending);
        const int desc = dq[qh & 7];
        ++qh;
        const int m = desc & 15;
        const bool last = (desc & 16) != 0;
        const int2* mp = meta + (desc >> 16);
        switch (m) {
            case 8: reduce_batch<8>(ring, head, mp, lane, acc, col_blocked); break;
            case 7: reduce_batch<7>(ring, head, mp, lane, acc, col_blocked); break;
            case 6: reduce_batch<6>(ring, head, mp, lane, acc, col_blocked); break;
            case 5: reduce_batch<5>(ring, head, mp, lane, acc, col_blocked); break;
            case 4: reduce_batch<4>(ring, head, mp, lane, acc, col_blocked); break;
            case 3: reduce_batch<3>(ring, head, mp, lane, acc, col_blocked); break;
            case 2: reduce_batch<2>(ring, head, mp, lane, acc, col_blocked); break;
            default: reduce_batch<1>(ring, head, mp, lane, acc, col_blocked); break;
        }
        int need = m;
        if (last) {
            float4 dl = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) {
                const float4 xs = ring[((head + m) & (kRing - 1)) * 32 + lane];
                const float4 out = finish_row(xs, acc, p.gamma);
                *reinterpret_cast<float4*>(p.Zn + (size_t)(r0 + ((desc >> 8) & 0xff)) * p.ld + c) = out;
                if (direct) dl = absdiff4(out, ring[((head + m + 1) & (kRing - 1)) * 32 + lane]);
            }
            // transpose scratch: the X slot of this batch (already read; not re-targeted before the next issue)
            if (direct) chunk_acc = chunk_add_row(chunk_acc, dl, nseg, lane, reinterpret_cast<float*>(ring + ((head + m) & (kRing - 1)) * 32));
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
            need += extra;
        }
        head += need; used -= need; --pending;
    }
    // the span is one whole level-0 chunk: its rows were added in order, skipped rows count +0
    if (direct) p.P0[(size_t)((r0 - p.row_lo) / p.G) * 32 + lane] = chunk_acc;
}

// ------------------------------------------------------------------------------------------
// hub role
// ------------------------------------------------------------------------------------------

// One (hub row, 32-column slab) per CTA of 13 warps, three roles, lock-stepped per 32-neighbour stage:
//   warps 5-12 copy: warp 5+c streams neighbours 4c..4c+3 of every stage (4 x 128 B) into the ring with one
//              cp.async per stage; (col*ld, w) are prefetched 8 stages ahead in statically indexed registers;
//   warps 1-4  pre-reduce: warp 1+b computes, for block b of the stage that has just landed, the two
//              sub-trees of the 8-neighbour pattern that do not involve the running sum,
//              X = fma(w5,z5, w7*z7) and Y = fma(w0,z0, w2*z2) + fma(w1,z1, w3*z3), and parks
//              {z6, z4, X, Y} per column as one float4;
//   warp 0     chain: a = fma(w6,z6,a); a = fma(w4,z4,a); a += X; a += Y for the blocks of the previous
//              stage -- 4 dependent operations per 8 neighbours, the minimum the reference's order allows,
//              fed by two 128-bit shared-memory loads per block.
// Columns in the sequential regime (>= 16*floor(d/16)) are reduced by warp 0 alone, straight from the ring.
__device__ void hub_slab_task(const SweepParams& p, int row, int slab32, float* ringf, float* wsm, float4* xy) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a = __ldg(p.rowptr + row), k = __ldg(p.rowptr + row + 1) - a;
    const int nst = (k + kHubStage - 1) / kHubStage;
    const int ccol = slab32 * 32 + lane;                 // reduce view: one column per lane
    const bool blk = ccol < (p.d / 16) * 16;             // k > hub_threshold >= 8
    // copy view (warps 5-12): this lane moves 16 bytes of neighbours nb and nb + 32 of every stage
    const int nb = (warp - 5) * 4 + (lane >> 3);
    const int pcol = slab32 * 32 + (lane & 7) * 4;
    const bool pact = pcol < p.ld;

    int cq[kHubMeta], cq2[kHubMeta];
    float wq[kHubMeta];    // warps 5 and 6 carry the stage's 64 weights (32 each)
#pragma unroll
    for (int i = 0; i < kHubMeta; ++i) { cq[i] = cq2[i] = 0; wq[i] = 0.0f; }
    const int wbase = (warp == 6) ? 32 : 0;
    const bool wcarrier = warp == 5 || warp == 6;
    if (warp >= 5) {
#pragma unroll
        for (int i = 0; i < kHubMeta; ++i) {
            const int i0 = i * kHubStage + nb, wi = i * kHubStage + wbase + lane;
            cq[i] = i0 < k ? __ldg(p.coloff + a + i0) : 0;
            cq2[i] = i0 + 32 < k ? __ldg(p.coloff + a + i0 + 32) : 0;
            wq[i] = (wcarrier && wi < k) ? __ldg(p.w + a + wi) : 0.0f;
        }
    }
    const float* zsrc = p.Zc + pcol;
    float* rdst = ringf + nb * 32 + (lane & 7) * 4;
    auto issue = [&](int si, int& c0, int& c1, float& wslot) {   // copy warps only
        const int slot = si % kHubStages;
        if (pact && si * kHubStage + nb < k) cp_async16(rdst + slot * (kHubStage * 32), zsrc + c0);
        if (pact && si * kHubStage + nb + 32 < k) cp_async16(rdst + slot * (kHubStage * 32) + 32 * 32, zsrc + c1);
        if (wcarrier) wsm[slot * kHubStage + wbase + lane] = wslot;
        cp_async_commit();
        const int i0 = (si + kHubMeta) * kHubStage + nb, wi = (si + kHubMeta) * kHubStage + wbase + lane;
        c0 = i0 < k ? __ldg(p.coloff + a + i0) : 0;
        c1 = i0 + 32 < k ? __ldg(p.coloff + a + i0 + 32) : 0;
        if (wcarrier) wslot = wi < k ? __ldg(p.w + a + wi) : 0.0f;
    };
    // pre-reduce block b of stage s (a full block) -> xy[s & 1][b][lane] = {z6, z4, X, Y}
    auto prereduce = [&](int s, int b) {
        const float* src = ringf + (size_t)(s % kHubStages) * kHubStage * 32 + b * 8 * 32 + lane;
        const float4* w4 = reinterpret_cast<const float4*>(wsm + (s % kHubStages) * kHubStage + b * 8);
        const float4 wa = w4[0], wb = w4[1];
        const float z0 = src[0], z1 = src[32], z2 = src[64], z3 = src[96], z4 = src[128], z5 = src[160],
                    z6 = src[192], z7 = src[224];
        const float x = ffma(wb.y, z5, fmul(wb.w, z7));
        const float y = fadd(ffma(wa.x, z0, fmul(wa.z, z2)), ffma(wa.y, z1, fmul(wa.w, z3)));
        xy[((s & 1) * 8 + b) * 32 + lane] = make_float4(z6, z4, x, y);
    };
    // chain over stage s (its {z6, z4, X, Y} were written during the previous interval)
    auto chain = [&](int s, float& acc) {
        const float* wrow = wsm + (s % kHubStages) * kHubStage;
        const int cnt = min(kHubStage, k - s * kHubStage);
        const int nfull = cnt >> 3;
        const float* src = ringf + (size_t)(s % kHubStages) * kHubStage * 32 + lane;
        if (blk) {
            float4 v[8], wv[8];
#pragma unroll
            for (int b = 0; b < 8; ++b)
                if (b < nfull) {
                    v[b] = xy[((s & 1) * 8 + b) * 32 + lane];
                    wv[b] = *reinterpret_cast<const float4*>(wrow + b * 8 + 4);   // {w4, w5, w6, w7}
                }
#pragma unroll
            for (int b = 0; b < 8; ++b)
                if (b < nfull) {
                    acc = ffma(wv[b].z, v[b].x, acc);
                    acc = ffma(wv[b].x, v[b].y, acc);
                    acc = fadd(acc, v[b].z);
                    acc = fadd(acc, v[b].w);
                }
        } else {
            for (int o = 0; o < nfull * 8; ++o) acc = ffma(wrow[o], src[o * 32], acc);
        }
        for (int o = nfull * 8; o < cnt; ++o) acc = ffma(wrow[o], src[o * 32], acc);   // k mod 8 leftovers
    };

    // prologue: stages 0 .. kAhead-1 in flight (slot = stage % 8).  The chain lags the copies by one more
    // stage than a plain ring would, so only kHubStages - 2 stages may be in flight: interval s overwrites the
    // ring slot of stage s - 2, which the chain finished in interval s - 1.
    constexpr int kAhead = kHubStages - 2;
    if (warp >= 5) {
#pragma unroll
        for (int s = 0; s < kAhead; ++s) {
            if (s < nst) issue(s, cq[s % kHubMeta], cq2[s % kHubMeta], wq[s % kHubMeta]);
            else cp_async_commit();
        }
    }
    float acc = 0.0f;
    // interval s (after barrier s): copy warps issue stage s+14; pre-reduce warps work on stage s; the chain
    // warp consumes stage s-1.  One more interval drains the chain.
    for (int sb = 0; sb <= nst; sb += kHubMeta) {
#pragma unroll
        for (int j = 0; j < kHubMeta; ++j) {
            const int s = sb + j;
            if (s <= nst) {                                   // CTA-uniform
                if (warp >= 5) cp_async_wait<kAhead - 1>();
                __syncthreads();                             // stage s landed; {z6,z4,X,Y} of stage s-1 written
                if (warp >= 5) {
                    const int si = s + kAhead;
                    if (si < nst) issue(si, cq[(j + kAhead) % kHubMeta], cq2[(j + kAhead) % kHubMeta], wq[(j + kAhead) % kHubMeta]);
                    else cp_async_commit();
                } else if (warp >= 1) {     // two blocks per pre-reduce warp
                    const int cnt = min(kHubStage, k - s * kHubStage);
                    if (s < nst && (warp - 1) * 8 + 8 <= cnt) prereduce(s, warp - 1);
                    if (s < nst && (warp + 3) * 8 + 8 <= cnt) prereduce(s, warp + 3);
                } else if (s >= 1) {
                    chain(s - 1, acc);
                }
            }
        }
    }
    cp_async_wait<0>();
    if (warp == 0 && ccol < p.ld) {
        const size_t off = (size_t)row * p.ld + ccol;
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

__global__ void __launch_bounds__(kRowThreads, CLANE_ROW_OCC) k_sweep_rows(SweepParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    if (p.st != nullptr && p.st->stop) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* mine = smem + (size_t)warp * kRowWarpSmem;
    float4* ring = reinterpret_cast<float4*>(mine);
    int2* meta = reinterpret_cast<int2*>(mine + (size_t)CLANE_RING * 32 * sizeof(float4));
    const int64_t task = (int64_t)blockIdx.x * kRowWarps + warp;
    const int64_t si = task / p.nslab;
    if (si >= p.n_spans) return;
    const int smeta = __ldg(p.span_meta + si);
    const int2 se = __ldg(p.span_edges + si);
    prefetch_span(p, si + kPrefetchAhead, (int)(task - si * p.nslab), lane);
    row_span_task(p, __ldg(p.span_row + si), smeta & 0xff, (smeta >> 8) != 0 && p.fuse, (int)(task - si * p.nslab), lane,
                  meta, ring, se.x, se.y);
}

__global__ void __launch_bounds__(kHubThreads, 1) k_sweep_hubs(SweepParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    if (p.st != nullptr && p.st->stop) return;
    float* ringf = reinterpret_cast<float*>(smem);
    float* wsm = ringf + kHubRingFloats;
    float4* xy = reinterpret_cast<float4*>(wsm + kHubStages * kHubStage);
    const int hr = blockIdx.x / p.nslab32;
    hub_slab_task(p, __ldg(p.hub_rows + hr), blockIdx.x - hr * p.nslab32, ringf, wsm, xy);
}

}  // namespace clane
