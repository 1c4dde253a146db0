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
    const int4* hub_info;       // per hub row (degree-descending): {row, first edge, degree, first scratch block}
    int n_hub_rows;
    int limit;                  // 16 * floor(d / 16): columns below it use the 8-block order
    int ntail4;                 // (ld - limit) / 4: float4 pieces per row in the sequential regime (0..4)
    int nslab32b;               // 32-column slabs below `limit`
    int sld;                    // 32 * nslab32b: columns per block in hubS
    float4* hubS;               // per hub row [32-column slab][block][32] {z6, z4, X, Y}
    float4* hubT;               // per hub row [sequential-regime column][neighbour] raw z (floats)
    int* hub_cnt;               // per hub row: segment warps that have parked their blocks (this sweep)
    int* hub_done;              // per chain CTA: 1 once the early (overlapped) chain pass has produced the row piece
    int hub_first;              // first hub row of this chain launch
    unsigned long long chain_spin_ns;   // early chain pass: give up waiting after this long (the late pass takes over)
    const clane_patience* st;
    // row-partitioned run: the other ranks' Znext buffers (peer memory over NVLink); every finished row is
    // stored to all of them from inside the kernel, so the exchange overlaps the sweep row by row
    float* peer[kMaxPeers];
    int n_remote;
    float* mc;                  // multicast (NVLS) address of Znext: one store reaches every rank; replaces peer[]
};

#ifndef CLANE_ROW_WARPS
#define CLANE_ROW_WARPS 2
#endif
constexpr int kRowWarps = CLANE_ROW_WARPS;     // row kernel: warps (= tasks) per CTA
constexpr int kRowThreads = 32 * kRowWarps;
// row kernel shared memory per warp: (offset, w) ring | 512-byte transpose scratch (fused L1) | 10 row-piece slots
// (8 neighbours + X + own Zcur) of the batch that goes through cp.async
constexpr size_t kRowWarpSmem = (size_t)(kMetaRing + 8) * sizeof(int2) + 512 + 10 * 512;
constexpr size_t kRowSmemBytes = (size_t)kRowWarps * kRowWarpSmem;

// hub chain kernel: one warp per CTA
constexpr int kChainGroup = 16;                // blocks per cp.async group (8 KB)
constexpr int kChainGroups = 12;               // groups in flight (12 x 8 KB = 96 KB: two chain CTAs per SM)
constexpr int kLightStages = 4;                // short hub rows
__host__ __device__ constexpr size_t chain_smem_bytes(int stages) {
    return (size_t)stages * kChainGroup * 32 * sizeof(float4) + (size_t)stages * kChainGroup * sizeof(float2);
}
constexpr size_t kChainSmemBytes = chain_smem_bytes(kChainGroups);
constexpr int kTailGroup = 32;                 // neighbours per stage of the sequential-regime chain
constexpr int kTailPitch = 36;                 // floats per (stage, column): 32 + 4, so that the 16 columns' 128-bit loads spread over the banks
constexpr int kMaxStages = 128;                // mbarrier pairs per chain CTA
constexpr int kChainThreads = 64;              // producer warp + chain warp
static_assert(kChainGroups <= kMaxStages, "mbarriers");

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
__device__ __forceinline__ void cp_async4_sa(unsigned smem_addr, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
// one store, replicated by the NVSwitch into every rank's buffer (multicast mapping of symmetric memory)
__device__ __forceinline__ void multimem_st4(float* mc, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st1(float* mc, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(mc), "f"(v) : "memory");
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// mbarriers of the hub chain's producer / consumer ring
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred done;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 done, [%0], %1;\n\t"
        "@!done bra WAIT_%=;\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

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

// L2 cache-policy hints (evict_last on the Zcur gathers, evict_first on X / Znext) were measured: the sector hit
// rate stays at 44 % either way and the policy descriptors cost 8 % more instructions (R2UR / UMOV), so plain
// accesses are used; Znext is written with the streaming (.cs) qualifier.
// neighbour row piece of this lane: 16 bytes at float4 index `off16` of the lane's column base
__device__ __forceinline__ float4 gather4(const float4* __restrict__ zb, int off16) {
    return __ldg(zb + (unsigned)off16);
}

// Predicated loads in straight-line code.  Inside the pipelined loop every global load is one of these: loads
// issued under divergent control flow (a switch on the batch length, an if on "last") make ptxas wait for the
// outstanding loads at the next control-flow join -- which is the loop's back edge, exactly where the next
// batch's loads must stay in flight.
__device__ __forceinline__ void ldg4_if(float4& v, const float4* p, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w) : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ void ldg4_stream_if(float4& v, const float* p, bool pred) {   // X
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w) : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ void ldg_i32_if(int& v, const int* p, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.global.nc.b32 %0, [%1];\n\t}"
                 : "+r"(v) : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ void ldg_f32_if(float& v, const float* p, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}"
                 : "+f"(v) : "l"(p), "r"((int)pred));
}

// the gathers of one batch: 128-bit loads straight into registers (lane L loads exactly the float4 of columns it
// will reduce)
template <int M>
__device__ __forceinline__ void load_batch(float4 (&buf)[8], const int2* __restrict__ mp, const float4* __restrict__ zb) {
#pragma unroll
    for (int i = 0; i < M; ++i) buf[i] = gather4(zb, mp[i].x);
}

__device__ __forceinline__ void cp_async16_cg(unsigned smem_addr, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
// the gathers of one batch through cp.async: slot i of the warp's shared-memory slots <- neighbour i's row piece
template <int M>
__device__ __forceinline__ void async_batch(unsigned slot_sa, const int2* __restrict__ mp, const float4* __restrict__ zb) {
#pragma unroll
    for (int i = 0; i < M; ++i) cp_async16_cg(slot_sa + i * 512, zb + (unsigned)mp[i].x);
}
template <int M>
__device__ __forceinline__ void slot_batch(float4 (&buf)[8], const float4* __restrict__ myslot) {
#pragma unroll
    for (int i = 0; i < M; ++i) buf[i] = myslot[i * 32];
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

// ------------------------------------------------------------------------------------------
// row kernel: one warp per (task, 128-column slab)
// ------------------------------------------------------------------------------------------
// Per-warp state of the two streams a task reads: batch descriptors (a 32-entry window in registers, read
// with a shuffle) and (offset, w) pairs (32-edge windows published to a 128-entry shared-memory ring).
// Everything is an int offset from a kernel parameter (constant bank), not a pointer: registers are what
// bounds the number of resident warps, and the resident warps are the memory-level parallelism.
struct Streams {
    int desc_first, e_first;
    int nb, e_total;
    int dwin, dnext;        // descriptor windows: current, prefetched
    int pc;                 // the next (offset, w) window to publish, held in registers
    float pw;
    int win_q;              // (offset, w) windows published so far
};

__device__ __forceinline__ void publish_window(int q, int pc, float pw, int lane, int2* meta) {
    const int base = (q & 3) * 32;
    const int2 v = make_int2(pc, __float_as_int(pw));
    meta[base + lane] = v;
    if (base == 0 && lane < 8) meta[kMetaRing + lane] = v;   // mirror: a batch never wraps
}

// Open a task's streams: descriptor windows 0 and 1 in registers; (offset, w) windows 0 and 1 published
// (one coalesced round trip for all of it), window 2 on its way.
__device__ __forceinline__ void open_streams(const SweepParams& p, Streams& s, int lane, int2* meta) {
    const int32_t* dp = p.descs + s.desc_first;
    const int* offp = p.coloff + s.e_first;
    const float* wp = p.w + s.e_first;
    s.dwin = lane < s.nb ? __ldg(dp + lane) : 0;
    int c0 = 0, c1 = 0;
    float w0 = 0.0f, w1 = 0.0f;
    if (lane < s.e_total) { c0 = __ldg(offp + lane); w0 = __ldg(wp + lane); }
    if (32 + lane < s.e_total) { c1 = __ldg(offp + 32 + lane); w1 = __ldg(wp + 32 + lane); }
    s.dnext = 32 + lane < s.nb ? __ldg(dp + 32 + lane) : 0;
    s.pc = 0; s.pw = 0.0f;
    if (64 + lane < s.e_total) { s.pc = __ldg(offp + 64 + lane); s.pw = __ldg(wp + 64 + lane); }
    // the rest of the streams: one L2 prefetch per 128-byte line now, so that the window loads further
    // down are L2 hits
    for (int i = 96 + lane * 32; i < s.e_total; i += 32 * 32) { prefetch_l2(offp + i); prefetch_l2(wp + i); }
    for (int i = 64 + lane * 32; i < s.nb; i += 32 * 32) prefetch_l2(dp + i);
    publish_window(0, c0, w0, lane, meta);
    publish_window(1, c1, w1, lane, meta);
    __syncwarp();
    s.win_q = 2;
}

// descriptor of batch ib; afterwards the (offset, w) ring holds the batch's edges
// kMayRoll = false for odd batch indices: the 32-descriptor window only rolls over at multiples of 32
template <bool kMayRoll>
__device__ __forceinline__ int next_desc(const SweepParams& p, Streams& s, int ib, int lane, int2* meta) {
    if (kMayRoll) {
        const bool roll = (ib & 31) == 0 && ib > 0;
        if (roll) { s.dwin = s.dnext; s.dnext = 0; }
        ldg_i32_if(s.dnext, p.descs + (s.desc_first + ib + 32 + lane), roll && ib + 32 + lane < s.nb);   // int index first: one IMAD.WIDE
    }
    const int id = __shfl_sync(kFull, s.dwin, ib & 31);
    const bool pub = (id & kDescPub) != 0;
    if (pub) {
        // This batch is the first to touch window win_q - 1: publish window win_q (fetched a whole window ago;
        // it replaces window win_q - 4, which the previous batch has left) and fetch the next into the same
        // registers.
        publish_window(s.win_q, s.pc, s.pw, lane, meta);
        __syncwarp();
        ++s.win_q;
    }
    const int off = s.win_q * 32 + lane;
    const bool fetch = pub && off < s.e_total;
    const int eidx = s.e_first + off;
    ldg_i32_if(s.pc, p.coloff + eidx, fetch);
    ldg_f32_if(s.pw, p.w + eidx, fetch);
    return id;
}

// span task: one batch per round -- descriptor, <= 10 loads straight into registers, reduction.  A warp has one
// batch of loads in flight; the memory-level parallelism comes from the resident warps per SM (measured with
// tools/l1pf_probe.cu: deeper per-warp pipelines or L1 / L2 prefetching do not beat more warps).
__device__ __forceinline__ void run_span(const SweepParams& p, const int4 t0, const int4 t1, int slab, int lane,
                                         int2* meta, float* scratch, float4* slots) {
    const int c0 = slab * 128 + lane * 4;
    const bool active = c0 < p.ld;
    const bool col_blocked = c0 < p.limit;
    const int cc = active ? c0 : 0;            // idle lanes shadow lane 0 (same sectors: no extra traffic)
    const int r0 = t1.x;
    const bool direct = (t1.y & kTaskDirect) != 0;   // runtime, not a template: one copy of the code in the instruction cache
    Streams s;
    s.desc_first = t0.x; s.nb = t0.y; s.e_first = t0.z; s.e_total = t0.w;
    open_streams(p, s, lane, meta);

    const float4* zb = reinterpret_cast<const float4*>(p.Zc + cc);
    const unsigned slot_sa = smem_u32(slots) + lane * 16;      // this lane's 16 bytes of slot 0
    const float4* myslot = slots + lane;
    float4 A[8], xs, own;
    xs = own = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float chunk_acc = 0.0f;
    const int nb = s.nb;

    auto finish = [&](const float4& xv, const float4& ov, int row_off) {
        const float4 out = finish_row(xv, acc, p.gamma);
        if (active) {
            st_stream4(p.Zn + row_off, out);
            if (p.mc != nullptr) multimem_st4(p.mc + row_off, out);
            for (int j = 0; j < p.n_remote; ++j) *reinterpret_cast<float4*>(p.peer[j] + row_off) = out;
        }
        if (direct) {
            const float4 dl = active ? absdiff4(out, ov) : make_float4(0.f, 0.f, 0.f, 0.f);
            chunk_acc = chunk_add_row(chunk_acc, dl, p.d >> 5, lane, scratch);
        }
        acc = make_float4(0.f, 0.f, 0.f, 0.f);
    };

    // Two batches in flight per warp, on two independent completion mechanisms: the odd batch goes through
    // cp.async into the warp's 10 shared-memory slots (completion = async group), the even batch straight into
    // registers (completion = scoreboard).  Both are issued before either is reduced; the reductions then run in
    // order.  (Two register buffers would cost 32 more registers, i.e. a fifth of the resident warps; and ptxas
    // tracks all 128-bit loads of a warp on one scoreboard, so waiting for the first would wait for both anyway.)
    for (int cb = 0; cb < nb; cb += 2) {
        const int idr = next_desc<true>(p, s, cb, lane, meta);
        const bool has_s = cb + 1 < nb;
        const int ids = has_s ? next_desc<false>(p, s, cb + 1, lane, meta) : 0;
        // ---- odd batch: cp.async ----
        const int2* mps = meta + ((ids >> kDescMetaShift) & 127);
        const int row_off_s = (r0 + ((ids >> kDescRowShift) & 31)) * p.ld + cc;   // n * ld < 2^31 (clane_plan_create)
        const bool last_s = (ids & kDescLast) != 0;
        switch (ids & 15) {
            case 8: async_batch<8>(slot_sa, mps, zb); break;
            case 7: async_batch<7>(slot_sa, mps, zb); break;
            case 6: async_batch<6>(slot_sa, mps, zb); break;
            case 5: async_batch<5>(slot_sa, mps, zb); break;
            case 4: async_batch<4>(slot_sa, mps, zb); break;
            case 3: async_batch<3>(slot_sa, mps, zb); break;
            case 2: async_batch<2>(slot_sa, mps, zb); break;
            case 1: async_batch<1>(slot_sa, mps, zb); break;
            default: break;
        }
        if (last_s) {
            cp_async16_cg(slot_sa + 8 * 512, p.X + row_off_s);
            if (direct) cp_async16_cg(slot_sa + 9 * 512, p.Zc + row_off_s);
        }
        cp_async_commit();
        // ---- even batch: registers; loads and reduction in the same switch case ----
        const int2* mp = meta + ((idr >> kDescMetaShift) & 127);
        const bool last = (idr & kDescLast) != 0;
        const int row_off = (r0 + ((idr >> kDescRowShift) & 31)) * p.ld + cc;
        // the row's X and own Zcur pieces ride with its last batch (predicated, straight-line: a branch here
        // would make ptxas wait for them at the join, before the gathers are even issued)
        ldg4_stream_if(xs, p.X + row_off, last);
        ldg4_if(own, reinterpret_cast<const float4*>(p.Zc + row_off), last && direct);
        switch (idr & 15) {
            case 8: load_batch<8>(A, mp, zb); reduce_batch<8>(A, mp, acc, col_blocked); break;
            case 7: load_batch<7>(A, mp, zb); reduce_batch<7>(A, mp, acc, col_blocked); break;
            case 6: load_batch<6>(A, mp, zb); reduce_batch<6>(A, mp, acc, col_blocked); break;
            case 5: load_batch<5>(A, mp, zb); reduce_batch<5>(A, mp, acc, col_blocked); break;
            case 4: load_batch<4>(A, mp, zb); reduce_batch<4>(A, mp, acc, col_blocked); break;
            case 3: load_batch<3>(A, mp, zb); reduce_batch<3>(A, mp, acc, col_blocked); break;
            case 2: load_batch<2>(A, mp, zb); reduce_batch<2>(A, mp, acc, col_blocked); break;
            default: load_batch<1>(A, mp, zb); reduce_batch<1>(A, mp, acc, col_blocked); break;
        }
        if (last) finish(xs, own, row_off);
        // ---- reduce the odd batch out of shared memory (each lane reads the 16 bytes it copied: no barrier) ----
        cp_async_wait<0>();
        if (has_s) {
            switch (ids & 15) {
                case 8: slot_batch<8>(A, myslot); reduce_batch<8>(A, mps, acc, col_blocked); break;
                case 7: slot_batch<7>(A, myslot); reduce_batch<7>(A, mps, acc, col_blocked); break;
                case 6: slot_batch<6>(A, myslot); reduce_batch<6>(A, mps, acc, col_blocked); break;
                case 5: slot_batch<5>(A, myslot); reduce_batch<5>(A, mps, acc, col_blocked); break;
                case 4: slot_batch<4>(A, myslot); reduce_batch<4>(A, mps, acc, col_blocked); break;
                case 3: slot_batch<3>(A, myslot); reduce_batch<3>(A, mps, acc, col_blocked); break;
                case 2: slot_batch<2>(A, myslot); reduce_batch<2>(A, mps, acc, col_blocked); break;
                default: slot_batch<1>(A, myslot); reduce_batch<1>(A, mps, acc, col_blocked); break;
            }
            if (last_s) finish(myslot[8 * 32], myslot[9 * 32], row_off_s);
        }
    }
    // the span is one whole level-0 chunk: its rows were added in order, skipped rows count +0
    if (direct && p.fuse) p.P0[(size_t)((r0 - p.row_lo) / p.G) * 32 + lane] = chunk_acc;
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
    // sequential-regime scratch of the row: [column][8 * nblk_row] floats; this lane's first column, this segment
    float* tdst = reinterpret_cast<float*>(p.hubT) + B0 * 32 * p.ntail4 +
                  (size_t)(col_blocked ? 0 : cc - p.limit) * nblk_row * 8 + (size_t)b_first * 8;
    const int* __restrict__ offp = p.coloff + t0.z;
    const float* __restrict__ wp = p.w + t0.z;
    // a segment is at most 128 edges: the whole (offset, w) stream fits the ring
    const int slab_bytes = min(128, p.ld - slab * 128) * 4;
    for (int i = lane; i < t0.w; i += 32) {
        const int off16 = __ldg(offp + i);
        meta[i] = make_int2(off16, __float_as_int(__ldg(wp + i)));
        if (i >= 16) {   // the first two blocks are loaded right away
            const char* a = reinterpret_cast<const char*>(reinterpret_cast<const float4*>(p.Zc + slab * 128) + (unsigned)off16);
            prefetch_l2(a);
            if (slab_bytes > 128) prefetch_l2(a + 128);
            if (slab_bytes > 256) prefetch_l2(a + 256);
            if (slab_bytes > 384) prefetch_l2(a + 384);
        }
    }
    __syncwarp();
    float4 A[8];
    for (int cb = 0; cb < nb; ++cb) {
        load_batch<8>(A, meta + cb * 8, zb);
        const int2* mp = meta + cb * 8;
        float w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = __int_as_float(mp[i].y);
        if (active) {
            if (col_blocked) {
                float4* o = sdst + (size_t)cb * 32;
                o[0] = park8(w, A[0].x, A[1].x, A[2].x, A[3].x, A[4].x, A[5].x, A[6].x, A[7].x);   // kept in L2 for the chain
                o[1] = park8(w, A[0].y, A[1].y, A[2].y, A[3].y, A[4].y, A[5].y, A[6].y, A[7].y);   // kept in L2 for the chain
                o[2] = park8(w, A[0].z, A[1].z, A[2].z, A[3].z, A[4].z, A[5].z, A[6].z, A[7].z);   // kept in L2 for the chain
                o[3] = park8(w, A[0].w, A[1].w, A[2].w, A[3].w, A[4].w, A[5].w, A[6].w, A[7].w);   // kept in L2 for the chain
            } else {
                // sequential regime: every column's values contiguous over the row's neighbours
                float* o = tdst + (size_t)cb * 8;
                const size_t cp = (size_t)nblk_row * 8;
                *reinterpret_cast<float4*>(o) = make_float4(A[0].x, A[1].x, A[2].x, A[3].x);
                *reinterpret_cast<float4*>(o + 4) = make_float4(A[4].x, A[5].x, A[6].x, A[7].x);
                *reinterpret_cast<float4*>(o + cp) = make_float4(A[0].y, A[1].y, A[2].y, A[3].y);
                *reinterpret_cast<float4*>(o + cp + 4) = make_float4(A[4].y, A[5].y, A[6].y, A[7].y);
                *reinterpret_cast<float4*>(o + 2 * cp) = make_float4(A[0].z, A[1].z, A[2].z, A[3].z);
                *reinterpret_cast<float4*>(o + 2 * cp + 4) = make_float4(A[4].z, A[5].z, A[6].z, A[7].z);
                *reinterpret_cast<float4*>(o + 3 * cp) = make_float4(A[0].w, A[1].w, A[2].w, A[3].w);
                *reinterpret_cast<float4*>(o + 3 * cp + 4) = make_float4(A[4].w, A[5].w, A[6].w, A[7].w);
            }
        }
    }
    // tell the row's chain warps (k_hub_chain, running beside this kernel) that these blocks are parked
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicAdd(p.hub_cnt + (t1.y >> kTaskHubShift), 1);
}

__global__ void __launch_bounds__(kRowThreads, CLANE_ROW_OCC * 4 / kRowWarps) k_sweep_rows(SweepParams p) {
    __shared__ __align__(16) unsigned char smem[kRowSmemBytes];
    if (p.st != nullptr && p.st->stop) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* mine = smem + (size_t)warp * kRowWarpSmem;
    int2* meta = reinterpret_cast<int2*>(mine);
    float* scratch = reinterpret_cast<float*>(meta + kMetaRing + 8);
    float4* slots = reinterpret_cast<float4*>(scratch + 128);
    const int wtask = blockIdx.x * kRowWarps + warp;
    int ti = wtask, slab = 0;
    if (p.nslab > 1) { ti = wtask / p.nslab; slab = wtask - ti * p.nslab; }
    if (ti >= p.n_tasks) return;
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(p.tasks + ti));
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(p.tasks + ti) + 1);
    if (t1.y & kTaskSegment) run_segment(p, t0, t1, slab, lane, meta);
    else run_span(p, t0, t1, slab, lane, meta, scratch, slots);
}

// ------------------------------------------------------------------------------------------
// hub chain: one warp per (hub row, 32 columns below `limit`), plus one warp per hub row for the
// sequential-regime columns.  Runs after k_sweep_rows (same stream).
// ------------------------------------------------------------------------------------------
// CTA = two warps: warp 1 copies the row's parked stream into a shared-memory ring (cp.async, completion
// signalled on "full" mbarriers), warp 0 runs the in-order chain out of the ring and hands the stages back
// ("empty" mbarriers) -- the chain warp issues nothing but the loads and the dependent operations of the chain.
//   slab s < nslab32b : 32 columns in the 8-block order, lane = column;  per block {z6, z4, X, Y} + {w4, w6}
//   slab s = nslab32b : the <= 16 sequential-regime columns, lane = column; per neighbour z (transposed by the
//                       segment warps: every column's values are contiguous) + w
// kEarly: launched on a side stream BEFORE k_sweep_rows and running beside it; every CTA waits (bounded) for
//         its row's segment warps, then chains.  The hub rows' segments are the first tasks of the row kernel,
//         so the chains finish long before the ordinary rows do and cost the sweep nothing.
// !kEarly: launched after both; chains whatever the early pass did not (it timed out: kernels serialised by a
//         profiler, or the device too busy to co-schedule), and resets the flags for the next sweep.
// kLight: short rows -- 4-stage ring, no register double buffer: a CTA that costs an SM next to nothing while it
//         waits.  !kLight: long rows (>= kLongBlocks blocks) -- 12 stages, the chain never waits for the ring.
template <bool kEarly, bool kLight>
__global__ void __launch_bounds__(kChainThreads) k_hub_chain(SweepParams p) {
    constexpr int kStages = kLight ? kLightStages : kChainGroups;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned long long full[kMaxStages], empty[kMaxStages];
    __shared__ int s_done;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = p.nslab32b + (p.ntail4 > 0 ? 1 : 0);
    const int hr = p.hub_first + blockIdx.x / per, s = blockIdx.x % per;
    const int cta = hr * per + s;                      // index into hub_done
    const bool stopped = p.st != nullptr && p.st->stop;
    if (kEarly && stopped) return;
    const int4 info = __ldg(p.hub_info + hr);          // one load: every access here queues behind the row kernel's
    const int row = info.x, a = info.y, k = info.z;
    const int nblk = k >> 3;
    if (kEarly) {
        const int expect = ((nblk + kSegEdges / 8 - 1) / (kSegEdges / 8)) * p.nslab;
        const volatile int* cnt = p.hub_cnt + hr;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        t1 = t0;
        bool ready = false;
        for (;;) {
            if (*cnt >= expect) { ready = true; break; }
            __nanosleep(200);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > p.chain_spin_ns) break;
        }
        ready = __syncthreads_and(ready);              // both warps agree
        if (!ready) return;                            // the late pass does it
        __threadfence();                               // acquire: the parked blocks of every segment warp
    } else {
        if (threadIdx.x == 0) {
            s_done = p.hub_done[cta];
            p.hub_done[cta] = 0;                       // also when stopped
            if (s == 0) p.hub_cnt[hr] = 0;
        }
        __syncthreads();
        if (s_done || stopped) return;
    }
    const size_t B0 = (size_t)info.w;
    const int nleft = k - nblk * 8;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(full + i, 32); mbar_init(empty + i, 1); }   // (unrolled: ~0.3 us)
        fence_mbar_init();
    }
    __syncthreads();
    const bool blocked = s < p.nslab32b;
    const int nt = p.ntail4, ntc = p.ld - p.limit;                  // sequential-regime float4 pieces / columns
    const int nnb = nblk * 8;
    const int ngroups = blocked ? (nblk + kChainGroup - 1) / kChainGroup : (nnb + kTailGroup - 1) / kTailGroup;
    float4* ringS = reinterpret_cast<float4*>(smem);                                  // blocked: [stage][16][32]
    float2* wq = reinterpret_cast<float2*>(ringS + kStages * kChainGroup * 32);       //          [stage][16]
    // sequential regime: stage = [ntc columns][kTailPitch] z | [32] w; as many stages as the shared memory holds
    float* zr = reinterpret_cast<float*>(smem);
    const int tstride = ntc * kTailPitch + kTailGroup;
    const int tstages = min(kMaxStages, (int)(chain_smem_bytes(kStages) / sizeof(float)) / max(tstride, 1));

    if (warp == 1) {
        // ---------------- producer ----------------
        if (blocked) {
            const float4* src = p.hubS + B0 * p.sld + (size_t)s * nblk * 32 + lane;   // contiguous 512 B per block
            // {w4, w6} of every block straight from w (constant during a sweep; no scratch copy that another SM's
            // segment warp could be rewriting while this SM's L1 holds a stale sector of it)
            const float* wsrc = p.w + a + 4;
            const unsigned ring_sa = smem_u32(ringS) + lane * 16;
            const unsigned wq_sa = smem_u32(wq) + (lane & 15) * 8;
            for (int g = 0; g < ngroups; ++g) {
                const int st = g % kStages;
                if (g >= kStages) mbar_wait(empty + st, (unsigned)(g / kStages - 1) & 1u);
                const int b0 = g * kChainGroup;
                const unsigned dst = ring_sa + (unsigned)st * (kChainGroup * 512);
#pragma unroll
                for (int j = 0; j < kChainGroup; ++j)
                    if (b0 + j < nblk) cp_async16_sa(dst + j * 512, src + (size_t)(b0 + j) * 32);
                if (lane < kChainGroup && b0 + lane < nblk) {
                    cp_async4_sa(wq_sa + (unsigned)st * (kChainGroup * 8), wsrc + (size_t)(b0 + lane) * 8);
                    cp_async4_sa(wq_sa + (unsigned)st * (kChainGroup * 8) + 4, wsrc + (size_t)(b0 + lane) * 8 + 2);
                }
                cp_async_arrive(full + st);
            }
        } else {
            const float* tsrc = reinterpret_cast<const float*>(p.hubT) + B0 * 32 * nt;   // [column][nnb]
            for (int g = 0; g < ngroups; ++g) {
                const int st = g % tstages;
                if (g >= tstages) mbar_wait(empty + st, (unsigned)(g / tstages - 1) & 1u);
                const int i0 = g * kTailGroup;
                const int cnt = min(kTailGroup, nnb - i0);            // multiple of 8
                // 16-byte pieces: column q / 8, neighbours 4 * (q % 8) .. + 3
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int q = lane + 32 * t, c = q >> 3, pc = (q & 7) * 4;
                    if (c < ntc && pc < cnt)
                        cp_async16_sa(smem_u32(zr + (size_t)st * tstride + c * kTailPitch + pc), tsrc + (size_t)c * nnb + i0 + pc);
                }
                if (lane < cnt) cp_async4_sa(smem_u32(zr + (size_t)st * tstride + ntc * kTailPitch + lane), p.w + a + i0 + lane);
                cp_async_arrive(full + st);
            }
        }
        return;
    }

    // ---------------- consumer: the chain ----------------
    float acc = 0.0f;
    const int col = blocked ? s * 32 + lane : p.limit + lane;
    const bool act = blocked ? col < p.limit : col < p.ld;
    const int ccol = act ? col : (blocked ? 0 : p.limit);
    const float xv = __ldg(p.X + (size_t)row * p.ld + ccol);   // needed last, fetched first
    // the k mod 8 leftovers: gathered now, added after the blocks
    float lz[7], lw[7];
#pragma unroll
    for (int o = 0; o < 7; ++o) {
        lz[o] = 0.0f; lw[o] = 0.0f;
        if (o < nleft) {
            lw[o] = __ldg(p.w + a + nnb + o);
            lz[o] = __ldg(p.Zc + (size_t)__ldg(p.coloff + a + nnb + o) * 4 + ccol);
        }
    }
    if (blocked) {
        // ---- 8-block order: a = fma(w6,z6,a); a = fma(w4,z4,a); a += X; a += Y per block ----
        // Register double buffer: the shared-memory loads of stage g + 1 are issued before the 64 dependent
        // operations of stage g, so the chain itself never waits for the ring.
        float4 v[kChainGroup], vn[kChainGroup];
        float2 wv[kChainGroup], wn[kChainGroup];
        auto fetch = [&](int g, float4 (&dv)[kChainGroup], float2 (&dw)[kChainGroup]) {
            const bool live = g < ngroups;   // past the end: no wait, the loads below return stale data nobody uses
            const int st = g % kStages;      //   (they stay unconditional: loads under a branch would be waited for at its join)
            if (live) mbar_wait(full + st, (unsigned)(g / kStages) & 1u);
            const float4* rs = ringS + (size_t)st * kChainGroup * 32 + lane;
            const float2* ws = wq + st * kChainGroup;
#pragma unroll
            for (int j = 0; j < kChainGroup; ++j) { dv[j] = rs[j * 32]; dw[j] = ws[j]; }   // past the row's end: stale, unused
            __syncwarp();                    // every lane has read the stage
            if (live && lane == 0) mbar_arrive(empty + st);
        };
        auto chain = [&](int g, const float4 (&dv)[kChainGroup], const float2 (&dw)[kChainGroup]) {
            const int cnt = min(kChainGroup, nblk - g * kChainGroup);
            if (cnt == kChainGroup) {
#pragma unroll
                for (int j = 0; j < kChainGroup; ++j) {
                    acc = ffma(dw[j].y, dv[j].x, acc);
                    acc = ffma(dw[j].x, dv[j].y, acc);
                    acc = fadd(acc, dv[j].z);
                    acc = fadd(acc, dv[j].w);
                }
            } else {
#pragma unroll
                for (int j = 0; j < kChainGroup; ++j)
                    if (j < cnt) {
                        acc = ffma(dw[j].y, dv[j].x, acc);
                        acc = ffma(dw[j].x, dv[j].y, acc);
                        acc = fadd(acc, dv[j].z);
                        acc = fadd(acc, dv[j].w);
                    }
            }
        };
        if (kLight) {
            for (int g = 0; g < ngroups; ++g) { fetch(g, v, wv); chain(g, v, wv); }
        } else {
            fetch(0, v, wv);
            for (int g = 0; g < ngroups; g += 2) {
                fetch(g + 1, vn, wn);
                chain(g, v, wv);
                if (g + 1 >= ngroups) break;
                fetch(g + 2, v, wv);
                chain(g + 1, vn, wn);
            }
        }
    } else {
        // ---- sequential regime: a = fma(w_i, z_i, a) over all neighbours (same register double buffer) ----
        const int cl = min(lane, ntc - 1);
        float4 zv[kTailGroup / 4], zn[kTailGroup / 4], wv[kTailGroup / 4], wn[kTailGroup / 4];
        auto fetch = [&](int g, float4 (&dz)[kTailGroup / 4], float4 (&dw)[kTailGroup / 4]) {
            const bool live = g < ngroups;
            const int st = g % tstages;
            if (live) mbar_wait(full + st, (unsigned)(g / tstages) & 1u);
            const float4* zs = reinterpret_cast<const float4*>(zr + (size_t)st * tstride + cl * kTailPitch);
            const float4* ws = reinterpret_cast<const float4*>(zr + (size_t)st * tstride + ntc * kTailPitch);
#pragma unroll
            for (int j = 0; j < kTailGroup / 4; ++j) { dz[j] = zs[j]; dw[j] = ws[j]; }
            __syncwarp();
            if (live && lane == 0) mbar_arrive(empty + st);
        };
        auto chain = [&](int g, const float4 (&dz)[kTailGroup / 4], const float4 (&dw)[kTailGroup / 4]) {
            const int cnt = min(kTailGroup, nnb - g * kTailGroup);   // multiple of 8
#pragma unroll
            for (int j = 0; j < kTailGroup / 4; ++j)
                if (cnt == kTailGroup || 4 * j < cnt) {
                    acc = ffma(dw[j].x, dz[j].x, acc);
                    acc = ffma(dw[j].y, dz[j].y, acc);
                    acc = ffma(dw[j].z, dz[j].z, acc);
                    acc = ffma(dw[j].w, dz[j].w, acc);
                }
        };
        if (kLight) {
            for (int g = 0; g < ngroups; ++g) { fetch(g, zv, wv); chain(g, zv, wv); }
        } else {
            fetch(0, zv, wv);
            for (int g = 0; g < ngroups; g += 2) {
                fetch(g + 1, zn, wn);
                chain(g, zv, wv);
                if (g + 1 >= ngroups) break;
                fetch(g + 2, zv, wv);
                chain(g + 1, zn, wn);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < 7; ++o)
        if (o < nleft) acc = ffma(lw[o], lz[o], acc);
    if (act) {
        const size_t off = (size_t)row * p.ld + col;
        const float v = fadd(xv, fmul(p.gamma, acc));
        p.Zn[off] = v;
        if (p.mc != nullptr) multimem_st1(p.mc + off, v);
        for (int j = 0; j < p.n_remote; ++j) p.peer[j][off] = v;
    }
    if (kEarly && lane == 0) p.hub_done[cta] = 1;
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
