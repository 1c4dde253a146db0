// ATen cascade sum on the device, bit-exact with torch.sum on one CPU thread
// (reference call sites /root/reference/clane/similarity.py:37 and
// /root/reference/clane/embedder.py:60,94).
//
// The CPU algorithm keeps 32 accumulators keyed by (flat index mod 32) -- exactly one warp --
// and four cascade levels: level 0 adds `step` consecutive 32-element rows sequentially, is
// added into level 1 and cleared; level 1 is dumped into level 2 every step^2 rows, level 2
// into level 3 every step^3 rows.  Every complete level-k node therefore starts from zero and
// is independent of its siblings: nodes are computed in parallel, and only the (short)
// sequences of node sums are added in order.
//
//   k_cascade_l01   : one CTA per level-1 node (step^2 rows); warps own level-0 chunks
//   k_cascade_finish: one CTA; levels 2 and 3, the ragged tails, the final lane combine,
//                     and (optionally) the patience state machine of Embedder.propagate.
#pragma once
#include "common.cuh"

namespace clane {

// An element source maps a flat index of the UNPADDED array to NQ fp32 values.
// prepare(t0) decomposes a chunk base once per warp; load(off) serves t0 + off.

struct ElemAbsDiff {  // |a[t] - b[t]| over an [n, d] matrix stored with leading dimension ld
    static constexpr int NQ = 1;
    const float* a;
    const float* b;
    int d, ld;
    struct Base { int64_t r0; uint32_t j0; };
    __device__ __forceinline__ Base prepare(int64_t t0) const {
        Base bs;
        if (d == ld) { bs.r0 = t0; bs.j0 = 0; }
        else { bs.r0 = t0 / d; bs.j0 = (uint32_t)(t0 - bs.r0 * d); }
        return bs;
    }
    __device__ __forceinline__ void load(const Base& bs, uint32_t off, float* v) const {
        size_t idx;
        if (d == ld) idx = (size_t)bs.r0 + off;
        else {
            uint32_t jj = bs.j0 + off, q = jj / (uint32_t)d;
            idx = (size_t)(bs.r0 + q) * ld + (jj - q * (uint32_t)d);
        }
        v[0] = fabsf(fsub(__ldg(a + idx), __ldg(b + idx)));
    }
    // streaming form for the level-0 loop: element t, then t + 32, t + 64, ... without a division per element
    struct Iter { size_t idx; int j; };
    __device__ __forceinline__ Iter iter(int64_t t) const {
        Iter it;
        if (d == ld) { it.idx = (size_t)t; it.j = 0; }
        else { const int64_t r = t / d; it.j = (int)(t - r * d); it.idx = (size_t)r * ld + it.j; }
        return it;
    }
    __device__ __forceinline__ void next(Iter& it, float* v) const {
        v[0] = fabsf(fsub(__ldg(a + it.idx), __ldg(b + it.idx)));
        it.idx += 32;
        if (d != ld) {
            it.j += 32;
            while (it.j >= d) { it.j -= d; it.idx += (size_t)(ld - d); }
        }
    }
};

struct ElemGatherSq2 {  // (Z[erow[e]][j]^2, Z[col[e]][j]^2) for flat t = e*d + j
    static constexpr int NQ = 2;
    const float* Z;
    const int32_t* erow;
    const int32_t* col;
    int d, ld;
    int64_t e_count;      // edges (the streaming form never reads an index past the last one)
    struct Base { int64_t e0; uint32_t j0; };
    __device__ __forceinline__ Base prepare(int64_t t0) const {
        Base bs;
        bs.e0 = t0 / d;
        bs.j0 = (uint32_t)(t0 - bs.e0 * d);
        return bs;
    }
    __device__ __forceinline__ void load(const Base& bs, uint32_t off, float* v) const {
        uint32_t jj = bs.j0 + off, q = jj / (uint32_t)d, j = jj - q * (uint32_t)d;
        int64_t e = bs.e0 + q;
        float x = __ldg(Z + (size_t)__ldg(erow + e) * ld + j);
        float y = __ldg(Z + (size_t)__ldg(col + e) * ld + j);
        v[0] = fmul(x, x);
        v[1] = fmul(y, y);
    }
    // streaming form: the edge's two row pointers are fetched once per edge, not once per element
    struct Iter { int64_t e; int j; const float* pa; const float* pb; };
    __device__ __forceinline__ Iter iter(int64_t t) const {
        Iter it;
        it.e = t / d;
        it.j = (int)(t - it.e * d);
        it.pa = Z + (size_t)__ldg(erow + it.e) * ld;
        it.pb = Z + (size_t)__ldg(col + it.e) * ld;
        return it;
    }
    __device__ __forceinline__ void next(Iter& it, float* v) const {
        const float x = __ldg(it.pa + it.j), y = __ldg(it.pb + it.j);
        v[0] = fmul(x, x);
        v[1] = fmul(y, y);
        it.j += 32;
        if (it.j >= d) {
            do { it.j -= d; ++it.e; } while (it.j >= d);
            // (the last element of the array may step one edge past the end: clamp the index loads, the pointers are unused)
            const int64_t ec = it.e < e_count ? it.e : e_count - 1;
            it.pa = Z + (size_t)__ldg(erow + ec) * ld;
            it.pb = Z + (size_t)__ldg(col + ec) * ld;
        }
    }
};

// The <= 31 elements the finish kernel reads itself (everything past the last complete cascade row; the whole
// array when n < 8), served from a 32-float buffer: a row-partitioned run all-reduces them with the level-1
// slots, so the finish never touches Z (which a faster rank may already be overwriting with its next sweep).
struct ElemValues {
    static constexpr int NQ = 1;
    const float* vals;
    struct Base { int dummy; };
    __device__ __forceinline__ Base prepare(int64_t) const { return Base{0}; }
    __device__ __forceinline__ void load(const Base&, uint32_t off, float* v) const { v[0] = vals[off]; }
    struct Iter { uint32_t off; };
    __device__ __forceinline__ Iter iter(int64_t t) const { return Iter{(uint32_t)t}; }
    __device__ __forceinline__ void next(Iter& it, float* v) const { v[0] = vals[it.off & 31]; it.off += 32; }
};

template <class Elem>
__global__ void k_tail_values(Elem elem, CascadeShape sh, float* __restrict__ vals) {
    const int64_t base = sh.n < 8 ? 0 : sh.ni * 32;
    const int off = threadIdx.x;
    float v[Elem::NQ];
    v[0] = 0.0f;
    if (base + off < sh.n) {
        typename Elem::Base bs = elem.prepare(base);
        elem.load(bs, (uint32_t)off, v);
    }
    vals[off] = v[0];
}

constexpr int kCascadeWarps = 16;

// level-1 buffer layout (floats): P1[(n1_nodes + 1)][NQ][32] | R0[NQ][32]   (= n1_nodes + 2 slots)
// level-2 scratch           : P2[(n2_full + 1)][NQ][32]
template <class Elem>
__global__ void __launch_bounds__(kCascadeWarps * 32)
k_cascade_l01(Elem elem, CascadeShape sh, float* __restrict__ ws, int64_t node_lo,
              const clane_patience* __restrict__ st) {
    constexpr int NQ = Elem::NQ;
    extern __shared__ float part[];  // [step][NQ][32]
    if (st != nullptr && st->stop) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t node = node_lo + blockIdx.x;
    const int step = (int)sh.step;
    const int64_t row0 = node * sh.node1_rows;
    const bool full = node < sh.n1_full;
    const int nchunks = full ? step : (int)sh.c_rem;
    float* P1 = ws;
    float* R0 = ws + (size_t)(sh.n1_nodes + 1) * 32 * NQ;

    for (int ch = warp; ch < nchunks; ch += kCascadeWarps) {
        const int64_t t0 = (row0 + (int64_t)ch * step) * 32;
        typename Elem::Iter it = elem.iter(t0 + lane);
        float acc[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[q] = 0.0f;
        for (int r = 0; r < step; r += 8) {
            float v[8][NQ];
#pragma unroll
            for (int u = 0; u < 8; ++u) elem.next(it, v[u]);
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int q = 0; q < NQ; ++q) acc[q] = fadd(acc[q], v[u][q]);
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) part[(ch * NQ + q) * 32 + lane] = acc[q];
    }
    if (!full && sh.r_rem > 0 && warp == (int)(sh.c_rem % kCascadeWarps)) {
        // leftover rows of the last, incomplete chunk: they stay in acc[0] to the end
        const int64_t t0 = (row0 + sh.c_rem * step) * 32;
        typename Elem::Iter it = elem.iter(t0 + lane);
        float acc[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[q] = 0.0f;
        for (int r = 0; r < (int)sh.r_rem; ++r) {
            float v[NQ];
            elem.next(it, v);
#pragma unroll
            for (int q = 0; q < NQ; ++q) acc[q] = fadd(acc[q], v[q]);
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) R0[q * 32 + lane] = acc[q];
    }
    __syncthreads();
    if (warp < NQ) {
        const int q = warp;
        float acc = 0.0f;
        for (int ch = 0; ch < nchunks; ++ch) acc = fadd(acc, part[(ch * NQ + q) * 32 + lane]);
        P1[((size_t)node * NQ + q) * 32 + lane] = acc;
    }
}

// Advance the patience state machine of Embedder.propagate (embedder.py:98-108).
__device__ __forceinline__ void patience_step(clane_patience* st, float amount, float* log, int log_cap) {
    const int sweep = st->sweeps;
    if (log != nullptr && sweep < log_cap) log[sweep] = amount;
    st->last_amount = amount;
    st->sweeps = sweep + 1;
    if (st->minimum > amount) { st->patience = st->tol; st->minimum = amount; }
    else st->patience -= 1;
    if (st->patience == 0) st->stop = 1;
    if (st->max_sweeps > 0 && st->sweeps >= st->max_sweeps) st->stop = 1;
}

template <class Elem>
__global__ void __launch_bounds__(1024)
k_cascade_finish(Elem elem, CascadeShape sh, const float* __restrict__ ws, float* __restrict__ P2,
                 float* __restrict__ out, clane_patience* __restrict__ st, float* __restrict__ log, int log_cap,
                 unsigned* __restrict__ reset_counter) {
    constexpr int NQ = Elem::NQ;
    __shared__ float lanes[NQ][32];
    if (st != nullptr && st->stop) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int step = (int)sh.step;
    const float* P1 = ws;
    const float* R0 = ws + (size_t)(sh.n1_nodes + 1) * 32 * NQ;
    if (threadIdx.x == 0 && reset_counter != nullptr) *reset_counter = 0u;

    // level 2: complete nodes in parallel (one warp per (node, quantity)); slot n2_full holds
    // the sum of the complete level-1 nodes after the last complete level-2 node.
    const int64_t n2_slots = sh.n2_full + 1;
    for (int64_t item = warp; item < n2_slots * NQ; item += nwarps) {
        const int64_t k = item / NQ;
        const int q = (int)(item - k * NQ);
        const int64_t first = k * step;
        const int cnt = (k < sh.n2_full) ? step : (int)(sh.n1_full - first);
        float acc = 0.0f;
        if (step <= 32) {
            float v[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) v[u] = (u < cnt) ? P1[((size_t)(first + u) * NQ + q) * 32 + lane] : 0.0f;
#pragma unroll
            for (int u = 0; u < 32; ++u)
                if (u < cnt) acc = fadd(acc, v[u]);
        } else {
            int i = 0;
            for (; i + 8 <= cnt; i += 8) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = P1[((size_t)(first + i + u) * NQ + q) * 32 + lane];
#pragma unroll
                for (int u = 0; u < 8; ++u) acc = fadd(acc, v[u]);
            }
            for (; i < cnt; ++i) acc = fadd(acc, P1[((size_t)(first + i) * NQ + q) * 32 + lane]);
        }
        P2[((size_t)k * NQ + q) * 32 + lane] = acc;
    }
    __syncthreads();
    if (warp < NQ) {
        const int q = warp;
        float acc3 = 0.0f;
        for (int64_t k = 0; k < sh.n2_full; ++k) acc3 = fadd(acc3, P2[((size_t)k * NQ + q) * 32 + lane]);
        float a = 0.0f;
        if (sh.ni > 0) {
            const float r0 = (sh.rem_rows > 0 && sh.r_rem > 0) ? R0[q * 32 + lane] : 0.0f;
            const float r1 = (sh.rem_rows > 0) ? P1[((size_t)sh.n1_full * NQ + q) * 32 + lane] : 0.0f;
            const float r2 = P2[((size_t)sh.n2_full * NQ + q) * 32 + lane];
            a = fadd(r0, r1);
            a = fadd(a, r2);
            a = fadd(a, acc3);
        }
        lanes[q][lane] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float result[NQ];
        if (sh.n < 8) {
            // ATen's scalar path: four interleaved accumulators
            float x[8][NQ];
            typename Elem::Base bs = elem.prepare(0);
            for (int i = 0; i < (int)sh.n; ++i) elem.load(bs, (uint32_t)i, x[i]);
            for (int q = 0; q < NQ; ++q) {
                float p4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                const int qd = (int)sh.n / 4;
                for (int i = 0; i < qd; ++i)
                    for (int k = 0; k < 4; ++k) p4[k] = fadd(p4[k], x[4 * i + k][q]);
                for (int i = 4 * qd; i < (int)sh.n; ++i) p4[0] = fadd(p4[0], x[i][q]);
                for (int k = 1; k < 4; ++k) p4[0] = fadd(p4[0], p4[k]);
                result[q] = p4[0];
            }
        } else {
            // leftover 8-vectors -> ILP row 0; combine the four ILP rows; scalar tail; 8 lanes
            typename Elem::Base bs = elem.prepare(sh.ni * 32);
            const int nleft = (int)(sh.nv - sh.ni * 4);
            for (int v = 0; v < nleft; ++v)
                for (int l = 0; l < 8; ++l) {
                    float x[NQ];
                    elem.load(bs, (uint32_t)(v * 8 + l), x);
                    for (int q = 0; q < NQ; ++q) lanes[q][l] = fadd(lanes[q][l], x[q]);
                }
            const int ntail = (int)(sh.n - sh.nv * 8);
            float tail[8][NQ];
            for (int k = 0; k < ntail; ++k) elem.load(bs, (uint32_t)(nleft * 8 + k), tail[k]);
            for (int q = 0; q < NQ; ++q) {
                for (int k = 1; k < 4; ++k)
                    for (int l = 0; l < 8; ++l) lanes[q][l] = fadd(lanes[q][l], lanes[q][k * 8 + l]);
                float o = 0.0f;
                for (int k = 0; k < ntail; ++k) o = fadd(o, tail[k][q]);
                for (int l = 0; l < 8; ++l) o = fadd(o, lanes[q][l]);
                result[q] = o;
            }
        }
        if (out != nullptr)
            for (int q = 0; q < NQ; ++q) out[q] = result[q];
        if (st != nullptr) patience_step(st, result[0], log, log_cap);
    }
}

// level-1 nodes [node_lo, node_hi) of the cascade over n elements -> p1
template <class Elem>
inline int cascade_launch_l01(const Elem& elem, int64_t n, int64_t node_lo, int64_t node_hi, float* p1,
                              const clane_patience* st, cudaStream_t s) {
    CascadeShape sh = cascade_shape(n);
    if (node_lo < 0 || node_hi > sh.n1_nodes || node_lo > node_hi) return CLANE_EINVAL;
    if (node_hi > node_lo) {
        const size_t smem = (size_t)sh.step * Elem::NQ * 32 * sizeof(float);
        k_cascade_l01<Elem><<<(unsigned)(node_hi - node_lo), kCascadeWarps * 32, smem, s>>>(elem, sh, p1, node_lo, st);
        CLANE_LAUNCH_CHECK();
    }
    return CLANE_OK;
}

// levels 2-3, tails, final combine (+ patience) from a complete p1
template <class Elem>
inline int cascade_launch_finish(const Elem& elem, int64_t n, const float* p1, float* p2, float* out,
                                 clane_patience* st, float* log, int log_cap, unsigned* reset_counter,
                                 cudaStream_t s, int threads = 1024) {
    CascadeShape sh = cascade_shape(n);
    k_cascade_finish<Elem><<<1, threads, 0, s>>>(elem, sh, p1, p2, out, st, log, log_cap, reset_counter);
    CLANE_LAUNCH_CHECK();
    return CLANE_OK;
}

// enqueue a full cascade sum of n elements; result(s) -> out[0..NQ)
template <class Elem>
inline int cascade_launch(const Elem& elem, int64_t n, float* p1, float* p2, float* out, clane_patience* st,
                          float* log, int log_cap, cudaStream_t s) {
    CascadeShape sh = cascade_shape(n);
    int rc = cascade_launch_l01(elem, n, 0, sh.n1_nodes, p1, st, s);
    if (rc != CLANE_OK) return rc;
    return cascade_launch_finish(elem, n, p1, p2, out, st, log, log_cap, nullptr, s);
}

// Fused path: the sweep kernel already produced one 32-lane partial per level-0 chunk
// (P0[chunk][32], chunk = group of rows); reduce `step` consecutive chunks per level-1 node.
__global__ void __launch_bounds__(256)
k_level1_from_p0(CascadeShape sh, const float* __restrict__ P0, float* __restrict__ p1,
                 const clane_patience* __restrict__ st) {
    if (st != nullptr && st->stop) return;
    const int lane = threadIdx.x & 31;
    const int64_t node = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (node >= sh.n1_nodes) return;
    const int step = (int)sh.step;
    const int cnt = node < sh.n1_full ? step : (int)sh.c_rem;
    const float* src = P0 + (size_t)node * step * 32 + lane;
    float acc = 0.0f;
    if (step <= 32) {   // fused mode: chunks of <= 1024 elements, so <= 32 chunks per node: one round trip
        float v[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] = (u < cnt) ? __ldg(src + (size_t)u * 32) : 0.0f;
#pragma unroll
        for (int u = 0; u < 32; ++u)
            if (u < cnt) acc = fadd(acc, v[u]);
    } else {
        int i = 0;
        for (; i + 8 <= cnt; i += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (size_t)(i + u) * 32);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = fadd(acc, v[u]);
        }
        for (; i < cnt; ++i) acc = fadd(acc, __ldg(src + (size_t)i * 32));
    }
    p1[(size_t)node * 32 + lane] = acc;
    if (node == sh.n1_nodes - 1 && sh.rem_rows > 0 && sh.r_rem > 0)   // leftover rows = the last, partial group
        p1[(size_t)(sh.n1_nodes + 1) * 32 + lane] = __ldg(P0 + (size_t)(sh.n1_full * step + sh.c_rem) * 32 + lane);
}

}  // namespace clane
