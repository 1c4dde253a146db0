// libclane_b200.so -- device kernels and the kernel-level C-ABI (include/clane_b200.h).
// Hand-written CUDA for sm_100a.  HBM/L2-bound gather work: no tensor cores.
//
// Reference call sites replaced (all under /root/reference/clane/):
//   graph.py:118-128  Graph.build_P          -> k_dots + cascade(ElemGatherSq2) + k_row_softmax
//   similarity.py:26-37 CosineSimilarity     -> k_dots + cascade(ElemGatherSq2)
//   embedder.py:84-94 Jacobi sweep + L1       -> k_sweep (sweep.cuh; fused L1 partials) + cascade levels 1-3
//   embedder.py:98-108 patience               -> patience_step (cascade.cuh)
#include <math_constants.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "cascade.cuh"
#include "common.cuh"
#include "plan.cuh"
#include "sweep.cuh"

namespace clane {

// ------------------------------------------------------------------------------------------
// edge -> source row (first row of A.indices(), graph.py:119)
// ------------------------------------------------------------------------------------------
__global__ void k_edge_rows(const int32_t* __restrict__ rowptr, int32_t n, int64_t e, int32_t* __restrict__ erow) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e) return;
    int lo = 0, hi = n;  // largest v with rowptr[v] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(rowptr + mid) <= i) lo = mid; else hi = mid;
    }
    erow[i] = lo;
}

// ------------------------------------------------------------------------------------------
// per-edge dot, sequential over the feature index (similarity.py:35-37)
//   d < 400 : ATen native bmm loop  acc = fl(acc + fl(a*b))
//   d >= 400: oneMKL                acc = fma(a, b, acc)
// ------------------------------------------------------------------------------------------
template <bool kFma>
__global__ void __launch_bounds__(256)
k_dots(const float* __restrict__ Z, int ld, int d, const int32_t* __restrict__ erow,
       const int32_t* __restrict__ col, int64_t e_lo, int64_t e_hi, float* __restrict__ dots) {
    const int64_t e = e_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= e_hi) return;
    const float* a = Z + (size_t)__ldg(erow + e) * ld;
    const float* b = Z + (size_t)__ldg(col + e) * ld;
    float acc = 0.0f;
    const int d4 = d & ~3;
    int j = 0;
    for (; j + 16 <= d4; j += 16) {
        float4 x[4], y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { x[u] = ldg4(a + j + 4 * u); y[u] = ldg4(b + j + 4 * u); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (kFma) {
                acc = ffma(x[u].x, y[u].x, acc); acc = ffma(x[u].y, y[u].y, acc);
                acc = ffma(x[u].z, y[u].z, acc); acc = ffma(x[u].w, y[u].w, acc);
            } else {
                acc = fadd(acc, fmul(x[u].x, y[u].x)); acc = fadd(acc, fmul(x[u].y, y[u].y));
                acc = fadd(acc, fmul(x[u].z, y[u].z)); acc = fadd(acc, fmul(x[u].w, y[u].w));
            }
        }
    }
    for (; j < d; ++j) {
        const float x = __ldg(a + j), y = __ldg(b + j);
        acc = kFma ? ffma(x, y, acc) : fadd(acc, fmul(x, y));
    }
    dots[e] = acc;
}

// ------------------------------------------------------------------------------------------
// Sleef_expf_u10 (what ATen's vectorised softmax calls)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float pow2if(int q) { return __int_as_float((q + 127) << 23); }

__device__ __forceinline__ float sleef_expf_u10(float d) {
    const float R_LN2f = 1.442695040888963407359924681001892137426645954152985934135449406931f;
    const float L2Uf = 0.693145751953125f, L2Lf = 1.428606765330187045e-06f;
    const float qf = rintf(fmul(d, R_LN2f));
    const int q = (int)qf;
    float s = ffma(qf, -L2Uf, d);
    s = ffma(qf, -L2Lf, s);
    float u = 0.000198527617612853646278381f;
    u = ffma(u, s, 0.00139304355252534151077271f);
    u = ffma(u, s, 0.00833336077630519866943359f);
    u = ffma(u, s, 0.0416664853692054748535156f);
    u = ffma(u, s, 0.166666671633720397949219f);
    u = ffma(u, s, 0.5f);
    u = fadd(1.0f, ffma(fmul(s, s), u, s));
    const int q1 = q >> 1;
    u = fmul(fmul(u, pow2if(q1)), pow2if(q - q1));
    if (d < -104.0f) u = 0.0f;
    if (d > 100.0f) u = CUDART_INF_F;
    return u;
}

// ------------------------------------------------------------------------------------------
// per-source softmax (graph.py:122-123): ATen last-dim softmax =
//   e_i = Sleef_expf_u10(s_i - max), sum by the 16-lane vec::reduce_all tree (sequential for
//   rows shorter than 16), p_i = e_i * (1 / sum).
// One warp per row.  Optional global divisor c = fl(sqrt(S1)) * fl(sqrt(S2)) (similarity.py:37).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_row_softmax(const float* scores, const float* __restrict__ norms2, int32_t row_lo, int32_t row_hi,
              const int32_t* __restrict__ rowptr, float* w) {   // scores and w may be the same array (in place)
    const int row = row_lo + (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= row_hi) return;
    const int a = __ldg(rowptr + row), k = __ldg(rowptr + row + 1) - a;
    if (k == 0) return;
    const bool div = norms2 != nullptr;
    float c = 1.0f;
    if (div) c = fmul(__fsqrt_rn(__ldg(norms2)), __fsqrt_rn(__ldg(norms2 + 1)));

    float m = -CUDART_INF_F;
    for (int i = lane; i < k; i += 32) {
        float s = scores[a + i];
        if (div) s = __fdiv_rn(s, c);
        m = fmaxf(m, s);
    }
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, h));

    float sum;
    float e_reg = 0.0f;  // rows of <= 32 edges keep e in a register
    if (k < 16) {
        if (lane < k) {
            float s = scores[a + lane];
            if (div) s = __fdiv_rn(s, c);
            e_reg = sleef_expf_u10(fsub(s, m));
        }
        sum = __shfl_sync(kFull, e_reg, 0);
        for (int i = 1; i < k; ++i) sum = fadd(sum, __shfl_sync(kFull, e_reg, i));
    } else {
        float acc = 0.0f;
        for (int base = 0; base < k; base += 32) {
            const int i = base + lane;
            float e = 0.0f;
            if (i < k) {
                float s = scores[a + i];
                if (div) s = __fdiv_rn(s, c);
                e = sleef_expf_u10(fsub(s, m));
                if (k > 32) w[a + i] = e;
            }
            e_reg = e;
            const float hi = __shfl_down_sync(kFull, e, 16);
            // lane l < 16 accumulates e_l, e_{l+16}, e_{l+32}, ... in order (missing = +0)
            acc = fadd(acc, e);
            acc = fadd(acc, hi);
        }
#pragma unroll
        for (int h = 8; h >= 1; h >>= 1) acc = fadd(acc, __shfl_xor_sync(kFull, acc, h));
        sum = __shfl_sync(kFull, acc, 0);
    }
    const float inv = __fdiv_rn(1.0f, sum);
    if (k <= 32) {
        if (lane < k) w[a + lane] = fmul(e_reg, inv);
    } else {
        for (int i = lane; i < k; i += 32) w[a + i] = fmul(w[a + i], inv);
    }
}

// col[e] * ld / 4: the float4 index the sweep gathers from (one multiply per edge, once per graph)
__global__ void k_col_offsets(const int32_t* __restrict__ col, int64_t e, int ld, int32_t* __restrict__ off) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < e) off[i] = col[i] * (ld >> 2);
}

__global__ void k_cosine_finalize(const float* __restrict__ dots, const float* __restrict__ norms2, int64_t e,
                                  float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e) return;
    const float c = fmul(__fsqrt_rn(__ldg(norms2)), __fsqrt_rn(__ldg(norms2 + 1)));
    out[i] = __fdiv_rn(dots[i], c);
}

__global__ void k_patience_reset(clane_patience* st, int tol, int max_sweeps) {
    st->minimum = CUDART_INF_F;
    st->patience = tol;
    st->tol = tol;
    st->sweeps = 0;
    st->max_sweeps = max_sweeps;
    st->stop = 0;
    st->last_amount = 0.0f;
    st->reserved = 0;
}

}  // namespace clane

using namespace clane;

// =============================================================================================
// C-ABI
// =============================================================================================
extern "C" {

int clane_version(void) { return 100; }

const char* clane_error_string(int code) {
    switch (code) {
        case CLANE_OK: return "ok";
        case CLANE_EINVAL: return "clane: invalid argument";
        case CLANE_ERANGE: return "clane: index out of range";
        case CLANE_EWORKSPACE: return "clane: workspace too small";
        case CLANE_ENODEVICE: return "clane: no sm_100 CUDA device";
        case CLANE_ENOENT: return "clane: cannot open file";
        case CLANE_EPARSE: return "clane: malformed edge line";
        case CLANE_EUNKNOWNID: return "clane: edge endpoint not in V";
        case CLANE_ENOMEM: return "clane: out of host memory";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "clane: unknown error";
    }
}

int clane_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    CLANE_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CLANE_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return CLANE_OK;
}

int32_t clane_padded_ld(int32_t d) { return (d + 3) & ~3; }

int clane_cascade_shape(int64_t n_elems, int64_t* n1_nodes, int64_t* elems_per_node) {
    if (n_elems < 0) return CLANE_EINVAL;
    CascadeShape sh = cascade_shape(n_elems);
    if (n1_nodes) *n1_nodes = sh.n1_nodes;
    if (elems_per_node) *elems_per_node = sh.node1_rows * 32;
    return CLANE_OK;
}

int clane_edge_rows(const int32_t* d_rowptr, int32_t n, int64_t e, int32_t* d_erow, clane_stream_t s) {
    if (!d_rowptr || !d_erow || n < 0 || e < 0) return CLANE_EINVAL;
    if (e == 0) return CLANE_OK;
    k_edge_rows<<<(unsigned)((e + 255) / 256), 256, 0, (cudaStream_t)s>>>(d_rowptr, n, e, d_erow);
    CLANE_LAUNCH_CHECK();
    return CLANE_OK;
}

int clane_scores_cosine(clane_plan* plan, const float* d_Z, const int32_t* d_erow, const int32_t* d_col,
                        int64_t edge_lo, int64_t edge_hi, float* d_dots, float* d_norms2, clane_stream_t s) {
    if (!plan || !d_Z || !d_erow || !d_col || !d_dots || !d_norms2) return CLANE_EINVAL;
    if (edge_lo < 0 || edge_hi > plan->e || edge_lo > edge_hi) return CLANE_EINVAL;
    cudaStream_t st = (cudaStream_t)s;
    if (edge_hi > edge_lo) {
        const unsigned grid = (unsigned)((edge_hi - edge_lo + 255) / 256);
        if (plan->d < 400) k_dots<false><<<grid, 256, 0, st>>>(d_Z, plan->ld, plan->d, d_erow, d_col, edge_lo, edge_hi, d_dots);
        else k_dots<true><<<grid, 256, 0, st>>>(d_Z, plan->ld, plan->d, d_erow, d_col, edge_lo, edge_hi, d_dots);
        CLANE_LAUNCH_CHECK();
    }
    ElemGatherSq2 el{d_Z, d_erow, d_col, plan->d, plan->ld};
    return cascade_launch(el, plan->e * (int64_t)plan->d, plan->d_p1, plan->d_p2, d_norms2, nullptr, nullptr, 0, st);
}

int clane_row_softmax(const float* d_scores, const float* d_norms2, int32_t row_lo, int32_t row_hi,
                      const int32_t* d_rowptr, float* d_w, clane_stream_t s) {
    if (!d_scores || !d_rowptr || !d_w || row_lo < 0 || row_hi < row_lo) return CLANE_EINVAL;
    if (row_hi == row_lo) return CLANE_OK;
    const unsigned grid = (unsigned)(((int64_t)(row_hi - row_lo) * 32 + 255) / 256);
    k_row_softmax<<<grid, 256, 0, (cudaStream_t)s>>>(d_scores, d_norms2, row_lo, row_hi, d_rowptr, d_w);
    CLANE_LAUNCH_CHECK();
    return CLANE_OK;
}

int clane_cosine_finalize(const float* d_dots, const float* d_norms2, int64_t e, float* d_out, clane_stream_t s) {
    if (!d_dots || !d_norms2 || !d_out || e < 0) return CLANE_EINVAL;
    if (e == 0) return CLANE_OK;
    k_cosine_finalize<<<(unsigned)((e + 255) / 256), 256, 0, (cudaStream_t)s>>>(d_dots, d_norms2, e, d_out);
    CLANE_LAUNCH_CHECK();
    return CLANE_OK;
}

int clane_build_p_cosine(clane_plan* plan, const float* d_Z, const int32_t* d_rowptr, const int32_t* d_erow,
                         const int32_t* d_col, float* d_w, float* d_norms2, clane_stream_t s) {
    if (!plan || !d_rowptr || !plan->has_schedule) return CLANE_EINVAL;
    // rows [row_lo, row_hi) own the contiguous edge range [rowptr[row_lo], rowptr[row_hi])
    int rc = clane_scores_cosine(plan, d_Z, d_erow, d_col, plan->edge_lo, plan->edge_hi, d_w, d_norms2, s);
    if (rc != CLANE_OK) return rc;
    return clane_row_softmax(d_w, d_norms2, plan->row_lo, plan->row_hi, d_rowptr, d_w, s);
}

// timing bracket: an ordinary record, or -- while the sweep is being captured -- an external event-record node of
// the graph, so that the brackets measure the kernels as they run in the replayed graph
static cudaError_t prof_record(cudaEvent_t ev, cudaStream_t st) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    return cudaEventRecordWithFlags(ev, st, cap == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault);
}

static int sweep_enqueue(clane_plan* plan, const float* d_X, const float* d_Zcur, float* d_Znext,
                         const int32_t* d_rowptr, const int32_t* d_col, const float* d_w, float gamma, float* d_amount,
                         clane_patience* d_state, float* d_amounts_log, int32_t log_cap, cudaStream_t st) {
    const bool want_l1 = d_amount != nullptr || d_state != nullptr;
    SweepParams p;
    p.X = d_X; p.Zc = d_Zcur; p.Zn = d_Znext;
    p.ld = plan->ld; p.d = plan->d; p.n = plan->n;
    p.rowptr = d_rowptr; p.coloff = plan->d_coloff; p.w = d_w; p.gamma = gamma;
    p.tasks = static_cast<const SweepTask*>(plan->d_tasks); p.n_tasks = plan->n_tasks;
    p.row_lo = plan->row_lo;
    p.G = plan->G; p.nslab = plan->nslab;
    p.fuse = (plan->fuse && want_l1) ? 1 : 0;
    p.P0 = plan->d_P0;
    p.hub_info = static_cast<const int4*>(plan->d_hub_info); p.n_hub_rows = plan->n_hub_rows;
    p.limit = plan->limit; p.ntail4 = plan->ntail4; p.nslab32b = plan->nslab32b; p.sld = plan->nslab32b * 32;
    p.hubS = static_cast<float4*>(plan->d_hubS);
    p.hubT = static_cast<float4*>(plan->d_hubT);
    p.hub_cnt = plan->d_hub_cnt; p.hub_done = plan->d_hub_done;
    p.chain_spin_ns = plan->chain_spin_ns;
    p.st = d_state;
    p.n_remote = 0;
    static const bool no_peer_stores = getenv("CLANE_DEBUG_NO_PEER_STORES") != nullptr;   // timing experiments only
    p.mc = nullptr;
    for (int t = 0; t < 2 && plan->n_peers > 1 && !no_peer_stores; ++t)
        if (plan->peers[t][plan->self_rank] == d_Znext) {
            if (plan->mc[t] != nullptr) { p.mc = plan->mc[t]; continue; }   // one multicast store instead of n - 1 unicast ones
            for (int r = 0; r < plan->n_peers; ++r)
                if (r != plan->self_rank) p.peer[p.n_remote++] = plan->peers[t][r];
        }
    // Hub rows: their segments are the first tasks of the row kernel; the chains follow on the same stream.
    const int64_t chain_ctas = (int64_t)p.n_hub_rows * (plan->nslab32b + (plan->ntail4 > 0 ? 1 : 0));
    const int64_t row_ctas = ((int64_t)p.n_tasks * plan->nslab + kRowWarps - 1) / kRowWarps;
    const bool prof = plan->profile;
    if (prof) CLANE_CUDA(prof_record(plan->ev_prof[0], st));
    // The early chain passes only help if the device runs them beside the row kernel; with a single hardware work
    // queue (CUDA_DEVICE_MAX_CONNECTIONS=1) kernels of different streams run in submission order and the early pass
    // would just spin into its time-out before every row kernel.
    static const bool overlap = [] {
        if (getenv("CLANE_NO_CHAIN_OVERLAP") != nullptr) return false;
        const char* q = getenv("CUDA_DEVICE_MAX_CONNECTIONS");
        return !(q != nullptr && atoi(q) == 1);
    }();
    const int per_row = plan->nslab32b + (plan->ntail4 > 0 ? 1 : 0);
    const int n_long = plan->n_long_hub_rows, n_short = plan->n_hub_rows - n_long;
    if (chain_ctas > 0 && overlap) {   // early chain passes: beside the row kernel, waiting on its segment warps
        CLANE_CUDA(cudaEventRecord(plan->ev_fork, st));
        CLANE_CUDA(cudaStreamWaitEvent(plan->side, plan->ev_fork, 0));
        if (n_long > 0) {
            p.hub_first = 0;
            k_hub_chain<true, false><<<(unsigned)(n_long * per_row), kChainThreads, kChainSmemBytes, plan->side>>>(p);
            CLANE_LAUNCH_CHECK();
        }
        CLANE_CUDA(cudaEventRecord(plan->ev_join, plan->side));
        if (n_short > 0) {             // its own stream: not behind the long rows' chains
            CLANE_CUDA(cudaStreamWaitEvent(plan->side2, plan->ev_fork, 0));
            p.hub_first = n_long;
            k_hub_chain<true, true><<<(unsigned)(n_short * per_row), kChainThreads, chain_smem_bytes(kLightStages), plan->side2>>>(p);
            CLANE_LAUNCH_CHECK();
            CLANE_CUDA(cudaEventRecord(plan->ev_join2, plan->side2));
        }
    }
    if (prof) CLANE_CUDA(prof_record(plan->ev_prof[1], st));
    if (row_ctas > 0) {
        k_sweep_rows<<<(unsigned)row_ctas, kRowThreads, 0, st>>>(p);
        CLANE_LAUNCH_CHECK();
    }
    if (prof) CLANE_CUDA(prof_record(plan->ev_prof[2], st));
    if (chain_ctas > 0 && overlap) {
        CLANE_CUDA(cudaStreamWaitEvent(st, plan->ev_join, 0));
        if (n_short > 0) CLANE_CUDA(cudaStreamWaitEvent(st, plan->ev_join2, 0));
    }
    if (prof) CLANE_CUDA(prof_record(plan->ev_prof[4], st));
    if (chain_ctas > 0) {              // late pass: whatever the early one left, and the reset of its flags
        p.hub_first = 0;
        k_hub_chain<false, true><<<(unsigned)chain_ctas, kChainThreads, chain_smem_bytes(kLightStages), st>>>(p);
        CLANE_LAUNCH_CHECK();
    }
    if (prof) CLANE_CUDA(prof_record(plan->ev_prof[5], st));
    struct ProfTail {   // records the end-of-sweep event on every exit path below
        clane_plan* pl; cudaStream_t s;
        ~ProfTail() { if (pl->profile) prof_record(pl->ev_prof[3], s); }
    } prof_tail{plan, st};
    if (!want_l1) return CLANE_OK;
    const int64_t n_elems = (int64_t)plan->n * plan->d;
    ElemAbsDiff el{d_Znext, d_Zcur, plan->d, plan->ld};
    if (p.fuse) {
        CascadeShape sh = cascade_shape(n_elems);
        if (plan->n_fix_groups > 0) {
            k_fix_chunks<<<(unsigned)(((int64_t)plan->n_fix_groups * 32 + 255) / 256), 256, 0, st>>>(
                d_Znext, d_Zcur, plan->d, plan->n, plan->G, plan->d_fix_groups, plan->n_fix_groups, plan->d_P0, d_state);
            CLANE_LAUNCH_CHECK();
        }
        if (sh.n1_nodes > 0) {
            k_level1_from_p0<<<(unsigned)((sh.n1_nodes * 32 + 255) / 256), 256, 0, st>>>(sh, plan->d_P0, plan->d_p1, d_state);
            CLANE_LAUNCH_CHECK();
        }
        return cascade_launch_finish(el, n_elems, plan->d_p1, plan->d_p2, d_amount, d_state, d_amounts_log, log_cap,
                                     nullptr, st);
    }
    return cascade_launch(el, n_elems, plan->d_p1, plan->d_p2, d_amount, d_state, d_amounts_log, log_cap, st);
}

int clane_sweep(clane_plan* plan, const float* d_X, const float* d_Zcur, float* d_Znext, const int32_t* d_rowptr,
                const int32_t* d_col, const float* d_w, float gamma, float* d_amount, clane_patience* d_state,
                float* d_amounts_log, int32_t log_cap, clane_stream_t s) {
    if (!plan || !plan->has_schedule || !d_X || !d_Zcur || !d_Znext || !d_rowptr) return CLANE_EINVAL;
    if (plan->e > 0 && (!d_col || !d_w)) return CLANE_EINVAL;
    cudaStream_t st = (cudaStream_t)s;
    if (plan->e > 0 && plan->coloff_src != d_col) {   // first sweep with this column array
        k_col_offsets<<<(unsigned)((plan->e + 255) / 256), 256, 0, st>>>(d_col, plan->e, plan->ld, plan->d_coloff);
        CLANE_LAUNCH_CHECK();
        plan->coloff_src = d_col;
    }
    static const bool env_graphs = getenv("CLANE_NO_GRAPHS") == nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (!plan->use_graphs || !env_graphs || cap != cudaStreamCaptureStatusNone || st == nullptr)
        return sweep_enqueue(plan, d_X, d_Zcur, d_Znext, d_rowptr, d_col, d_w, gamma, d_amount, d_state, d_amounts_log,
                             log_cap, st);
    // One sweep = up to five kernels on two streams.  Replay it as a CUDA graph: a propagate() call
    // alternates between two argument sets (Zcur/Znext swapped), so each is captured once.
    const void* key[10] = {d_X, d_Zcur, d_Znext, d_rowptr, d_col, d_w, d_amount, d_state, d_amounts_log, st};
    clane_plan::SweepGraph* slot = nullptr;
    for (auto& g : plan->graphs)
        if (g.exec && g.gamma == gamma && g.log_cap == log_cap && memcmp(g.key, key, sizeof(key)) == 0) slot = &g;
    if (!slot) {
        slot = plan->graphs[0].last_use <= plan->graphs[1].last_use ? &plan->graphs[0] : &plan->graphs[1];
        if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
        cudaGraph_t graph = nullptr;
        CLANE_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        int rc = sweep_enqueue(plan, d_X, d_Zcur, d_Znext, d_rowptr, d_col, d_w, gamma, d_amount, d_state,
                               d_amounts_log, log_cap, st);
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc != CLANE_OK || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            if (rc != CLANE_OK) return rc;
            plan->use_graphs = false;   // capture unsupported here: direct launches from now on
            return sweep_enqueue(plan, d_X, d_Zcur, d_Znext, d_rowptr, d_col, d_w, gamma, d_amount, d_state,
                                 d_amounts_log, log_cap, st);
        }
        ce = cudaGraphInstantiate(&slot->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) { slot->exec = nullptr; return (int)ce; }
        memcpy(slot->key, key, sizeof(key));
        slot->gamma = gamma;
        slot->log_cap = log_cap;
    }
    slot->last_use = ++plan->graph_clock;
    CLANE_CUDA(cudaGraphLaunch(slot->exec, st));
    return CLANE_OK;
}

int clane_l1_diff(clane_plan* plan, const float* d_Za, const float* d_Zb, float* d_out, clane_stream_t s) {
    if (!plan || !d_Za || !d_Zb || !d_out) return CLANE_EINVAL;
    ElemAbsDiff el{d_Za, d_Zb, plan->d, plan->ld};
    return cascade_launch(el, (int64_t)plan->n * plan->d, plan->d_p1, plan->d_p2, d_out, nullptr, nullptr, 0,
                          (cudaStream_t)s);
}

int clane_l1_partial(clane_plan* plan, const float* d_Za, const float* d_Zb, int64_t node_lo, int64_t node_hi,
                     float* d_p1, clane_stream_t s) {
    if (!plan || !d_Za || !d_Zb || !d_p1) return CLANE_EINVAL;
    ElemAbsDiff el{d_Za, d_Zb, plan->d, plan->ld};
    return cascade_launch_l01(el, (int64_t)plan->n * plan->d, node_lo, node_hi, d_p1, nullptr, (cudaStream_t)s);
}

int clane_l1_finish(clane_plan* plan, const float* d_Za, const float* d_Zb, float* d_p1, float* d_out,
                    clane_patience* d_state, float* d_amounts_log, int32_t log_cap, clane_stream_t s) {
    if (!plan || !d_Za || !d_Zb || !d_p1) return CLANE_EINVAL;
    ElemAbsDiff el{d_Za, d_Zb, plan->d, plan->ld};
    return cascade_launch_finish(el, (int64_t)plan->n * plan->d, d_p1, plan->d_p2, d_out, d_state, d_amounts_log,
                                 log_cap, nullptr, (cudaStream_t)s);
}

int clane_plan_set_peers(clane_plan* plan, int32_t n_peers, int32_t self_rank, const uint64_t* h_ptrs_a,
                         const uint64_t* h_ptrs_b) {
    if (!plan || n_peers < 0 || n_peers > kMaxPeers + 1 || (n_peers > 0 && (self_rank < 0 || self_rank >= n_peers)))
        return CLANE_EINVAL;
    if (n_peers > 0 && (!h_ptrs_a || !h_ptrs_b)) return CLANE_EINVAL;
    plan->n_peers = n_peers;
    plan->self_rank = self_rank;
    for (int r = 0; r < n_peers; ++r) {
        plan->peers[0][r] = reinterpret_cast<float*>(h_ptrs_a[r]);
        plan->peers[1][r] = reinterpret_cast<float*>(h_ptrs_b[r]);
    }
    for (auto& g : plan->graphs)   // cached sweeps were captured with the old peer set
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    return CLANE_OK;
}

int clane_plan_set_multicast(clane_plan* plan, uint64_t mc_a, uint64_t mc_b) {
    if (!plan) return CLANE_EINVAL;
    plan->mc[0] = reinterpret_cast<float*>(mc_a);
    plan->mc[1] = reinterpret_cast<float*>(mc_b);
    for (auto& g : plan->graphs)
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    return CLANE_OK;
}

int clane_l1_tail_values(clane_plan* plan, const float* d_Za, const float* d_Zb, float* d_vals, clane_stream_t s) {
    if (!plan || !d_Za || !d_Zb || !d_vals) return CLANE_EINVAL;
    ElemAbsDiff el{d_Za, d_Zb, plan->d, plan->ld};
    k_tail_values<<<1, 32, 0, (cudaStream_t)s>>>(el, cascade_shape((int64_t)plan->n * plan->d), d_vals);
    CLANE_LAUNCH_CHECK();
    return CLANE_OK;
}

int clane_l1_finish_values(clane_plan* plan, float* d_p1, const float* d_vals, float* d_out, clane_patience* d_state,
                           float* d_amounts_log, int32_t log_cap, clane_stream_t s) {
    if (!plan || !d_p1 || !d_vals) return CLANE_EINVAL;
    ElemValues el{d_vals};
    return cascade_launch_finish(el, (int64_t)plan->n * plan->d, d_p1, plan->d_p2, d_out, d_state, d_amounts_log,
                                 log_cap, nullptr, (cudaStream_t)s);
}

int clane_plan_profile(clane_plan* plan, int enable) {
    if (!plan) return CLANE_EINVAL;
    if (enable && !plan->ev_prof[0])
        for (int i = 0; i < 6; ++i) CLANE_CUDA(cudaEventCreate(&plan->ev_prof[i]));
    if (plan->profile != (enable != 0))   // the brackets are event-record nodes of the replayed sweep graphs: re-capture
        for (auto& g : plan->graphs)
            if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    plan->profile = enable != 0;
    return CLANE_OK;
}

int clane_plan_profile_read(clane_plan* plan, float* h_ms) {
    if (!plan || !h_ms || !plan->ev_prof[0]) return CLANE_EINVAL;
    CLANE_CUDA(cudaEventSynchronize(plan->ev_prof[3]));
    CLANE_CUDA(cudaEventElapsedTime(&h_ms[0], plan->ev_prof[1], plan->ev_prof[2]));   // row kernel
    CLANE_CUDA(cudaEventElapsedTime(&h_ms[1], plan->ev_prof[0], plan->ev_prof[3]));   // whole sweep
    CLANE_CUDA(cudaEventElapsedTime(&h_ms[2], plan->ev_prof[5], plan->ev_prof[3]));   // exact L1 tail
    CLANE_CUDA(cudaEventElapsedTime(&h_ms[3], plan->ev_prof[4], plan->ev_prof[5]));   // hub chain kernel
    return CLANE_OK;
}

int clane_patience_reset(clane_patience* d_state, int32_t tol, int32_t max_sweeps, clane_stream_t s) {
    if (!d_state) return CLANE_EINVAL;
    k_patience_reset<<<1, 1, 0, (cudaStream_t)s>>>(d_state, tol, max_sweeps);
    CLANE_LAUNCH_CHECK();
    return CLANE_OK;
}

// one-time: opt in to > 48 KB dynamic shared memory for the sweep kernel
int clane_internal_prepare_kernels(void) {
    static bool done = false;
    if (done) return CLANE_OK;
    CLANE_CUDA(cudaFuncSetAttribute(k_hub_chain<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemBytes));
    // the cascade level-0/1 kernel needs step*NQ*128 bytes (<= 32 KB for step = 128, NQ = 2)
    done = true;
    return CLANE_OK;
}

}  // extern "C"
