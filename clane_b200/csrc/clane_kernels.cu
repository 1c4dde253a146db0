// libclane_b200.so -- device kernels and the kernel-level C-ABI (include/clane_b200.h).
// Hand-written CUDA for sm_100a.  HBM/L2-bound gather work: no tensor cores.
//
// Reference call sites replaced (all under /root/reference/clane/):
//   graph.py:118-128  Graph.build_P          -> k_dots + cascade(ElemGatherSq2) + k_row_softmax
//   similarity.py:26-37 CosineSimilarity     -> k_dots + cascade(ElemGatherSq2)
//   embedder.py:84-94 Jacobi sweep + L1       -> k_sweep_rows (+ hub path) + cascade(ElemAbsDiff)
//   embedder.py:98-108 patience               -> patience_step (cascade.cuh)
#include <math_constants.h>

#include "cascade.cuh"
#include "common.cuh"

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
       const int32_t* __restrict__ col, int64_t e_cnt, float* __restrict__ dots) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= e_cnt) return;
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
k_row_softmax(const float* __restrict__ scores, const float* __restrict__ norms2, int32_t n,
              const int32_t* __restrict__ rowptr, float* __restrict__ w) {
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
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

// ------------------------------------------------------------------------------------------
// Jacobi sweep, row update (embedder.py:92):  w[1,k] @ Z[k,d] in oneMKL sgemm order
// (SURVEY 7.1 step 4 / Appendix A.1), then t = fl(gamma*acc), z = fl(x + t).
//
// One warp per (row, 128-column slab); lane = one float4 of columns.  A column group is
// "blocked" (fixed 8-neighbour tree) iff it lies below 16*floor(d/16) and the row has >= 8
// neighbours; otherwise a sequential fma chain over the neighbours in ascending column id.
// A row's neighbours are never split across lanes: the summation order is the reference's.
// ------------------------------------------------------------------------------------------
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

__global__ void __launch_bounds__(256)
k_sweep_rows(const float* __restrict__ X, const float* __restrict__ Zc, float* __restrict__ Zn, int ld, int d,
             const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ w,
             float gamma, const int32_t* __restrict__ order, int n_rows, int nslab,
             const clane_patience* __restrict__ st) {
    if (st != nullptr && st->stop) return;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int task = (int)(gw / nslab), slab = (int)(gw - (int64_t)task * nslab);
    if (task >= n_rows) return;
    const int row = __ldg(order + task);
    const int a = __ldg(rowptr + row), k = __ldg(rowptr + row + 1) - a;
    if (k == 0) return;
    const int c = slab * 128 + lane * 4;
    const bool active = c < ld;
    const int cc = active ? c : 0;
    const int dm = (d / 16) * 16;
    const bool blk = (c < dm) && (k >= 8);
    const float* zb = Zc + cc;
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);

    for (int base = 0; base < k; base += 32) {
        const int my = base + lane;
        int cj = 0;
        float wj = 0.0f;
        if (my < k) { cj = __ldg(col + a + my); wj = __ldg(w + a + my); }
        const int cnt = min(32, k - base);
        int o = 0;
        for (; o + 8 <= cnt; o += 8) {
            float4 z[8];
            float ww[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = __shfl_sync(kFull, cj, o + i);
                ww[i] = __shfl_sync(kFull, wj, o + i);
                z[i] = ldg4(zb + (size_t)r * ld);
            }
            if (blk) {
                acc.x = blocked8(acc.x, ww, z[0].x, z[1].x, z[2].x, z[3].x, z[4].x, z[5].x, z[6].x, z[7].x);
                acc.y = blocked8(acc.y, ww, z[0].y, z[1].y, z[2].y, z[3].y, z[4].y, z[5].y, z[6].y, z[7].y);
                acc.z = blocked8(acc.z, ww, z[0].z, z[1].z, z[2].z, z[3].z, z[4].z, z[5].z, z[6].z, z[7].z);
                acc.w = blocked8(acc.w, ww, z[0].w, z[1].w, z[2].w, z[3].w, z[4].w, z[5].w, z[6].w, z[7].w);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) fma4(ww[i], z[i], acc);
            }
        }
        const int m = cnt - o;  // 0..7 leftover neighbours: sequential fma in both regimes
        if (m > 0) {
            float4 z[7];
            float ww[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const int r = __shfl_sync(kFull, cj, (o + i) & 31);
                ww[i] = __shfl_sync(kFull, wj, (o + i) & 31);
                if (i < m) z[i] = ldg4(zb + (size_t)r * ld);
            }
#pragma unroll
            for (int i = 0; i < 7; ++i)
                if (i < m) fma4(ww[i], z[i], acc);
        }
    }
    if (active) {
        const size_t off = (size_t)row * ld + c;
        const float4 x = ld_stream4(X + off);
        float4 out;
        out.x = fadd(x.x, fmul(gamma, acc.x));
        out.y = fadd(x.y, fmul(gamma, acc.y));
        out.z = fadd(x.z, fmul(gamma, acc.z));
        out.w = fadd(x.w, fmul(gamma, acc.w));
        *reinterpret_cast<float4*>(Zn + off) = out;
    }
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

size_t clane_workspace_bytes(int64_t n, int64_t e, int32_t d) {
    if (n < 0 || e < 0 || d < 1) return 0;
    size_t a = cascade_ws_floats(n * (int64_t)d, 1);
    size_t b = cascade_ws_floats(e * (int64_t)d, 2);
    return ((a > b ? a : b) + 64) * sizeof(float);
}

int clane_edge_rows(const int32_t* d_rowptr, int32_t n, int64_t e, int32_t* d_erow, clane_stream_t s) {
    if (!d_rowptr || !d_erow || n < 0 || e < 0) return CLANE_EINVAL;
    if (e == 0) return CLANE_OK;
    k_edge_rows<<<(unsigned)((e + 255) / 256), 256, 0, (cudaStream_t)s>>>(d_rowptr, n, e, d_erow);
    CLANE_LAUNCH_CHECK();
    return CLANE_OK;
}

int clane_scores_cosine(const float* d_Z, int32_t ld, int32_t d, int32_t n, int64_t e, const int32_t* d_erow,
                        const int32_t* d_col, float* d_dots, float* d_norms2, void* d_ws, size_t ws_bytes,
                        clane_stream_t s) {
    if (!d_Z || !d_erow || !d_col || !d_dots || !d_norms2 || !d_ws || d < 1 || ld < d || (ld & 3) || n < 0 || e < 0)
        return CLANE_EINVAL;
    cudaStream_t st = (cudaStream_t)s;
    if (e > 0) {
        const unsigned grid = (unsigned)((e + 255) / 256);
        if (d < 400) k_dots<false><<<grid, 256, 0, st>>>(d_Z, ld, d, d_erow, d_col, e, d_dots);
        else k_dots<true><<<grid, 256, 0, st>>>(d_Z, ld, d, d_erow, d_col, e, d_dots);
        CLANE_LAUNCH_CHECK();
    }
    ElemGatherSq2 el{d_Z, d_erow, d_col, d, ld};
    return cascade_launch(el, e * (int64_t)d, (float*)d_ws, ws_bytes, d_norms2, nullptr, nullptr, 0, st);
}

int clane_row_softmax(const float* d_scores, const float* d_norms2, int32_t n, const int32_t* d_rowptr, float* d_w,
                      clane_stream_t s) {
    if (!d_scores || !d_rowptr || !d_w || n < 0) return CLANE_EINVAL;
    if (n == 0) return CLANE_OK;
    const unsigned grid = (unsigned)(((int64_t)n * 32 + 255) / 256);
    k_row_softmax<<<grid, 256, 0, (cudaStream_t)s>>>(d_scores, d_norms2, n, d_rowptr, d_w);
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

int clane_build_p_cosine(const float* d_Z, int32_t ld, int32_t d, int32_t n, int64_t e, const int32_t* d_rowptr,
                         const int32_t* d_erow, const int32_t* d_col, float* d_w, float* d_norms2, void* d_ws,
                         size_t ws_bytes, clane_stream_t s) {
    if (!d_rowptr) return CLANE_EINVAL;
    int rc = clane_scores_cosine(d_Z, ld, d, n, e, d_erow, d_col, d_w, d_norms2, d_ws, ws_bytes, s);
    if (rc != CLANE_OK) return rc;
    return clane_row_softmax(d_w, d_norms2, n, d_rowptr, d_w, s);
}

int clane_sweep(const float* d_X, const float* d_Zcur, float* d_Znext, int32_t ld, int32_t d, int32_t n,
                const int32_t* d_rowptr, const int32_t* d_col, const float* d_w, float gamma,
                const int32_t* d_light_order, int32_t n_light, const int32_t* d_hub_rows, int32_t n_hub,
                float* d_amount, clane_patience* d_state, float* d_amounts_log, int32_t log_cap, void* d_ws,
                size_t ws_bytes, clane_stream_t s) {
    if (!d_X || !d_Zcur || !d_Znext || !d_rowptr || d < 1 || ld < d || (ld & 3) || n < 0 || n_light < 0 || n_hub < 0)
        return CLANE_EINVAL;
    if ((n_light > 0 && !d_light_order) || (n_hub > 0 && !d_hub_rows)) return CLANE_EINVAL;
    if ((n_light > 0 || n_hub > 0) && (!d_col || !d_w)) return CLANE_EINVAL;
    cudaStream_t st = (cudaStream_t)s;
    const int nslab = (ld + 127) / 128;
    if (n_hub > 0) {
        const int64_t warps = (int64_t)n_hub * nslab;
        k_sweep_rows<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(
            d_X, d_Zcur, d_Znext, ld, d, d_rowptr, d_col, d_w, gamma, d_hub_rows, n_hub, nslab, d_state);
        CLANE_LAUNCH_CHECK();
    }
    if (n_light > 0) {
        const int64_t warps = (int64_t)n_light * nslab;
        k_sweep_rows<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(
            d_X, d_Zcur, d_Znext, ld, d, d_rowptr, d_col, d_w, gamma, d_light_order, n_light, nslab, d_state);
        CLANE_LAUNCH_CHECK();
    }
    if (d_amount != nullptr || d_state != nullptr) {
        if (!d_ws) return CLANE_EINVAL;
        ElemAbsDiff el{d_Znext, d_Zcur, d, ld};
        return cascade_launch(el, (int64_t)n * d, (float*)d_ws, ws_bytes, d_amount, d_state, d_amounts_log, log_cap, st);
    }
    return CLANE_OK;
}

int clane_l1_diff(const float* d_Za, const float* d_Zb, int32_t ld, int32_t d, int32_t n, float* d_out, void* d_ws,
                  size_t ws_bytes, clane_stream_t s) {
    if (!d_Za || !d_Zb || !d_out || !d_ws || d < 1 || ld < d || n < 0) return CLANE_EINVAL;
    ElemAbsDiff el{d_Za, d_Zb, d, ld};
    return cascade_launch(el, (int64_t)n * d, (float*)d_ws, ws_bytes, d_out, nullptr, nullptr, 0, (cudaStream_t)s);
}

int clane_patience_reset(clane_patience* d_state, int32_t tol, int32_t max_sweeps, clane_stream_t s) {
    if (!d_state) return CLANE_EINVAL;
    k_patience_reset<<<1, 1, 0, (cudaStream_t)s>>>(d_state, tol, max_sweeps);
    CLANE_LAUNCH_CHECK();
    return CLANE_OK;
}

}  // extern "C"
