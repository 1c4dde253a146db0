// Shared device/host helpers for libclane_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "clane_b200.h"

#define CLANE_CUDA(x)                                   \
    do {                                                \
        cudaError_t err__ = (x);                        \
        if (err__ != cudaSuccess) return (int)err__;    \
    } while (0)

#define CLANE_LAUNCH_CHECK() CLANE_CUDA(cudaGetLastError())

namespace clane {

constexpr unsigned kFull = 0xffffffffu;

// ---- exactly rounded fp32 primitives (never contracted; -fmad=false is also set) ----------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// streaming (evict-first) 128-bit load/store for arrays touched once per sweep (X, Znext)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream4(float* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- ATen cascade-sum shape -------------------------------------------
// n flat elements -> nv 8-vectors -> ni cascade rows of 32; 4 levels with level_step = 2^p.
struct CascadeShape {
    int64_t n, nv, ni;
    int p;
    int64_t step;        // rows per level-0 chunk
    int64_t node1_rows;  // step^2 rows per level-1 node
    int64_t n1_full;     // complete level-1 nodes
    int64_t rem_rows;    // rows in the trailing, incomplete level-1 node
    int64_t c_rem;       // complete chunks in it
    int64_t r_rem;       // leftover rows after them (stay in acc[0])
    int64_t n1_nodes;    // n1_full + (rem_rows > 0)
    int64_t n2_full;     // complete level-2 nodes
};

__host__ __device__ inline int ceil_log2_i64(int64_t x) {
    if (x <= 1) return 0;
    int l = 0;
    int64_t v = x - 1;
    while (v > 0) { v >>= 1; ++l; }
    return l;
}

__host__ __device__ inline CascadeShape cascade_shape(int64_t n) {
    CascadeShape s;
    s.n = n;
    s.nv = n / 8;
    s.ni = s.nv / 4;
    int p = ceil_log2_i64(s.ni) / 4;
    s.p = p < 4 ? 4 : p;
    s.step = (int64_t)1 << s.p;
    s.node1_rows = s.step * s.step;
    s.n1_full = s.ni / s.node1_rows;
    s.rem_rows = s.ni - s.n1_full * s.node1_rows;
    s.c_rem = s.rem_rows / s.step;
    s.r_rem = s.rem_rows - s.c_rem * s.step;
    s.n1_nodes = s.n1_full + (s.rem_rows > 0 ? 1 : 0);
    s.n2_full = s.n1_full / s.step;
    return s;
}

constexpr int kStashFloats = 4096;  // |delta| of one group (one level-0 chunk: <= 32 * 2^7 floats)

// floats of the two scratch regions for a cascade over n elements
__host__ inline size_t cascade_p1_floats(int64_t n, int nq) {
    CascadeShape s = cascade_shape(n);
    return (size_t)(s.n1_nodes + 2) * 32 * nq;
}
__host__ inline size_t cascade_p2_floats(int64_t n, int nq) {
    CascadeShape s = cascade_shape(n);
    return (size_t)(s.n2_full + 1) * 32 * nq;
}

}  // namespace clane
