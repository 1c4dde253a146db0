// Microbenchmark (not product code): does gathering HALF of every 512-byte row per pass keep the table L2-resident?
// The arxiv-shape Zcur is 87 MB; the B200's 126 MB L2 is two partitions that each cache what their own SMs touch,
// so random full-row gathers from all SMs see ~half of it.  Variant B reads every row twice, 256 bytes per pass
// (the four 64-byte pieces whose columns fall in cascade lanes 0-15, then 16-31): 43 MB per pass.
//   A: one pass, float4 per lane (full rows)         B: two passes, float2 per lane (half rows)
// Both start from a flushed L2 (like a sweep: Zcur was last written, with streaming stores, one sweep ago).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o colsplit_probe colsplit_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <cmath>
#include <cuda_runtime.h>

template <int U>
__global__ void k_full(const float4* __restrict__ Z, const int* __restrict__ idx, int n_idx, float4* __restrict__ out, int per_warp) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float4 acc = make_float4(0, 0, 0, 0);
    const int lo = warp * per_warp, hi = min(n_idx, lo + per_warp);
    for (int base = lo; base + U <= hi; base += U) {
        float4 z[U];
        int r[U];
#pragma unroll
        for (int i = 0; i < U; ++i) r[i] = __ldg(idx + base + i);
#pragma unroll
        for (int i = 0; i < U; ++i) z[i] = __ldg(Z + (size_t)r[i] * 32 + lane);
#pragma unroll
        for (int i = 0; i < U; ++i) { acc.x += z[i].x; acc.y += z[i].y; acc.z += z[i].z; acc.w += z[i].w; }
    }
    out[(size_t)warp * 32 + lane] = acc;
}

// pass = blockIdx.x / blocks_per_pass; lane L reads 8 bytes: 64-byte piece L / 8 of the row's half `pass`
template <int U>
__global__ void k_half(const float2* __restrict__ Z, const int* __restrict__ idx, int n_idx, float2* __restrict__ out,
                       int per_warp, int blocks_per_pass) {
    const int lane = threadIdx.x & 31;
    const int pass = blockIdx.x / blocks_per_pass;
    const int warp = ((blockIdx.x - pass * blocks_per_pass) * blockDim.x + threadIdx.x) >> 5;
    const int col2 = (lane >> 3) * 16 + pass * 8 + (lane & 7);     // float2 index inside the 64-float2 row
    float2 acc = make_float2(0, 0);
    const int lo = warp * per_warp, hi = min(n_idx, lo + per_warp);
    for (int base = lo; base + U <= hi; base += U) {
        float2 z[U];
        int r[U];
#pragma unroll
        for (int i = 0; i < U; ++i) r[i] = __ldg(idx + base + i);
#pragma unroll
        for (int i = 0; i < U; ++i) z[i] = __ldg(Z + (size_t)r[i] * 64 + col2);
#pragma unroll
        for (int i = 0; i < U; ++i) { acc.x += z[i].x; acc.y += z[i].y; }
    }
    out[((size_t)pass * gridDim.x * 4 + warp) * 32 + lane] = acc;
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 169343, E = 1166243;
    std::mt19937 rng(1);
    std::vector<int> pl(E);
    std::uniform_real_distribution<double> u01(0, 1);
    std::vector<int> perm(N);
    for (int i = 0; i < N; ++i) perm[i] = i;
    for (int i = N - 1; i > 0; --i) std::swap(perm[i], perm[rng() % (i + 1)]);
    for (int e = 0; e < E; ++e) pl[e] = perm[(int)(N * std::pow(u01(rng), 2.5)) % N];
    float4 *Z, *out; int* idx;
    cudaMalloc(&Z, (size_t)N * 512); cudaMemset(Z, 0, (size_t)N * 512);
    cudaMalloc(&out, (size_t)1 << 27); cudaMalloc(&idx, (size_t)E * 4);
    float4* flush; cudaMalloc(&flush, 512u << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemcpy(idx, pl.data(), (size_t)E * 4, cudaMemcpyHostToDevice);
    for (int tps : {1, 4, 8}) {
        for (int wpsm : {16, 24, 32, 48, 64}) {
            const int nwarps = 148 * wpsm * tps;
            const int per_warp = (E + nwarps - 1) / nwarps;
            const int blocks = (nwarps + 3) / 4;
            float tA[2] = {1e9f, 1e9f}, tB[2] = {1e9f, 1e9f};     // [cold, warm]
            for (int it = 0; it < 6; ++it) {
                const int warm = it & 1;
                if (!warm) cudaMemset(flush, it, 512u << 20);
                cudaEventRecord(e0);
                k_full<8><<<blocks, 128>>>(Z, idx, E, out, per_warp);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                tA[warm] = ms < tA[warm] ? ms : tA[warm];
            }
            for (int it = 0; it < 6; ++it) {
                const int warm = it & 1;
                if (!warm) cudaMemset(flush, it, 512u << 20);
                cudaEventRecord(e0);
                k_half<8><<<2 * blocks, 128>>>(reinterpret_cast<const float2*>(Z), idx, E, reinterpret_cast<float2*>(out), per_warp, blocks);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                tB[warm] = ms < tB[warm] ? ms : tB[warm];
            }
            printf("tasks/slot=%d warps/SM=%2d  full rows: cold %6.1f us warm %6.1f us   two half-row passes: cold %6.1f us warm %6.1f us\n",
                   tps, wpsm, tA[0] * 1e3, tA[1] * 1e3, tB[0] * 1e3, tB[1] * 1e3);
        }
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
