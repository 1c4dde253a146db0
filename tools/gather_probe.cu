// Microbenchmark (not product code): how fast can sm_100a gather random 512-byte rows?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe gather_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <cmath>
#include <cuda_runtime.h>

template <int U>
__global__ void k_gather(const float4* __restrict__ Z, const int* __restrict__ idx, int n_idx, float4* __restrict__ out,
                         int ld4) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int base = warp * U; base + U <= n_idx; base += nwarps * U) {
        float4 z[U];
        int r[U];
#pragma unroll
        for (int i = 0; i < U; ++i) r[i] = __ldg(idx + base + i);
#pragma unroll
        for (int i = 0; i < U; ++i) z[i] = __ldg(Z + (size_t)r[i] * ld4 + lane);
#pragma unroll
        for (int i = 0; i < U; ++i) { acc.x += z[i].x; acc.y += z[i].y; acc.z += z[i].z; acc.w += z[i].w; }
    }
    out[(size_t)warp * 32 + lane] = acc;
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 169343, D4 = 32, E = 1166243;
    std::mt19937 rng(1);
    std::vector<int> uni(E), pl(E);
    std::uniform_real_distribution<double> u01(0, 1);
    std::vector<int> perm(N);
    for (int i = 0; i < N; ++i) perm[i] = i;
    for (int i = N - 1; i > 0; --i) std::swap(perm[i], perm[rng() % (i + 1)]);
    for (int e = 0; e < E; ++e) { uni[e] = rng() % N; pl[e] = perm[(int)(N * std::pow(u01(rng), 2.5)) % N]; }
    float4 *Z, *out; int* idx;
    cudaMalloc(&Z, (size_t)N * D4 * 16); cudaMemset(Z, 0, (size_t)N * D4 * 16);
    cudaMalloc(&out, (size_t)1 << 26); cudaMalloc(&idx, (size_t)E * 4);
    float4* flush; cudaMalloc(&flush, 256u << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int dist = 0; dist < 2; ++dist) {
        cudaMemcpy(idx, dist ? pl.data() : uni.data(), (size_t)E * 4, cudaMemcpyHostToDevice);
        for (int wpsm : {16, 24, 32, 64}) {
            auto run = [&](auto kern, int U) {
                const int blocks = 148 * wpsm / 8;   // 256-thread blocks
                float best = 1e9;
                for (int it = 0; it < 6; ++it) {
                    if (it % 2 == 0) cudaMemset(flush, it, 256u << 20);   // cold L2 on even iterations
                    cudaEventRecord(e0);
                    kern<<<blocks, 256>>>(Z, idx, E, out, D4);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    if (it >= 2) best = ms < best ? ms : best;
                }
                printf("dist=%s warps/SM=%2d U=%2d  %.1f us  %.2f TB/s gathered  %.2f Gedges/s\n", dist ? "powerlaw" : "uniform ", wpsm, U,
                       best * 1e3, (double)E * 512 / (best * 1e-3) / 1e12, E / (best * 1e-3) / 1e9);
            };
            run(k_gather<8>, 8);
        }
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
