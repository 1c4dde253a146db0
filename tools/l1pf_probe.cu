// Microbenchmark (not product code): which issue pattern gathers random 512-byte rows fastest on sm_100a?
//   U loads per round into registers, optional prefetch.global.L1 / .L2 of the rows PFD rounds ahead.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l1pf_probe l1pf_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <cmath>
#include <cuda_runtime.h>

template <int U, int PF, int PFD>   // PF: 0 none, 1 L1, 2 L2
__global__ void k_gather(const float4* __restrict__ Z, const int* __restrict__ idx, int n_idx, float4* __restrict__ out,
                         int ld4, int per_warp) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float4 acc = make_float4(0, 0, 0, 0);
    const int lo = warp * per_warp, hi = min(n_idx, lo + per_warp);
    for (int base = lo; base + U <= hi; base += U) {
        if (PF) {
            // lane -> (row lane / 4, 128-byte line lane % 4) of the round PFD ahead
            const int t = base + PFD * U + (lane >> 2);
            if ((lane >> 2) < U && t < hi) {
                const char* a = reinterpret_cast<const char*>(Z + (size_t)__ldg(idx + t) * ld4) + (lane & 3) * 128;
                if (PF == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
                else asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            }
            if (U > 8) {
                const int t2 = t + 8;
                if ((lane >> 2) + 8 < U && t2 < hi) {
                    const char* a = reinterpret_cast<const char*>(Z + (size_t)__ldg(idx + t2) * ld4) + (lane & 3) * 128;
                    if (PF == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
                    else asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                }
            }
        }
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
    std::vector<int> pl(E);
    std::uniform_real_distribution<double> u01(0, 1);
    std::vector<int> perm(N);
    for (int i = 0; i < N; ++i) perm[i] = i;
    for (int i = N - 1; i > 0; --i) std::swap(perm[i], perm[rng() % (i + 1)]);
    for (int e = 0; e < E; ++e) pl[e] = perm[(int)(N * std::pow(u01(rng), 2.5)) % N];
    float4 *Z, *out; int* idx;
    cudaMalloc(&Z, (size_t)N * D4 * 16); cudaMemset(Z, 0, (size_t)N * D4 * 16);
    cudaMalloc(&out, (size_t)1 << 26); cudaMalloc(&idx, (size_t)E * 4);
    float4* flush; cudaMalloc(&flush, 256u << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemcpy(idx, pl.data(), (size_t)E * 4, cudaMemcpyHostToDevice);
    auto run = [&](auto kern, const char* name, int tasks_per_slot) {
        for (int wpsm : {16, 24, 32, 48, 64}) {
            // many more warps than slots (like the product kernel: ~10 waves of short tasks)
            const int slots = 148 * wpsm;
            const int nwarps = slots * tasks_per_slot;
            const int per_warp = (E + nwarps - 1) / nwarps;
            const int blocks = nwarps / 4;
            float best = 1e9;
            for (int it = 0; it < 6; ++it) {
                if (it % 2 == 0) cudaMemset(flush, it, 256u << 20);   // cold L2 on even iterations
                cudaEventRecord(e0);
                kern<<<blocks, 128>>>(Z, idx, E, out, D4, per_warp);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (it >= 2) best = ms < best ? ms : best;
            }
            printf("%-22s warps/SM=%2d edges/warp=%4d  %6.1f us  %5.2f TB/s\n", name, wpsm, per_warp, best * 1e3,
                   (double)E * 512 / (best * 1e-3) / 1e12);
        }
    };
    for (int tps : {1, 8}) {
        printf("--- tasks per warp slot: %d\n", tps);
        run(k_gather<8, 0, 0>, "U=8", tps);
        run(k_gather<16, 0, 0>, "U=16", tps);
        run(k_gather<8, 1, 1>, "U=8 L1pf+1", tps);
        run(k_gather<8, 1, 2>, "U=8 L1pf+2", tps);
        run(k_gather<8, 2, 4>, "U=8 L2pf+4", tps);
        run(k_gather<4, 0, 0>, "U=4", tps);
        run(k_gather<4, 1, 2>, "U=4 L1pf+2", tps);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
