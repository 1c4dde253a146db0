// Microbenchmark (not product code): cp.async ring streaming of random 128-byte row pieces, one consumer warp.
#include <cstdio>
#include <vector>
#include <random>
#include <cuda_runtime.h>

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// MODE 0: cp.async ring (like the hub kernel); MODE 1: LDG + STS ring
template <int STAGES, int MODE>
__global__ void __launch_bounds__(256) k_ring(const float* __restrict__ Z, const int* __restrict__ idx, int k, float* out, int ld) {
    extern __shared__ __align__(16) float ring[];   // [STAGES][32][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int* my = idx + (size_t)blockIdx.x * k;
    const int nst = k / 32;
    const int nb = warp * 4 + (lane >> 3), pc = (lane & 7) * 4;
    float acc = 0.f;
    if (MODE == 0) {
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < nst) cp_async16(ring + (s % STAGES) * 1024 + nb * 32 + pc, Z + (size_t)__ldg(my + s * 32 + nb) * ld + pc);
            cp_async_commit();
        }
        for (int s = 0; s < nst; ++s) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            const int si = s + STAGES - 1;
            if (si < nst) cp_async16(ring + (si % STAGES) * 1024 + nb * 32 + pc, Z + (size_t)__ldg(my + si * 32 + nb) * ld + pc);
            cp_async_commit();
            if (warp == 0) {
                const float* src = ring + (s % STAGES) * 1024 + lane;
#pragma unroll
                for (int i = 0; i < 32; ++i) acc += src[i * 32];
            }
        }
    } else if (MODE >= 2) {
        // like MODE 0 but the row ids are prefetched 8 stages ahead into statically indexed registers
        int cq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) cq[i] = __ldg(my + i * 32 + nb);
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            cp_async16(ring + (s % STAGES) * 1024 + nb * 32 + pc, Z + (size_t)cq[s % 8] * ld + pc);
            cp_async_commit();
            cq[s % 8] = (s + 8) * 32 + nb < k ? __ldg(my + (s + 8) * 32 + nb) : 0;
        }
        for (int sb = 0; sb < nst; sb += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int s = sb + j;
                if (MODE != 5) { cp_async_wait<STAGES - 2>(); __syncthreads(); }
                const int si = s + STAGES - 1;
                const int slot = (j + STAGES - 1) % 8;
                if (si < nst && MODE != 4) cp_async16(ring + (si % STAGES) * 1024 + nb * 32 + pc, Z + (size_t)cq[slot] * ld + pc);
                cp_async_commit();
                cq[slot] = (si + 8) * 32 + nb < k ? __ldg(my + (si + 8) * 32 + nb) : 0;
                if (warp == 0 && MODE != 3) {
                    const float* src = ring + (s % STAGES) * 1024 + lane;
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc += src[i * 32];
                }
            }
        }
    } else {
        float4 v[2];
        // software pipeline of depth 2 stages in registers, ring of STAGES in smem
        for (int s = 0; s < nst + 2; ++s) {
            float4 nv = make_float4(0, 0, 0, 0);
            if (s < nst) nv = __ldg(reinterpret_cast<const float4*>(Z + (size_t)__ldg(my + s * 32 + nb) * ld + pc));
            if (s >= 2) {
                *reinterpret_cast<float4*>(ring + ((s - 2) % STAGES) * 1024 + nb * 32 + pc) = v[s & 1];
                __syncthreads();
                if (warp == 0) {
                    const float* src = ring + ((s - 2) % STAGES) * 1024 + lane;
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc += src[i * 32];
                }
            }
            v[s & 1] = nv;
        }
    }
    if (warp == 0) out[blockIdx.x * 32 + lane] = acc;
}

int main() {
    const int N = 169343, LD = 128, K = 8192;
    std::mt19937 rng(1);
    float *Z, *out; int* idx;
    cudaMalloc(&Z, (size_t)N * LD * 4); cudaMemset(Z, 0, (size_t)N * LD * 4);
    cudaMalloc(&out, 1 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ctas : {1, 148}) {
        std::vector<int> h((size_t)ctas * K);
        for (auto& x : h) x = rng() % N;
        cudaMalloc(&idx, h.size() * 4); cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        auto run = [&](auto kern, int stages, const char* name) {
            size_t smem = (size_t)stages * 4096;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            float best = 1e9;
            for (int it = 0; it < 5; ++it) {
                cudaEventRecord(e0); kern<<<ctas, 256, smem>>>(Z, idx, K, out, LD); cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (it) best = ms < best ? ms : best;
            }
            printf("ctas=%3d %-18s stages=%2d  %.1f us total, %.0f ns per 32-neighbour stage\n", ctas, name, stages, best * 1e3, best * 1e6 / (K / 32));
        };

        run(k_ring<16, 2>, 16, "full");
        run(k_ring<16, 3>, 16, "no consumer");
        run(k_ring<16, 4>, 16, "no copies");
        run(k_ring<16, 5>, 16, "no wait/sync");
        cudaFree(idx);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
