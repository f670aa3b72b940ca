// dsmem_probe.cu -- can a thread-block cluster hold one time row on chip and exchange it by SCATTERED remote stores?
// (round 2, the cluster-per-row variant of the MBD pipeline the round-1 verdict suggested.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_probe.bin dsmem_probe.cu ; ./dsmem_probe.bin
// Clusters of 8 CTAs x 512 threads, one CTA per SM (100 KB of shared memory each).  Every thread performs ITER
// operations on pseudo-random words of a pseudo-random CTA of its cluster; reports SM-cycles per warp instruction
// (per SM) for: 8-byte remote stores, 4-byte remote atomics (red.shared::cluster), 8-byte local stores (baseline),
// and a coalesced remote copy (lane-contiguous 8-byte stores to one remote CTA: what a grouped exchange would do).
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
typedef unsigned int u32;
typedef unsigned long long u64;
constexpr int ITER = 64, CL = 8, THREADS = 512, WORDS = 12800;  // 100 KB of u64 per CTA

__device__ __forceinline__ u32 rnd(u32 &s) { s = s * 1664525u + 1013904223u; return s >> 8; }
__device__ __forceinline__ u32 mapa(u32 saddr, u32 rank) {
    u32 r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}

template <int MODE>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1) probe(u32 *out, long long *cycles) {
    extern __shared__ __align__(16) u64 buf[];
    cg::cluster_group cluster = cg::this_cluster();
    for (int i = threadIdx.x; i < WORDS; i += THREADS) buf[i] = 0;
    cluster.sync();
    const u32 base = (u32)__cvta_generic_to_shared(buf);
    u32 s = (blockIdx.x * THREADS + threadIdx.x) * 2654435761u + 12345u;
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const u32 r = rnd(s);
            const u32 w = (r >> 3) % WORDS, cta = r & 7u;
            if (MODE == 0) {  // scattered remote 8-byte store
                const u32 a = mapa(base + w * 8u, cta);
                asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(a), "l"((u64)r) : "memory");
            } else if (MODE == 1) {  // scattered remote 4-byte atomic add (no return)
                const u32 a = mapa(base + w * 8u, cta);
                asm volatile("red.shared::cluster.add.u32 [%0], %1;" ::"r"(a), "r"(1u) : "memory");
            } else if (MODE == 2) {  // scattered LOCAL 8-byte store
                buf[w] = r;
            } else if (MODE == 3) {  // lane-contiguous remote 8-byte stores: one remote CTA per warp instruction
                const u32 r0 = __shfl_sync(0xffffffffu, r, 0);
                const u32 w0 = ((r0 >> 3) % (WORDS - 32)) + lane;
                const u32 a = mapa(base + w0 * 8u, r0 & 7u);
                asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(a), "l"((u64)r) : "memory");
            } else if (MODE == 4) {  // scattered remote 4-byte atomic WITH return
                const u32 a = mapa(base + w * 8u, cta);
                u32 old;
                asm volatile("atom.shared::cluster.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(1u) : "memory");
                s += old;
            }
        }
    }
    cluster.sync();
    long long t1 = clock64();
    if (threadIdx.x == 0) atomicMax((unsigned long long *)cycles, (unsigned long long)(t1 - t0));
    out[blockIdx.x * THREADS + threadIdx.x] = s + (u32)buf[threadIdx.x];
}

template <int MODE>
void run(const char *name, u32 *out, long long *cyc) {
    const int grid = 18 * CL;  // 18 clusters of 8 on 148 SMs
    const size_t smem = WORDS * 8;
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<MODE><<<grid, THREADS, smem>>>(out, cyc);
    cudaMemset(cyc, 0, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<grid, THREADS, smem>>>(out, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double winstr = (double)(THREADS / 32) * ITER * 8;  // warp instructions per SM
    printf("%-58s %.3f ms  clock64 max %lld -> %.2f SM-cycles per warp-op (%s)\n", name, ms, h, (double)h / winstr,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    u32 *out; long long *cyc;
    cudaMalloc(&out, 18 * CL * THREADS * 4); cudaMalloc(&cyc, 8);
    run<2>("scattered LOCAL 8-byte store", out, cyc);
    run<0>("scattered REMOTE 8-byte store (st.shared::cluster)", out, cyc);
    run<3>("lane-contiguous REMOTE 8-byte stores, one CTA per warp-op", out, cyc);
    run<1>("scattered REMOTE 4-byte red.shared::cluster.add", out, cyc);
    run<4>("scattered REMOTE 4-byte atom.shared::cluster.add (return)", out, cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
