// smem_probe.cu -- per-SM throughput of the shared-memory operations the MBD partition / rank kernels lean on
// (round 2).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_probe smem_probe.cu ; ./smem_probe
// Every test: 148*4 CTAs x 256 threads, each warp issues ITER x 16 operations on pseudo-random addresses of a
// table of NADDR words; reports SM cycles per warp-instruction (per SM, all resident warps together).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned int u32;
constexpr int ITER = 256;

__device__ __forceinline__ u32 rnd(u32 &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(256) probe(u32 *out, int naddr_mask, long long *cycles) {
    __shared__ u32 tab[4096];
    for (int i = threadIdx.x; i < 4096; i += 256) tab[i] = 0;
    __syncthreads();
    u32 s = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const u32 a = rnd(s) & naddr_mask;
            if (MODE == 0) acc += atomicAdd(&tab[a], 1u);                  // ATOMS with return
            else if (MODE == 1) atomicAdd(&tab[a], 1u);                    // ATOMS, result unused
            else if (MODE == 2) acc += tab[a];                             // LDS random
            else if (MODE == 3) tab[a] = acc + k;                          // STS random
            else if (MODE == 4) {                                          // match.any + popc ranking
                const u32 m = __match_any_sync(0xffffffffu, a);
                acc += __popc(m & ((1u << lane) - 1u)) + (u32)__popc(m);
            } else if (MODE == 5) {                                        // 8 ballots emulate match.any (256 ids)
                u32 m = 0xffffffffu;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const u32 bal = __ballot_sync(0xffffffffu, (a >> b) & 1u);
                    m &= ((a >> b) & 1u) ? bal : ~bal;
                }
                acc += __popc(m & ((1u << lane) - 1u)) + (u32)__popc(m);
            } else if (MODE == 6) {                                        // aggregated: match + leader RMW + shfl
                const u32 m = __match_any_sync(0xffffffffu, a);
                const int leader = __ffs(m) - 1;
                u32 base = 0;
                if (lane == leader) { base = tab[a]; tab[a] = base + __popc(m); }
                base = __shfl_sync(0xffffffffu, base, leader);
                acc += base + __popc(m & ((1u << lane) - 1u));
                __syncwarp();
            } else if (MODE == 7) {                                        // red.shared via inline asm (no return)
                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"((u32)__cvta_generic_to_shared(&tab[a])) : "memory");
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) atomicMax((unsigned long long *)cycles, (unsigned long long)(t1 - t0));
    out[blockIdx.x * 256 + threadIdx.x] = acc + tab[threadIdx.x];
}

template <int MODE>
void run(const char *name, int naddr, u32 *out, long long *cyc) {
    const int ctas_per_sm = 4, grid = 148 * ctas_per_sm;
    cudaMemset(cyc, 0, 8);
    probe<MODE><<<grid, 256>>>(out, naddr - 1, cyc);
    cudaMemset(cyc, 0, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<grid, 256>>>(out, naddr - 1, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double winstr_per_sm = (double)ctas_per_sm * 8 * ITER * 16;  // warp-instructions per SM
    printf("%-44s naddr=%4d  %.3f ms  %7.2f SM-cycles per warp-op (clock64 max %lld -> %.2f)\n", name, naddr, ms,
           ms * 1e-3 * 1.965e9 / winstr_per_sm, h, (double)h / winstr_per_sm);
}

int main() {
    u32 *out; long long *cyc;
    cudaMalloc(&out, 148 * 4 * 256 * 4); cudaMalloc(&cyc, 8);
    for (int naddr : {64, 256, 1024, 4096}) {
        run<0>("ATOMS.ADD with return, random", naddr, out, cyc);
        run<1>("atomicAdd result unused, random", naddr, out, cyc);
        run<7>("red.shared.add, random", naddr, out, cyc);
        run<2>("LDS random", naddr, out, cyc);
        run<3>("STS random", naddr, out, cyc);
        run<4>("match.any + 2 popc", naddr, out, cyc);
        run<5>("8 ballots + 2 popc", naddr, out, cyc);
        run<6>("match + leader LDS/STS + shfl (aggregated)", naddr, out, cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
