// sortnet.cuh -- register-resident bitonic sorting network shared by the rank pipeline (mbd.cu, u32 keys)
// and the sign-vector matcher (bd_match.cu, u64 keys).
#pragma once
#include "common.cuh"

namespace sd {

template <typename K>
__device__ __forceinline__ void ce_key(K &a, K &b) {
    const K lo = min(a, b), hi = max(a, b);
    a = lo;
    b = hi;
}

// Register-resident bitonic network over 32*EPL keys held as v[i] at element index e = lane*EPL + i,
// in its "always ascending" form: the first stage of every merge pairs e with its mirror e ^ (k-1), the
// remaining stages pair e with e ^ j, and every compare-exchange puts the minimum at the lower index.
// Stages inside a lane are unrolled on registers; stages across lanes are ROLLED loops over the lane
// mask -- the fully unrolled network was ~35k SASS instructions and the rank kernel spent 70% of its
// stall samples in stall_no_inst (profiles/ncu_mbd_r01a_summary.md): code size matters more than loop
// overhead here.
template <int EPL, typename K>
__device__ __forceinline__ void lane_tail(K (&v)[EPL]) {  // xor stages j = EPL/2 .. 1
#pragma unroll
    for (int j = EPL >> 1; j > 0; j >>= 1) {
#pragma unroll
        for (int i = 0; i < EPL; ++i)
            if ((i & j) == 0) ce_key(v[i], v[i | j]);
    }
}

template <int EPL, typename K>
__device__ __forceinline__ void lane_sort(K (&v)[EPL]) {  // every lane sorts its own EPL keys
#pragma unroll
    for (int k = 2; k <= EPL; k <<= 1) {
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const int p = i ^ (k - 1);
            if (i < p) ce_key(v[i], v[p]);
        }
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < EPL; ++i)
                if ((i & j) == 0) ce_key(v[i], v[i | j]);
        }
    }
}

// xor stages with lane masks jl_first, jl_first/2, .., 1 followed by the in-lane tail
template <int EPL, typename K>
__device__ __forceinline__ void warp_merge_tail(K (&v)[EPL], const int lane, const int jl_first) {
#pragma unroll 1
    for (int jl = jl_first; jl > 0; jl >>= 1) {
        const bool lower = (lane & jl) == 0;
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const K o = __shfl_xor_sync(0xffffffffu, v[i], jl);
            v[i] = lower ? min(v[i], o) : max(v[i], o);
        }
    }
    lane_tail<EPL, K>(v);
}

template <int EPL, typename K>
__device__ __forceinline__ void warp_bitonic_sort(K (&v)[EPL], const int lane) {
    lane_sort<EPL, K>(v);
#pragma unroll 1
    for (int kl = 2; kl <= 32; kl <<= 1) {  // merges across kl lanes
        {   // mirror stage: partner lane = lane ^ (kl-1), partner register = EPL-1-i
            const bool lower = (lane & (kl >> 1)) == 0;
            K o[EPL];
#pragma unroll
            for (int i = 0; i < EPL; ++i) o[i] = __shfl_xor_sync(0xffffffffu, v[EPL - 1 - i], kl - 1);
#pragma unroll
            for (int i = 0; i < EPL; ++i) v[i] = lower ? min(v[i], o[i]) : max(v[i], o[i]);
        }
        warp_merge_tail<EPL, K>(v, lane, kl >> 2);
    }
}

// 16 keys of ONE thread, 60 compare-exchanges in 10 layers (a minimal-size 16-input network; checked with the
// 0-1 principle over all 65536 inputs).  The sub-bin ranking of mbd.cu sorts its bins of <= 16 keys with it:
// every compare-exchange works on the thread's own registers, so there are no shuffles and no selects.
template <typename K>
__device__ __forceinline__ void thread_sort16(K (&w)[16]) {
#define SD_CE(a, b) ce_key(w[a], w[b])
    SD_CE(0, 13); SD_CE(1, 12); SD_CE(2, 15); SD_CE(3, 14); SD_CE(4, 8); SD_CE(5, 6); SD_CE(7, 11); SD_CE(9, 10);
    SD_CE(0, 5); SD_CE(1, 7); SD_CE(2, 9); SD_CE(3, 4); SD_CE(6, 13); SD_CE(8, 14); SD_CE(10, 15); SD_CE(11, 12);
    SD_CE(0, 1); SD_CE(2, 3); SD_CE(4, 5); SD_CE(6, 8); SD_CE(7, 9); SD_CE(10, 11); SD_CE(12, 13); SD_CE(14, 15);
    SD_CE(0, 2); SD_CE(1, 3); SD_CE(4, 10); SD_CE(5, 11); SD_CE(6, 7); SD_CE(8, 9); SD_CE(12, 14); SD_CE(13, 15);
    SD_CE(1, 2); SD_CE(3, 12); SD_CE(4, 6); SD_CE(5, 7); SD_CE(8, 10); SD_CE(9, 11); SD_CE(13, 14);
    SD_CE(1, 4); SD_CE(2, 6); SD_CE(5, 8); SD_CE(7, 10); SD_CE(9, 13); SD_CE(11, 14);
    SD_CE(2, 4); SD_CE(3, 6); SD_CE(9, 12); SD_CE(11, 13);
    SD_CE(3, 5); SD_CE(6, 8); SD_CE(7, 9); SD_CE(10, 12);
    SD_CE(3, 4); SD_CE(5, 6); SD_CE(7, 8); SD_CE(9, 10); SD_CE(11, 12);
    SD_CE(6, 7); SD_CE(8, 9);
#undef SD_CE
}

}  // namespace sd
