// bd_bits.cu -- strict band depth (relax=False) as bit-packed sign masks + early-exit pair tests.
//
// Replaces the strict branch of _r2_containment (_containment.py:68-80: containment // len(curve))
// enumerated over all pairs by _univariate_band_depth (_functional.py:238-253).  For query curve q
//     Sb[c][t] = [X[t,c] < X[t,q]],  Sa[c][t] = [X[t,c] > X[t,q]]
// and a pair (c1 < c2) of OTHER curves contains q at every time point iff
//     (Sb[c1] & Sb[c2]) | (Sa[c1] & Sa[c2]) == 0      over all T bits
// i.e. iff the violation Gram entry V[c1,c2] = Sb.Sb^T + Sa.Sa^T is zero ("K2").
// This file is the CUDA-core variant: masks are packed 32 time points per word, word-major so the
// first word of every curve (which rejects almost all pairs) is contiguous; survivors of word 0
// are queued in shared memory and verified on the remaining words by the whole CTA.
// The tcgen05 int8 Gram variant lives in bd_gemm.cu; sd_set_option(SD_OPT_BD_IMPL) selects.
#include "common.cuh"

namespace sd {

constexpr int MASK_QT = 8;     // queries handled per thread by the mask kernel
constexpr int PAIR_TJ = 256;   // curves per tile (threads per CTA)
constexpr int PAIR_QCAP = 4096;

// M[(q*W + w)*n + c] = {below bits, above bits} of curve c vs query q over time points 32w..32w+31
__global__ void __launch_bounds__(128) bd_mask_kernel(const double *__restrict__ X, const i64 T, const i64 n,
                                                      const i64 ld, const i64 *__restrict__ qidx, const int nqb,
                                                      const int W, uint2 *__restrict__ M,
                                                      int *__restrict__ status) {
    __shared__ double sq[32][MASK_QT];
    const int w = blockIdx.y;
    const int q0 = blockIdx.z * MASK_QT;
    for (int i = threadIdx.x; i < 32 * MASK_QT; i += blockDim.x) {
        const int tt = i / MASK_QT, qq = i % MASK_QT;
        const i64 t = (i64)w * 32 + tt;
        double v = 0.0;
        if (t < T && q0 + qq < nqb) v = X[t * ld + qidx[q0 + qq]];
        sq[tt][qq] = v;
    }
    __syncthreads();
    const i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    u32 b[MASK_QT], a[MASK_QT];
#pragma unroll
    for (int qq = 0; qq < MASK_QT; ++qq) b[qq] = a[qq] = 0u;
    bool bad = false;
    const int tmax = (T - (i64)w * 32) < 32 ? (int)(T - (i64)w * 32) : 32;
    for (int tt = 0; tt < tmax; ++tt) {
        const double x = X[((i64)w * 32 + tt) * ld + c];
        bad |= !isfinite(x);
#pragma unroll
        for (int qq = 0; qq < MASK_QT; ++qq) {
            const double xq = sq[tt][qq];
            b[qq] |= (u32)(x < xq) << tt;
            a[qq] |= (u32)(x > xq) << tt;
        }
    }
    if (bad) atomicOr(status, ST_NONFINITE);
#pragma unroll
    for (int qq = 0; qq < MASK_QT; ++qq)
        if (q0 + qq < nqb) M[((i64)(q0 + qq) * W + w) * n + c] = make_uint2(b[qq], a[qq]);
}

__device__ __forceinline__ bool pair_survives_tail(const uint2 *__restrict__ Mq, const int W, const i64 n, const int c1,
                                                   const int c2) {
    for (int w = 1; w < W; ++w) {
        const uint2 m1 = Mq[(i64)w * n + c1], m2 = Mq[(i64)w * n + c2];
        if ((m1.x & m2.x) | (m1.y & m2.y)) return false;
    }
    return true;
}

// grid (tiles, queries).  CTA (jt, q): curves c1 in tile jt against every c2 > c1.
__global__ void __launch_bounds__(PAIR_TJ) bd_pair_kernel(const uint2 *__restrict__ M, const i64 n, const int W,
                                                          const i64 *__restrict__ qidx, i64 *__restrict__ out) {
    __shared__ uint2 s_tile[PAIR_TJ];
    __shared__ uint2 s_queue[PAIR_QCAP];
    __shared__ int s_qn;
    __shared__ u64 s_total;
    const int q = blockIdx.y;
    const int jt = blockIdx.x;
    const int ntiles = (int)ceil_div(n, PAIR_TJ);
    const int qi = (int)qidx[q];
    const uint2 *Mq = M + (i64)q * W * n;
    const int c1 = jt * PAIR_TJ + threadIdx.x;
    const bool valid1 = c1 < n && c1 != qi;
    uint2 m1 = make_uint2(0u, 0u);
    if (valid1) m1 = Mq[c1];
    if (threadIdx.x == 0) { s_qn = 0; s_total = 0ull; }
    u64 count = 0;
    for (int kt = jt; kt < ntiles; ++kt) {
        __syncthreads();  // previous tile fully consumed (and queue drained state visible)
        const int k0 = kt * PAIR_TJ;
        const int klen = (n - k0) < PAIR_TJ ? (int)(n - k0) : PAIR_TJ;
        if ((int)threadIdx.x < klen) s_tile[threadIdx.x] = Mq[k0 + threadIdx.x];
        __syncthreads();
        for (int kk = 0; kk < klen; kk += 32) {
            const int kend = kk + 32 < klen ? kk + 32 : klen;
            if (valid1) {
                for (int u = kk; u < kend; ++u) {
                    const uint2 m2 = s_tile[u];
                    const int c2 = k0 + u;
                    const u32 viol = (m1.x & m2.x) | (m1.y & m2.y);
                    if (viol == 0u && c2 > c1 && c2 != qi) {
                        if (W == 1) {
                            ++count;
                        } else {
                            const int pos = atomicAdd(&s_qn, 1);
                            if (pos < PAIR_QCAP) s_queue[pos] = make_uint2((u32)c1, (u32)c2);
                            else count += pair_survives_tail(Mq, W, n, c1, c2);  // queue full: verify in place
                        }
                    }
                }
            }
            if (W > 1) {
                __syncthreads();
                const int qn = s_qn < PAIR_QCAP ? s_qn : PAIR_QCAP;
                __syncthreads();  // every thread has read the same qn before anyone pushes again
                if (qn >= PAIR_QCAP / 2) {  // uniform decision: drain with the whole CTA
                    for (int i = threadIdx.x; i < qn; i += PAIR_TJ) {
                        const uint2 pr = s_queue[i];
                        count += pair_survives_tail(Mq, W, n, (int)pr.x, (int)pr.y);
                    }
                    __syncthreads();
                    if (threadIdx.x == 0) s_qn = 0;
                    __syncthreads();
                }
            }
        }
    }
    __syncthreads();
    if (W > 1) {
        const int qn = s_qn < PAIR_QCAP ? s_qn : PAIR_QCAP;
        for (int i = threadIdx.x; i < qn; i += PAIR_TJ) {
            const uint2 pr = s_queue[i];
            count += pair_survives_tail(Mq, W, n, (int)pr.x, (int)pr.y);
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) count += __shfl_xor_sync(0xffffffffu, count, s);
    if ((threadIdx.x & 31) == 0 && count) atomicAdd(&s_total, count);
    __syncthreads();
    if (threadIdx.x == 0 && s_total) atomicAdd((u64 *)&out[q], s_total);
}

// J = 3, strict: triples (c1 < c2 < c3) of other curves never all-below / all-above.  O(n^3 W)
// per query: meant for the small n the reference itself can handle.  One thread per (c1, c2).
__global__ void __launch_bounds__(256) bd_triple_kernel(const uint2 *__restrict__ M, const i64 n, const int W,
                                                        const i64 *__restrict__ qidx, i64 *__restrict__ out) {
    const int q = blockIdx.y;
    const int qi = (int)qidx[q];
    const uint2 *Mq = M + (i64)q * W * n;
    const i64 npairs = n * (n - 1) / 2;
    const i64 pid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 count = 0;
    if (pid < npairs) {
        // unrank pid -> (c1 < c2): c2 = largest with c2(c2-1)/2 <= pid
        i64 c2 = (i64)((1.0 + sqrt(1.0 + 8.0 * (double)pid)) * 0.5);
        while (c2 * (c2 - 1) / 2 > pid) --c2;
        while ((c2 + 1) * c2 / 2 <= pid) ++c2;
        const i64 c1 = pid - c2 * (c2 - 1) / 2;
        if (c1 != qi && c2 != qi) {
            for (i64 c3 = c2 + 1; c3 < n; ++c3) {
                if (c3 == qi) continue;
                bool ok = true;
                for (int w = 0; w < W && ok; ++w) {
                    const uint2 a = Mq[(i64)w * n + c1], b = Mq[(i64)w * n + c2], c = Mq[(i64)w * n + c3];
                    ok = ((a.x & b.x & c.x) | (a.y & b.y & c.y)) == 0u;
                }
                count += ok;
            }
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) count += __shfl_xor_sync(0xffffffffu, count, s);
    if ((threadIdx.x & 31) == 0 && count) atomicAdd((u64 *)&out[q], count);
}

__global__ void iota_i64_kernel(i64 *p, i64 count) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = i;
}

// d_q == nullptr means all curves.  d_out[nq] receives the strict numerator for subset size j.
int bd_strict_device(sd_ctx *ctx, const double *dX, i64 T, i64 n, i64 ld, const i64 *d_q, i64 nq, int j,
                     i64 *d_out) {
    if (n < 1 || T < 1 || ld < n || nq < 0) {
        set_error("strict band depth: bad shape T=%lld n=%lld ld=%lld nq=%lld", (long long)T, (long long)n,
                  (long long)ld, (long long)nq);
        return SD_ERR_INVALID;
    }
    if (n >= (1ll << 31)) {
        set_error("strict band depth: n=%lld exceeds 2^31-1", (long long)n);
        return SD_ERR_UNSUPPORTED;
    }
    cudaStream_t st = ctx->stream;
    SD_CUDA(cudaMemsetAsync(d_out, 0, (size_t)nq * sizeof(i64), st));
    if (nq == 0) return SD_OK;
    if (!d_q) {
        SD_TRY(ctx->buf[BUF_QIDX].reserve((size_t)nq * sizeof(i64)));
        i64 *iq = ctx->buf[BUF_QIDX].as<i64>();
        iota_i64_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(iq, nq);
        ctx->last.launches++;
        d_q = iq;
    }
    const i64 W = ceil_div(T, 32);
    if (W > 65535) {
        set_error("strict band depth: T=%lld too large (max %d)", (long long)T, 65535 * 32);
        return SD_ERR_UNSUPPORTED;
    }
    // queries per batch: masks are W*n*8 bytes per query, keep the batch under ~2 GB
    const size_t per_q = (size_t)W * (size_t)n * sizeof(uint2);
    i64 QB = (i64)((2ull << 30) / per_q);
    if (QB < MASK_QT) QB = MASK_QT;
    if (QB > 32768) QB = 32768;
    if (QB > nq) QB = nq;
    SD_TRY(ctx->buf[BUF_MASK].reserve((size_t)QB * per_q));
    uint2 *M = ctx->buf[BUF_MASK].as<uint2>();
    for (i64 q0 = 0; q0 < nq; q0 += QB) {
        const int nqb = (int)(nq - q0 < QB ? nq - q0 : QB);
        dim3 mgrid((unsigned)ceil_div(n, 128), (unsigned)W, (unsigned)ceil_div(nqb, MASK_QT));
        SD_TRY(prof_begin(ctx, SD_PHASE_BD_MASKS));
        bd_mask_kernel<<<mgrid, 128, 0, st>>>(dX, T, n, ld, d_q + q0, nqb, (int)W, M, ctx->d_status);
        SD_TRY(prof_end(ctx));
        ctx->last.launches++;
        SD_TRY(prof_begin(ctx, SD_PHASE_BD_PAIRS));
        if (j == 2) {
            dim3 pgrid((unsigned)ceil_div(n, PAIR_TJ), (unsigned)nqb);
            bd_pair_kernel<<<pgrid, PAIR_TJ, 0, st>>>(M, n, (int)W, d_q + q0, d_out + q0);
        } else {
            const i64 npairs = n * (n - 1) / 2;
            if (npairs == 0) { SD_TRY(prof_end(ctx)); continue; }
            dim3 tgrid((unsigned)ceil_div(npairs, 256), (unsigned)nqb);
            bd_triple_kernel<<<tgrid, 256, 0, st>>>(M, n, (int)W, d_q + q0, d_out + q0);
        }
        SD_TRY(prof_end(ctx));
        ctx->last.launches++;
        SD_CUDA(cudaGetLastError());
    }
    return SD_OK;
}

}  // namespace sd
