// api.cu -- extern "C" entry points taking HOST buffers (copies + kernels + copy back), and the
// device-pointer variant used for HBM-resident timing.  See include/statdepth_b200.h.
#include <thread>
#include <vector>

#include "common.cuh"

namespace sd {

// strict J = 2 by enumeration: bit kernel, or the dense Gram when a probe of 8 queries shows that the bit
// kernel would drown in survivors (non-crossing / tie-heavy curves)
static int strict_enumerate(sd_ctx *ctx, const double *dX, i64 T, i64 n, i64 ld, const i64 *d_q, i64 nq, i64 *d_out) {
    if (ctx->bd_impl == SD_BD_GEMM) {
        ctx->last.bd_impl_used = SD_BD_GEMM;
        return bd_strict_gemm_device(ctx, dX, T, n, ld, d_q, nq, d_out);
    }
    if (ctx->last.bd_impl_used == 0) ctx->last.bd_impl_used = SD_BD_BITS;
    constexpr i64 PROBE = 8;
    if (ctx->bd_impl == SD_BD_BITS || nq <= 4 * PROBE || n < 1024)
        return bd_strict_device(ctx, dX, T, n, ld, d_q, nq, 2, d_out);
    SD_TRY(ctx->buf[BUF_WORK].reserve(sizeof(u64)));
    u64 *d_hits = ctx->buf[BUF_WORK].as<u64>();
    SD_CUDA(cudaMemsetAsync(d_hits, 0, sizeof(u64), ctx->stream));
    i64 *iq = nullptr;
    if (!d_q) {  // materialise the identity so that the probe and the remainder can be offset
        SD_TRY(ctx->buf[BUF_QIDX].reserve((size_t)nq * sizeof(i64) * 2));
        iq = ctx->buf[BUF_QIDX].as<i64>() + nq;
        std::vector<i64> h((size_t)nq);
        for (i64 i = 0; i < nq; ++i) h[(size_t)i] = i;
        SD_CUDA(cudaMemcpyAsync(iq, h.data(), (size_t)nq * sizeof(i64), cudaMemcpyHostToDevice, ctx->stream));
        SD_CUDA(cudaStreamSynchronize(ctx->stream));
        d_q = iq;
    }
    SD_TRY(bd_strict_device(ctx, dX, T, n, ld, d_q, PROBE, 2, d_out, d_hits));
    u64 h_hits = 0;
    SD_CUDA(cudaMemcpyAsync(&h_hits, d_hits, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(cudaStreamSynchronize(ctx->stream));
    const double pairs = (double)PROBE * 0.5 * (double)(n - 1) * (double)(n - 2);
    if ((double)h_hits > 0.02 * pairs) {
        ctx->last.bd_impl_used = SD_BD_GEMM;
        return bd_strict_gemm_device(ctx, dX, T, n, ld, d_q + PROBE, nq - PROBE, d_out + PROBE);
    }
    ctx->last.bd_impl_used = SD_BD_BITS;
    return bd_strict_device(ctx, dX, T, n, ld, d_q + PROBE, nq - PROBE, 2, d_out + PROBE);
}

// dispatch one subset size j on device-resident data
int band_depth_device(sd_ctx *ctx, const double *dX, i64 T, i64 n, i64 ld, const i64 *d_q, i64 nq, int j,
                      int relax, i64 *d_out) {
    if (j != 2 && j != 3) {
        set_error("band depth: subset size j=%d not supported on the device (2 or 3)", j);
        return SD_ERR_UNSUPPORTED;
    }
    if (!relax) {
        if (j == 3) {
            ctx->last.bd_impl_used = SD_BD_BITS;
            return bd_strict_device(ctx, dX, T, n, ld, d_q, nq, 3, d_out);
        }
        ctx->last.bd_impl_used = 0;
        const bool match = (ctx->bd_impl == SD_BD_MATCH || ctx->bd_impl == SD_BD_AUTO) && bd_match_supported(T, n);
        if (!match) return strict_enumerate(ctx, dX, T, n, ld, d_q, nq, d_out);
        i64 nfb = 0;
        SD_TRY(bd_strict_match_device(ctx, dX, T, n, ld, d_q, nq, d_out, strict_enumerate, &nfb));
        if (2 * nfb <= nq) ctx->last.bd_impl_used = SD_BD_MATCH;  // else: what strict_enumerate chose
        return SD_OK;
    }
    // overflow guard: T * C(n-1, j) must fit in int64
    {
        const long double full = (j == 2) ? (long double)(n - 1) * (n - 2) / 2.0L
                                          : (long double)(n - 1) * (n - 2) * (n - 3) / 6.0L;
        // j = 3: the kernels form 3 * C(m, 3) before dividing (comb3_dev), so that intermediate must fit as well
        if (full * (long double)(T > 0 ? T : 1) >= 9.0e18L || (j == 3 && 3.0L * full >= 9.0e18L)) {
            set_error("band depth: T*C(n-1,%d) overflows int64 for T=%lld n=%lld", j, (long long)T, (long long)n);
            return SD_ERR_OVERFLOW;
        }
    }
    SD_TRY(ctx->buf[BUF_ACC].reserve((size_t)n * 2 * sizeof(i64)));
    i64 *acc2 = ctx->buf[BUF_ACC].as<i64>();
    i64 *acc3 = acc2 + n;
    if (!d_q && j == 2)  // all curves, in order: the finish kernel writes the caller's buffer, no gather launch
        return mbd_all_device(ctx, dX, T, n, ld, false, d_out, nullptr, nullptr, nullptr);
    SD_TRY(mbd_all_device(ctx, dX, T, n, ld, j == 3, acc2, acc3, nullptr, nullptr));
    return gather_i64_device(ctx, j == 2 ? acc2 : acc3, d_q, nq, d_out);
}

static int check_common(sd_ctx *ctx, const void *in, const void *out, const char *who) {
    if (!ctx) {
        set_error("%s: NULL context", who);
        return SD_ERR_INVALID;
    }
    if (!in || !out) {
        set_error("%s: NULL buffer", who);
        return SD_ERR_INVALID;
    }
    return SD_OK;
}

// upload helper: host matrix [rows, cols] with leading dimension ld -> dense device [rows, cols]
static int upload_matrix(sd_ctx *ctx, int slot, const double *h, i64 rows, i64 cols, i64 ld, double **d) {
    SD_TRY(ctx->buf[slot].reserve((size_t)(rows * cols > 0 ? rows * cols : 1) * sizeof(double)));
    *d = ctx->buf[slot].as<double>();
    if (rows == 0 || cols == 0) return SD_OK;
    if (ld == cols || rows == 1) {
        SD_CUDA(cudaMemcpyAsync(*d, h, (size_t)rows * cols * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    } else {
        SD_CUDA(cudaMemcpy2DAsync(*d, (size_t)cols * sizeof(double), h, (size_t)ld * sizeof(double),
                                  (size_t)cols * sizeof(double), (size_t)rows, cudaMemcpyHostToDevice,
                                  ctx->stream));
    }
    return SD_OK;
}

static int upload_queries(sd_ctx *ctx, const int64_t *h_q, i64 nq, i64 n, const i64 **d_q, const char *who) {
    *d_q = nullptr;
    if (!h_q) {
        if (nq != n) {
            set_error("%s: query_idx == NULL requires nq == n", who);
            return SD_ERR_INVALID;
        }
        return SD_OK;
    }
    for (i64 i = 0; i < nq; ++i)
        if (h_q[i] < 0 || h_q[i] >= n) {
            set_error("%s: query_idx[%lld]=%lld out of range [0,%lld)", who, (long long)i, (long long)h_q[i],
                      (long long)n);
            return SD_ERR_INVALID;
        }
    SD_TRY(ctx->buf[BUF_MISC].reserve((size_t)(nq > 0 ? nq : 1) * sizeof(i64)));
    i64 *dq = ctx->buf[BUF_MISC].as<i64>();
    if (nq > 0)
        SD_CUDA(cudaMemcpyAsync(dq, h_q, (size_t)nq * sizeof(i64), cudaMemcpyHostToDevice, ctx->stream));
    *d_q = dq;
    return SD_OK;
}

// Pageable host input (what FunctionalDepth([DataFrame]) hands over): cudaMemcpyAsync from unregistered memory is a
// synchronous, single-threaded staged copy (~11 GB/s measured: 74 ms for the 819 MB of config 2 against 15 ms from
// pinned memory).  The blocks are instead copied into two pinned staging buffers by a few host threads while the
// previous block's DMA and ranking run.
static bool host_is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

static void parallel_copy_rows(double *dst, const double *src, i64 rows, i64 n, i64 ld, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    const i64 per = ceil_div(rows, nthreads);
    for (int t = 0; t < nthreads; ++t) {
        const i64 r0 = t * per, r1 = r0 + per < rows ? r0 + per : rows;
        if (r0 >= r1) break;
        th.emplace_back([=]() {
            if (ld == n) memcpy(dst + r0 * n, src + r0 * ld, (size_t)(r1 - r0) * n * sizeof(double));
            else
                for (i64 r = r0; r < r1; ++r) memcpy(dst + r * n, src + r * ld, (size_t)n * sizeof(double));
        });
    }
    for (auto &t : th) t.join();
}

// Relaxed band depth of a host matrix, streamed in row blocks (see sd_band_depth_f64).  Timings: h2d_ns is 0
// and kernel_ns covers the overlapped copy + ranking region.
static int band_depth_pipelined(sd_ctx *ctx, const double *X, i64 T, i64 n, i64 ld, const int64_t *query_idx, i64 nq,
                                int j, int64_t *count_out) {
    {
        const long double full = (j == 2) ? (long double)(n - 1) * (n - 2) / 2.0L
                                          : (long double)(n - 1) * (n - 2) * (n - 3) / 6.0L;
        if (full * (long double)T >= 9.0e18L || (j == 3 && 3.0L * full >= 9.0e18L)) {
            set_error("band depth: T*C(n-1,%d) overflows int64 for T=%lld n=%lld", j, (long long)T, (long long)n);
            return SD_ERR_OVERFLOW;
        }
    }
    i64 RB = ceil_div(T, 8);                       // 8 blocks ...
    const i64 min_rows = ceil_div((i64)(32u << 20), n * (i64)sizeof(double));
    if (RB < min_rows) RB = min_rows;              // ... of at least 32 MB
    if (RB > T) RB = T;
    SD_TRY(ctx->buf[BUF_IN].reserve((size_t)2 * RB * n * sizeof(double)));
    double *dbuf[2] = {ctx->buf[BUF_IN].as<double>(), ctx->buf[BUF_IN].as<double>() + (size_t)RB * n};
    const i64 *d_q = nullptr;
    SD_TRY(upload_queries(ctx, query_idx, nq, n, &d_q, "sd_band_depth_f64"));
    SD_TRY(ctx->buf[BUF_OUT].reserve((size_t)(nq > 0 ? nq : 1) * sizeof(i64)));
    SD_TRY(ctx->buf[BUF_ACC].reserve((size_t)n * 2 * sizeof(i64)));
    i64 *d_out = ctx->buf[BUF_OUT].as<i64>();
    i64 *acc2 = ctx->buf[BUF_ACC].as<i64>(), *acc3 = acc2 + n;
    SD_TRY(mark(ctx, 1));
    // the copy stream must not overtake work already queued on the main stream (query upload, status reset)
    SD_CUDA(cudaEventRecord(ctx->ev_pipe[2], ctx->stream));
    SD_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_pipe[2], 0));
    const bool staged = host_is_pageable(X);
    int nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads > 8) nthreads = 8;
    if (staged) {
        for (int s = 0; s < 2; ++s) {
            if (ctx->stage_cap[s] < (size_t)RB * n * sizeof(double)) {
                if (ctx->stage[s]) cudaFreeHost(ctx->stage[s]);
                ctx->stage[s] = nullptr;
                ctx->stage_cap[s] = 0;
                SD_CUDA(cudaHostAlloc(&ctx->stage[s], (size_t)RB * n * sizeof(double), cudaHostAllocDefault));
                ctx->stage_cap[s] = (size_t)RB * n * sizeof(double);
            }
            if (!ctx->ev_stage[s]) SD_CUDA(cudaEventCreateWithFlags(&ctx->ev_stage[s], cudaEventDisableTiming));
        }
    }
    // an error inside the loop must not return while a copy out of the caller's buffer is still in flight
    ctx->mbd_no_wait = 1;  // the host must keep feeding the copy stream while earlier blocks are ranked
    const int loop_status = [&]() -> int {
        int k = 0;
        for (i64 r0 = 0; r0 < T; r0 += RB, ++k) {
            const int s = k & 1;
            const i64 rows = T - r0 < RB ? T - r0 : RB;
            if (k >= 2) SD_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_pipe[2 + s], 0));  // block k-2 consumed
            if (staged) {
                if (k >= 2) SD_CUDA(cudaEventSynchronize(ctx->ev_stage[s]));  // the DMA out of this staging buffer is done
                parallel_copy_rows((double *)ctx->stage[s], X + r0 * ld, rows, n, ld, nthreads);
                SD_CUDA(cudaMemcpyAsync(dbuf[s], ctx->stage[s], (size_t)rows * n * sizeof(double), cudaMemcpyHostToDevice,
                                        ctx->copy_stream));
                SD_CUDA(cudaEventRecord(ctx->ev_stage[s], ctx->copy_stream));
            } else if (ld == n) {
                SD_CUDA(cudaMemcpyAsync(dbuf[s], X + r0 * ld, (size_t)rows * n * sizeof(double), cudaMemcpyHostToDevice,
                                        ctx->copy_stream));
            } else {
                SD_CUDA(cudaMemcpy2DAsync(dbuf[s], (size_t)n * sizeof(double), X + r0 * ld, (size_t)ld * sizeof(double),
                                          (size_t)n * sizeof(double), (size_t)rows, cudaMemcpyHostToDevice,
                                          ctx->copy_stream));
            }
            SD_CUDA(cudaEventRecord(ctx->ev_pipe[s], ctx->copy_stream));
            SD_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_pipe[s], 0));
            SD_TRY(mbd_all_device(ctx, dbuf[s], rows, n, n, j == 3, acc2, acc3, nullptr, nullptr, k > 0));
            SD_CUDA(cudaEventRecord(ctx->ev_pipe[2 + s], ctx->stream));
        }
        return SD_OK;
    }();
    ctx->mbd_no_wait = 0;
    if (loop_status != SD_OK) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->stream);
        return loop_status;
    }
    SD_TRY(gather_i64_device(ctx, j == 2 ? acc2 : acc3, d_q, nq, d_out));
    SD_TRY(mark(ctx, 2));
    if (nq > 0)
        SD_CUDA(cudaMemcpyAsync(count_out, d_out, (size_t)nq * sizeof(i64), cudaMemcpyDeviceToHost, ctx->stream));
    const int st = end_call(ctx, true);
    ctx->last.h2d_ns = 0;
    return st;
}

}  // namespace sd

using namespace sd;

extern "C" {

static int sd_band_depth_f64_impl(sd_ctx *ctx, const double *X, int64_t T, int64_t n, int64_t ld, int layout,
                      const int64_t *query_idx, int64_t nq, int j, int relax, int64_t *count_out) {
    SD_TRY(check_common(ctx, X, count_out, "sd_band_depth_f64"));
    SD_REQUIRE(T >= 1 && n >= 1 && nq >= 0, "sd_band_depth_f64: bad sizes T=%lld n=%lld nq=%lld", (long long)T,
               (long long)n, (long long)nq);
    SD_REQUIRE(layout == SD_LAYOUT_TN || layout == SD_LAYOUT_NT, "sd_band_depth_f64: bad layout %d", layout);
    SD_REQUIRE(ld >= (layout == SD_LAYOUT_TN ? n : T), "sd_band_depth_f64: ld=%lld too small", (long long)ld);
    SD_TRY(begin_call(ctx));
    // Large relaxed inputs: the H2D copy (PCIe) is ~7x longer than the ranking, and the relaxed numerator is
    // additive over time rows -> stream the matrix in row blocks, copy of block k+1 overlapping ranking of k.
    if (relax && layout == SD_LAYOUT_TN && (j == 2 || j == 3) && T >= 16 && (size_t)T * n * sizeof(double) >= (64u << 20))
        return band_depth_pipelined(ctx, X, T, n, ld, query_idx, nq, j, count_out);
    double *dX = nullptr;
    if (layout == SD_LAYOUT_TN) {
        SD_TRY(upload_matrix(ctx, BUF_IN, X, T, n, ld, &dX));
    } else {  // [n, T] curve-major: upload, then transpose on the device
        double *dN = nullptr;
        SD_TRY(upload_matrix(ctx, BUF_IN2, X, n, T, ld, &dN));
        SD_TRY(ctx->buf[BUF_IN].reserve((size_t)T * n * sizeof(double)));
        dX = ctx->buf[BUF_IN].as<double>();
    }
    const i64 *d_q = nullptr;
    SD_TRY(upload_queries(ctx, query_idx, nq, n, &d_q, "sd_band_depth_f64"));
    SD_TRY(ctx->buf[BUF_OUT].reserve((size_t)(nq > 0 ? nq : 1) * sizeof(i64)));
    i64 *d_out = ctx->buf[BUF_OUT].as<i64>();
    SD_TRY(mark(ctx, 1));
    if (layout == SD_LAYOUT_NT) SD_TRY(transpose_device(ctx, ctx->buf[BUF_IN2].as<double>(), n, T, T, dX));
    SD_TRY(band_depth_device(ctx, dX, T, n, n, d_q, nq, j, relax, d_out));
    SD_TRY(mark(ctx, 2));
    if (nq > 0)
        SD_CUDA(cudaMemcpyAsync(count_out, d_out, (size_t)nq * sizeof(i64), cudaMemcpyDeviceToHost, ctx->stream));
    return end_call(ctx, true);
}

static int sd_band_depth_f64_dev_impl(sd_ctx *ctx, const double *dX, int64_t T, int64_t n, int64_t ld,
                          const int64_t *d_query_idx, int64_t nq, int j, int relax, int64_t *d_count_out) {
    SD_TRY(check_common(ctx, dX, d_count_out, "sd_band_depth_f64_dev"));
    SD_REQUIRE(T >= 1 && n >= 1 && nq >= 0 && ld >= n, "sd_band_depth_f64_dev: bad sizes");
    SD_REQUIRE(d_query_idx || nq == n, "sd_band_depth_f64_dev: query_idx == NULL requires nq == n");
    SD_TRY(begin_call(ctx));
    SD_TRY(mark(ctx, 1));
    SD_TRY(band_depth_device(ctx, dX, T, n, ld, (const i64 *)d_query_idx, nq, j, relax, (i64 *)d_count_out));
    SD_TRY(mark(ctx, 2));
    if (ctx->async_device && relax) {  // everything is queued on ctx->stream; sd_sync() completes the call
        ctx->pending = 1;
        return SD_OK;
    }
    return end_call(ctx, false);
}

int sd_band_ranks_f64(sd_ctx *ctx, const double *X, int64_t T, int64_t n, int64_t ld, int layout,
                      int32_t *below_out, int32_t *above_out) {
    SD_TRY(check_common(ctx, X, below_out, "sd_band_ranks_f64"));
    SD_REQUIRE(above_out != nullptr, "sd_band_ranks_f64: above_out is NULL");
    SD_REQUIRE(T >= 1 && n >= 1, "sd_band_ranks_f64: bad sizes");
    SD_REQUIRE(layout == SD_LAYOUT_TN || layout == SD_LAYOUT_NT, "sd_band_ranks_f64: bad layout %d", layout);
    SD_REQUIRE(ld >= (layout == SD_LAYOUT_TN ? n : T), "sd_band_ranks_f64: ld too small");
    SD_TRY(begin_call(ctx));
    double *dX = nullptr;
    if (layout == SD_LAYOUT_TN) {
        SD_TRY(upload_matrix(ctx, BUF_IN, X, T, n, ld, &dX));
    } else {
        double *dN = nullptr;
        SD_TRY(upload_matrix(ctx, BUF_IN2, X, n, T, ld, &dN));
        SD_TRY(ctx->buf[BUF_IN].reserve((size_t)T * n * sizeof(double)));
        dX = ctx->buf[BUF_IN].as<double>();
    }
    SD_TRY(ctx->buf[BUF_OUT].reserve((size_t)T * n * 2 * sizeof(int)));
    int *d_b = ctx->buf[BUF_OUT].as<int>();
    int *d_a = d_b + (size_t)T * n;
    SD_TRY(mark(ctx, 1));
    if (layout == SD_LAYOUT_NT) SD_TRY(transpose_device(ctx, ctx->buf[BUF_IN2].as<double>(), n, T, T, dX));
    SD_TRY(mbd_all_device(ctx, dX, T, n, n, false, nullptr, nullptr, d_b, d_a));  // ranks only
    SD_TRY(mark(ctx, 2));
    SD_CUDA(cudaMemcpyAsync(below_out, d_b, (size_t)T * n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(cudaMemcpyAsync(above_out, d_a, (size_t)T * n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return end_call(ctx, true);
}

int sd_simplex_depth_f64(sd_ctx *ctx, const double *F, int64_t N, int64_t T, int d, const int64_t *query_idx,
                         int64_t nq, int relax, double tol, int64_t *count_out) {
    SD_TRY(check_common(ctx, F, count_out, "sd_simplex_depth_f64"));
    SD_REQUIRE(N >= 1 && T >= 1 && nq >= 0, "sd_simplex_depth_f64: bad sizes");
    SD_REQUIRE(d >= 1 && d <= 3, "sd_simplex_depth_f64: d=%d not supported (1..3)", d);
    SD_REQUIRE(tol >= 0.0, "sd_simplex_depth_f64: tol must be >= 0");
    SD_TRY(begin_call(ctx));
    double *dF = nullptr;
    SD_TRY(upload_matrix(ctx, BUF_IN, F, 1, N * T * d, N * T * d, &dF));
    const i64 *d_q = nullptr;
    SD_TRY(upload_queries(ctx, query_idx, nq, N, &d_q, "sd_simplex_depth_f64"));
    SD_TRY(ctx->buf[BUF_OUT].reserve((size_t)(nq > 0 ? nq : 1) * sizeof(i64)));
    i64 *d_out = ctx->buf[BUF_OUT].as<i64>();
    SD_TRY(mark(ctx, 1));
    SD_TRY(simplex_depth_device(ctx, dF, N, T, d, d_q, nq, relax, tol, d_out));
    SD_TRY(mark(ctx, 2));
    if (nq > 0)
        SD_CUDA(cudaMemcpyAsync(count_out, d_out, (size_t)nq * sizeof(i64), cudaMemcpyDeviceToHost, ctx->stream));
    return end_call(ctx, true);
}

int sd_pointcloud_simplicial_f64(sd_ctx *ctx, const double *P, int64_t n, int d, const int64_t *query_idx,
                                 int64_t nq, double tol, int64_t *count_out) {
    SD_TRY(check_common(ctx, P, count_out, "sd_pointcloud_simplicial_f64"));
    SD_REQUIRE(n >= 1 && nq >= 0, "sd_pointcloud_simplicial_f64: bad sizes");
    SD_REQUIRE(d >= 1 && d <= 3, "sd_pointcloud_simplicial_f64: d=%d not supported (1..3)", d);
    SD_REQUIRE(tol >= 0.0, "sd_pointcloud_simplicial_f64: tol must be >= 0");
    SD_TRY(begin_call(ctx));
    double *dP = nullptr;
    SD_TRY(upload_matrix(ctx, BUF_IN, P, 1, n * d, n * d, &dP));
    const i64 *d_q = nullptr;
    SD_TRY(upload_queries(ctx, query_idx, nq, n, &d_q, "sd_pointcloud_simplicial_f64"));
    SD_TRY(ctx->buf[BUF_OUT].reserve((size_t)(nq > 0 ? nq : 1) * sizeof(i64)));
    i64 *d_out = ctx->buf[BUF_OUT].as<i64>();
    SD_TRY(mark(ctx, 1));
    SD_TRY(simplicial_device(ctx, dP, n, d, d_q, nq, tol, d_out));
    SD_TRY(mark(ctx, 2));
    if (nq > 0)
        SD_CUDA(cudaMemcpyAsync(count_out, d_out, (size_t)nq * sizeof(i64), cudaMemcpyDeviceToHost, ctx->stream));
    return end_call(ctx, true);
}

int sd_pointcloud_l1_f64(sd_ctx *ctx, const double *P, int64_t n, int d, const int64_t *query_idx, int64_t nq,
                         double *depth_out) {
    SD_TRY(check_common(ctx, P, depth_out, "sd_pointcloud_l1_f64"));
    SD_REQUIRE(n >= 1 && nq >= 0, "sd_pointcloud_l1_f64: bad sizes");
    SD_REQUIRE(d >= 1 && d <= 16, "sd_pointcloud_l1_f64: d=%d not supported (1..16)", d);
    SD_TRY(begin_call(ctx));
    double *dP = nullptr;
    SD_TRY(upload_matrix(ctx, BUF_IN, P, 1, n * d, n * d, &dP));
    const i64 *d_q = nullptr;
    SD_TRY(upload_queries(ctx, query_idx, nq, n, &d_q, "sd_pointcloud_l1_f64"));
    SD_TRY(ctx->buf[BUF_OUT].reserve((size_t)(nq > 0 ? nq : 1) * sizeof(double)));
    double *d_out = ctx->buf[BUF_OUT].as<double>();
    SD_TRY(mark(ctx, 1));
    SD_TRY(l1_device(ctx, dP, n, d, d_q, nq, d_out));
    SD_TRY(mark(ctx, 2));
    if (nq > 0)
        SD_CUDA(cudaMemcpyAsync(depth_out, d_out, (size_t)nq * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return end_call(ctx, true);
}

int sd_pointcloud_oja_f64(sd_ctx *ctx, const double *P, int64_t n, int d, const int64_t *query_idx, int64_t nq,
                          const int64_t *pool, int64_t npool, double hull_volume, double *out) {
    SD_TRY(check_common(ctx, P, out, "sd_pointcloud_oja_f64"));
    SD_REQUIRE(n >= 1 && nq >= 0 && npool >= 0, "sd_pointcloud_oja_f64: bad sizes");
    SD_REQUIRE(d == 2 || d == 3, "sd_pointcloud_oja_f64: d=%d not supported (2 or 3)", d);
    SD_REQUIRE(pool || npool == n, "sd_pointcloud_oja_f64: pool == NULL requires npool == n");
    SD_TRY(begin_call(ctx));
    double *dP = nullptr;
    SD_TRY(upload_matrix(ctx, BUF_IN, P, 1, n * d, n * d, &dP));
    const i64 *d_q = nullptr;
    SD_TRY(upload_queries(ctx, query_idx, nq, n, &d_q, "sd_pointcloud_oja_f64"));
    const i64 *d_pool = nullptr;
    if (pool) {
        for (i64 i = 0; i < npool; ++i)
            SD_REQUIRE(pool[i] >= 0 && pool[i] < n, "sd_pointcloud_oja_f64: pool[%lld] out of range", (long long)i);
        SD_TRY(ctx->buf[BUF_AUX].reserve((size_t)(npool > 0 ? npool : 1) * sizeof(i64)));
        if (npool > 0)
            SD_CUDA(cudaMemcpyAsync(ctx->buf[BUF_AUX].p, pool, (size_t)npool * sizeof(i64), cudaMemcpyHostToDevice,
                                    ctx->stream));
        d_pool = ctx->buf[BUF_AUX].as<i64>();
    }
    SD_TRY(ctx->buf[BUF_OUT].reserve((size_t)(nq > 0 ? nq : 1) * sizeof(double)));
    double *d_out = ctx->buf[BUF_OUT].as<double>();
    SD_TRY(mark(ctx, 1));
    SD_TRY(oja_device(ctx, dP, n, d, d_q, nq, d_pool, npool, hull_volume, d_out));
    SD_TRY(mark(ctx, 2));
    if (nq > 0)
        SD_CUDA(cudaMemcpyAsync(out, d_out, (size_t)nq * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return end_call(ctx, true);
}


static int sd_band_depth_batched_f64_impl(sd_ctx *ctx, const double *X, int64_t T, int64_t n, int64_t ld,
                              const uint8_t *membership, int64_t B, const int64_t *queries, int64_t nqb, int j,
                              int relax, int64_t *count_out) {
    SD_TRY(check_common(ctx, X, count_out, "sd_band_depth_batched_f64"));
    SD_REQUIRE(membership && queries, "sd_band_depth_batched_f64: NULL membership/queries");
    SD_REQUIRE(T >= 1 && n >= 1 && B >= 0 && nqb >= 1 && ld >= n, "sd_band_depth_batched_f64: bad sizes");
    // host side: member lists and query positions inside each batch
    std::vector<i64> cols, offs((size_t)B + 1, 0), qloc((size_t)(B * nqb));
    std::vector<i64> pos((size_t)n);
    for (i64 b = 0; b < B; ++b) {
        const uint8_t *mb = membership + b * n;
        i64 m = 0;
        for (i64 c = 0; c < n; ++c) {
            pos[(size_t)c] = -1;
            if (mb[c]) {
                pos[(size_t)c] = m++;
                cols.push_back(c);
            }
        }
        offs[(size_t)b + 1] = offs[(size_t)b] + m;
        for (i64 i = 0; i < nqb; ++i) {
            const i64 g = queries[b * nqb + i];
            SD_REQUIRE(g >= 0 && g < n && pos[(size_t)g] >= 0,
                       "sd_band_depth_batched_f64: query %lld of batch %lld is not a member", (long long)g,
                       (long long)b);
            qloc[(size_t)(b * nqb + i)] = pos[(size_t)g];
        }
    }
    // relaxed depth with equally sized batches (permutation tests): ALL batches go through the rank pipeline as
    // one matrix of B*T rows whose row groups accumulate separately -- a handful of launches instead of ~10 per batch
    const bool strict_match = !relax && j == 2 && (ctx->bd_impl == SD_BD_AUTO || ctx->bd_impl == SD_BD_MATCH);
    bool uniform = B > 1 && ((relax != 0 && (j == 2 || j == 3)) || strict_match);
    const i64 m0 = B > 0 ? offs[1] - offs[0] : 0;
    for (i64 b = 0; b < B && uniform; ++b) uniform = offs[(size_t)b + 1] - offs[(size_t)b] == m0;
    if (uniform && !relax && !bd_match_supported(T, m0)) uniform = false;
    if (uniform && m0 >= 1 && relax) {
        const long double full = (j == 2) ? (long double)(m0 - 1) * (m0 - 2) / 2.0L
                                          : (long double)(m0 - 1) * (m0 - 2) * (m0 - 3) / 6.0L;
        if (full * (long double)T >= 9.0e18L) uniform = false;  // let the per-batch path report the overflow
    }
    if (uniform && m0 >= 1 && relax)
        for (i64 b = 0; b < B; ++b)  // flat index of every query in the [B][m0] accumulator
            for (i64 i = 0; i < nqb; ++i) qloc[(size_t)(b * nqb + i)] += b * m0;
    SD_TRY(begin_call(ctx));
    double *dX = nullptr;
    SD_TRY(upload_matrix(ctx, BUF_IN, X, T, n, ld, &dX));
    SD_TRY(ctx->buf[BUF_AUX].reserve((cols.size() + 1) * sizeof(i64)));
    SD_TRY(ctx->buf[BUF_MISC].reserve((qloc.size() + 1) * sizeof(i64)));
    SD_TRY(ctx->buf[BUF_OUT].reserve((qloc.size() + 1) * sizeof(i64)));
    i64 *d_cols = ctx->buf[BUF_AUX].as<i64>();
    i64 *d_ql = ctx->buf[BUF_MISC].as<i64>();
    i64 *d_out = ctx->buf[BUF_OUT].as<i64>();
    if (!cols.empty())
        SD_CUDA(cudaMemcpyAsync(d_cols, cols.data(), cols.size() * sizeof(i64), cudaMemcpyHostToDevice, ctx->stream));
    if (!qloc.empty())
        SD_CUDA(cudaMemcpyAsync(d_ql, qloc.data(), qloc.size() * sizeof(i64), cudaMemcpyHostToDevice, ctx->stream));
    SD_TRY(mark(ctx, 1));
    if (uniform && m0 >= 1) {
        // batches per pass: the gathered matrix stays below ~1 GB
        i64 BB = (i64)((1ull << 30) / ((size_t)T * m0 * sizeof(double)));
        if (BB < 1) BB = 1;
        if (BB > B) BB = B;
        SD_TRY(ctx->buf[BUF_IN2].reserve((size_t)BB * T * m0 * sizeof(double)));
        SD_TRY(ctx->buf[BUF_ACC].reserve((size_t)B * m0 * 2 * sizeof(i64)));
        double *dXg = ctx->buf[BUF_IN2].as<double>();
        i64 *acc2 = ctx->buf[BUF_ACC].as<i64>(), *acc3 = acc2 + B * m0;
        for (i64 b0 = 0; b0 < B; b0 += BB) {
            const i64 nb = B - b0 < BB ? B - b0 : BB;
            SD_TRY(compact_batches_device(ctx, dX, T, n, d_cols + b0 * m0, m0, nb, dXg));
            if (relax) {
                SD_TRY(mbd_all_device(ctx, dXg, nb * T, m0, m0, j == 3, acc2 + b0 * m0, acc3 + b0 * m0, nullptr,
                                      nullptr, false, T));
            } else {  // strict: sign-vector matcher over all sub-populations of the pass
                ctx->last.bd_impl_used = SD_BD_MATCH;
                SD_TRY(bd_strict_match_batched_device(ctx, dXg, nb, T, m0, d_ql + b0 * nqb, nqb, d_out + b0 * nqb,
                                                      strict_enumerate));
            }
        }
        if (relax) SD_TRY(gather_i64_device(ctx, j == 2 ? acc2 : acc3, d_ql, (i64)qloc.size(), d_out));
    } else {
        SD_TRY(ctx->buf[BUF_IN2].reserve((size_t)T * n * sizeof(double)));
        double *dXb = ctx->buf[BUF_IN2].as<double>();
        for (i64 b = 0; b < B; ++b) {
            const i64 m = offs[(size_t)b + 1] - offs[(size_t)b];
            SD_TRY(compact_columns_device(ctx, dX, T, n, d_cols + offs[(size_t)b], m, dXb));
            SD_TRY(band_depth_device(ctx, dXb, T, m, m, d_ql + b * nqb, nqb, j, relax, d_out + b * nqb));
        }
    }
    SD_TRY(mark(ctx, 2));
    if (!qloc.empty())
        SD_CUDA(cudaMemcpyAsync(count_out, d_out, qloc.size() * sizeof(i64), cudaMemcpyDeviceToHost, ctx->stream));
    // cols / qloc are pageable host vectors: the async copies above were staged synchronously
    return end_call(ctx, true);
}

int sd_band_depth_f64(sd_ctx *ctx, const double *X, int64_t T, int64_t n, int64_t ld, int layout,
                      const int64_t *query_idx, int64_t nq, int j, int relax, int64_t *count_out) {
    try {
        return sd_band_depth_f64_impl(ctx, X, T, n, ld, layout, query_idx, nq, j, relax, count_out);
    } catch (...) {  // std::bad_alloc from host-side staging vectors: the ABI never throws
        sd::set_error("sd_band_depth_f64: out of host memory");
        return SD_ERR_INVALID;
    }
}

int sd_band_depth_f64_dev(sd_ctx *ctx, const double *dX, int64_t T, int64_t n, int64_t ld,
                          const int64_t *d_query_idx, int64_t nq, int j, int relax, int64_t *d_count_out) {
    try {
        return sd_band_depth_f64_dev_impl(ctx, dX, T, n, ld, d_query_idx, nq, j, relax, d_count_out);
    } catch (...) {  // std::bad_alloc from host-side staging vectors: the ABI never throws
        sd::set_error("sd_band_depth_f64_dev: out of host memory");
        return SD_ERR_INVALID;
    }
}

int sd_pointcloud_blocks_f64(sd_ctx *ctx, const double *P, int64_t n, int d, const int64_t *member,
                             const int64_t *block_off, const int64_t *query_pos, int64_t B, int kind, double tol,
                             const double *hull_volume, double *out) {
    SD_TRY(check_common(ctx, P, out, "sd_pointcloud_blocks_f64"));
    SD_REQUIRE(n >= 1 && B >= 0 && member && block_off && query_pos, "sd_pointcloud_blocks_f64: bad arguments");
    SD_REQUIRE(d >= 1 && d <= 3, "sd_pointcloud_blocks_f64: d=%d not supported (1..3)", d);
    SD_REQUIRE(kind >= 0 && kind <= 2, "sd_pointcloud_blocks_f64: kind=%d (0 simplicial, 1 l1, 2 oja)", kind);
    SD_REQUIRE(kind != 2 || (hull_volume && d >= 2), "sd_pointcloud_blocks_f64: oja needs d in (2, 3) and hull volumes");
    SD_REQUIRE(tol >= 0.0, "sd_pointcloud_blocks_f64: tol must be >= 0");
    SD_REQUIRE(block_off[0] == 0, "sd_pointcloud_blocks_f64: block_off[0] must be 0");
    for (i64 b = 0; b < B; ++b) {
        const i64 m = block_off[b + 1] - block_off[b];
        SD_REQUIRE(m >= 1 && m * d <= 4096, "sd_pointcloud_blocks_f64: block %lld has %lld members (1 .. %d supported)",
                   (long long)b, (long long)m, 4096 / d);
        SD_REQUIRE(query_pos[b] >= 0 && query_pos[b] < m, "sd_pointcloud_blocks_f64: query_pos[%lld] out of range",
                   (long long)b);
    }
    const i64 total = B > 0 ? block_off[B] : 0;
    for (i64 i = 0; i < total; ++i)
        SD_REQUIRE(member[i] >= 0 && member[i] < n, "sd_pointcloud_blocks_f64: member[%lld] out of range", (long long)i);
    SD_TRY(begin_call(ctx));
    double *dP = nullptr;
    SD_TRY(upload_matrix(ctx, BUF_IN, P, 1, n * d, n * d, &dP));
    // member | block_off | query_pos | hull volumes | out, in one workspace buffer
    const size_t words = (size_t)total + (size_t)(B + 1) + (size_t)B + (size_t)B + (size_t)B + 8;
    SD_TRY(ctx->buf[BUF_AUX].reserve(words * 8));
    i64 *d_member = ctx->buf[BUF_AUX].as<i64>();
    i64 *d_off = d_member + total;
    i64 *d_qpos = d_off + (B + 1);
    double *d_vol = reinterpret_cast<double *>(d_qpos + B);
    double *d_out = d_vol + B;
    if (total > 0) SD_CUDA(cudaMemcpyAsync(d_member, member, (size_t)total * 8, cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(cudaMemcpyAsync(d_off, block_off, (size_t)(B + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (B > 0) {
        SD_CUDA(cudaMemcpyAsync(d_qpos, query_pos, (size_t)B * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (hull_volume) SD_CUDA(cudaMemcpyAsync(d_vol, hull_volume, (size_t)B * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    SD_TRY(mark(ctx, 1));
    SD_TRY(cloud_blocks_device(ctx, dP, d, d_member, d_off, d_qpos, B, kind, tol, hull_volume ? d_vol : nullptr, d_out));
    SD_TRY(mark(ctx, 2));
    if (B > 0) SD_CUDA(cudaMemcpyAsync(out, d_out, (size_t)B * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return end_call(ctx, true);
}

int sd_band_depth_batched_f64(sd_ctx *ctx, const double *X, int64_t T, int64_t n, int64_t ld,
                              const uint8_t *membership, int64_t B, const int64_t *queries, int64_t nqb, int j,
                              int relax, int64_t *count_out) {
    try {
        return sd_band_depth_batched_f64_impl(ctx, X, T, n, ld, membership, B, queries, nqb, j, relax, count_out);
    } catch (...) {  // std::bad_alloc from host-side staging vectors: the ABI never throws
        sd::set_error("sd_band_depth_batched_f64: out of host memory");
        return SD_ERR_INVALID;
    }
}

}  // extern "C"
