// common.cuh -- shared host/device plumbing of libsdepth.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/statdepth_b200.h"

namespace sd {

typedef long long i64;
typedef unsigned long long u64;
typedef unsigned int u32;

void set_error(const char *fmt, ...);

#define SD_CUDA(call)                                                                            \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            sd::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return SD_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

#define SD_TRY(call)                \
    do {                            \
        int _s = (call);            \
        if (_s != SD_OK) return _s; \
    } while (0)

#define SD_REQUIRE(cond, ...)         \
    do {                              \
        if (!(cond)) {                \
            sd::set_error(__VA_ARGS__); \
            return SD_ERR_INVALID;    \
        }                             \
    } while (0)

// Grow-only device buffer owned by the context.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

// workspace slots of a context
enum BufId {
    BUF_IN = 0,     // device copy of the caller's host input
    BUF_IN2,        // transposed / secondary input
    BUF_QIDX,       // query indices
    BUF_OUT,        // result staging
    BUF_ACC,        // int64 accumulators (MBD: acc_j2[n], acc_j3[n])
    BUF_SPLIT,      // MBD: per-row splitters
    BUF_CURSOR,     // MBD: per-(row, part) fill counts + per-row overflow flags
    BUF_PART_X,     // MBD: partitioned values  / fallback sort scratch
    BUF_PART_J,     // MBD: partitioned curve ids
    BUF_MASK,       // strict BD: per-query sign masks
    BUF_MISC,       // small odds and ends
    BUF_AUX,        // batched / extra
    BUF_WORK,       // MBD: work list of big parts
    BUF_RANKS,      // strict BD matcher: per-time-point ranks of all curves (packed pairs)
    BUF_GEOM,       // Oja counting: the pool's points and the queries' positions inside the pool
    BUF_SLAB,       // MBD slab path: per-row maps, bucket tables, bin starts
    NUM_BUFS
};

// status bits written by kernels into ctx->d_status
enum { ST_NONFINITE = 1, ST_INTERNAL = 2 };

__host__ __device__ static inline i64 ceil_div(i64 a, i64 b) { return (a + b - 1) / b; }

// order-preserving map double -> u64 (with -0.0 folded onto +0.0 so that == ties stay ties)
__host__ __device__ static inline u64 sortable_key(double x) {
    if (x == 0.0) x = 0.0;
#ifdef __CUDA_ARCH__
    u64 b = (u64)__double_as_longlong(x);
#else
    u64 b;
    memcpy(&b, &x, 8);
#endif
    const u64 mask = (u64)(-(i64)(b >> 63)) | 0x8000000000000000ull;
    return b ^ mask;
}

}  // namespace sd

struct sd_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;                         // H2D of row blocks overlapped with ranking
    cudaEvent_t ev_pipe[4] = {nullptr, nullptr, nullptr, nullptr};  // copied[2], consumed[2]
    void *stage[2] = {nullptr, nullptr};                        // pinned staging of pageable host input (api.cu)
    size_t stage_cap[2] = {0, 0};
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};
    cudaEvent_t ev_slab[2] = {nullptr, nullptr};                // MBD slab path: unfit-row counts copied (mbd.cu)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // h2d start | kernels start | kernels end | d2h end
    sd_timings last = {0, 0, 0, 0, 0, 0};
    int bd_impl = SD_BD_AUTO;
    int mbd_force_fallback = 0;
    int mbd_no_wait = 0;    // set around pipelined host calls: mbd_all_device must not wait for device results
    int profile = 0;
    int simplicial_impl = SD_SIMPLICIAL_AUTO;
    int async_device = 0;   // SD_OPT_ASYNC_DEVICE
    int pending = 0;        // an asynchronous device call has been queued and not yet completed by sd_sync()
    static const int MAX_PROF = 256;                 // event pairs per call when profiling
    cudaEvent_t prof_ev[2 * MAX_PROF] = {};          // created lazily
    int prof_phase[MAX_PROF] = {};
    int prof_n = 0;
    int64_t phase_ns[SD_PHASE_COUNT] = {};
    sd::DevBuf buf[sd::NUM_BUFS];
    int *d_status = nullptr;  // device int[4]: [0] status bits, [1] MBD rows ranked by the generic path
    int *h_status = nullptr;  // pinned mirror; [2] and [3] also receive the slab path's unfit-row counts mid-call (mbd.cu)
};

namespace sd {

// timing helpers: record the four phase boundaries of a host-buffer call
int begin_call(sd_ctx *ctx);                 // resets status, records ev[0]
int mark(sd_ctx *ctx, int which);            // records ev[which]
int end_call(sd_ctx *ctx, bool had_copies);  // syncs, fills ctx->last, maps status bits to errors
int check_status(sd_ctx *ctx);               // (after sync) translate *h_status
// profiling brackets (no-ops unless SD_OPT_PROFILE): prof_begin before a phase's launches, prof_end after
int prof_begin(sd_ctx *ctx, int phase);
int prof_end(sd_ctx *ctx);

// kernels / drivers implemented in the other translation units (all stream-ordered on ctx->stream)
int band_depth_device(sd_ctx *ctx, const double *dX, i64 T, i64 n, i64 ld, const i64 *d_q, i64 nq, int j,
                      int relax, i64 *d_out);
int mbd_all_device(sd_ctx *ctx, const double *dX, i64 T, i64 n, i64 ld, bool want_j3, i64 *d_acc2,
                   i64 *d_acc3, int *d_rank_b, int *d_rank_a, bool accumulate = false, i64 group_rows = 0);
int bd_strict_device(sd_ctx *ctx, const double *dX, i64 T, i64 n, i64 ld, const i64 *d_q, i64 nq, int j,
                     i64 *d_out, u64 *d_hits = nullptr);
int bd_strict_gemm_device(sd_ctx *ctx, const double *dX, i64 T, i64 n, i64 ld, const i64 *d_q, i64 nq,
                          i64 *d_out);
int probe_int8_peak(sd_ctx *ctx, double *ops_per_s);
bool bd_match_supported(i64 T, i64 n);
int bd_strict_match_device(sd_ctx *ctx, const double *dX, i64 T, i64 n, i64 ld, const i64 *d_q, i64 nq, i64 *d_out,
                           int (*fallback)(sd_ctx *, const double *, i64, i64, i64, const i64 *, i64, i64 *),
                           i64 *n_fallback);
int bd_strict_match_batched_device(sd_ctx *ctx, const double *dXg, i64 nb, i64 T, i64 n, const i64 *d_ql, i64 nqb,
                                   i64 *d_out,
                                   int (*fallback)(sd_ctx *, const double *, i64, i64, i64, const i64 *, i64, i64 *));
int transpose_device(sd_ctx *ctx, const double *d_in, i64 rows, i64 cols, i64 ld_in, double *d_out);
int gather_i64_device(sd_ctx *ctx, const i64 *d_src, const i64 *d_idx, i64 nq, i64 *d_out);
int compact_columns_device(sd_ctx *ctx, const double *dX, i64 T, i64 n, const i64 *d_cols, i64 m, double *d_out);
int compact_batches_device(sd_ctx *ctx, const double *dX, i64 T, i64 n, const i64 *d_cols, i64 m, i64 nb,
                           double *d_out);
int l1_device(sd_ctx *ctx, const double *dP, i64 n, int d, const i64 *d_q, i64 nq, double *d_out);
int oja_device(sd_ctx *ctx, const double *dP, i64 n, int d, const i64 *d_q, i64 nq, const i64 *d_pool,
               i64 npool, double hull_volume, double *d_out);
int cloud_blocks_device(sd_ctx *ctx, const double *dP, int d, const i64 *d_member, const i64 *d_off,
                        const i64 *d_qpos, i64 B, int kind, double tol, const double *d_volume, double *d_out);
int simplicial_device(sd_ctx *ctx, const double *dP, i64 n, int d, const i64 *d_q, i64 nq, double tol,
                      i64 *d_out);
int simplicial2_count_device(sd_ctx *ctx, const double *d_pts, i64 n, i64 stride_j, i64 T, const i64 *d_q, i64 nq,
                             double tol, i64 *d_out);
int oja2_count_device(sd_ctx *ctx, const double *d_pts, i64 n, const i64 *d_q, i64 nq, double hull_volume,
                      double *d_out);
int simplex_depth_device(sd_ctx *ctx, const double *dF, i64 N, i64 T, int d, const i64 *d_q, i64 nq,
                         int relax, double tol, i64 *d_out);

}  // namespace sd
