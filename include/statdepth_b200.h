/*
 * statdepth_b200.h -- C ABI of the B200-native depth engine (libsdepth.so).
 *
 * Drop-in boundary for statdepth's data-parallel hot path.  The reference has no FFI: its seam is
 * two private Python drivers imported by name in statdepth/depth/depth.py:7-8,
 *     _functionaldepth(data, to_compute, J, containment, relax, deep_check, quiet) -> pd.Series
 *         (statdepth/depth/calculations/_functional.py:17-97)
 *     _pointwisedepth(data, to_compute, containment, quiet) -> pd.Series
 *         (statdepth/depth/calculations/_pointcloud.py:14-66)
 * Each entry point below replaces the inner loops of one of those drivers and returns the
 * INTEGER numerators (or float64 depths where the reference's result is a float sum); the Python
 * host (statdepth_b200/) turns them into the same pd.Series the reference returns.  The ctypes
 * binding a maintainer of the reference would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host buffer, the library owns all
 *     device memory (per-context workspace, grown on demand, freed by sd_destroy);
 *   - every function returns an int status (SD_OK == 0), never throws or aborts;
 *     sd_last_error() returns a thread-local, NUL-terminated description of the last failure;
 *   - every call is synchronous on return; internally stream-ordered on the context's stream;
 *   - one context per GPU per process (one process per GPU; multi-GPU sharding is done by the
 *     host with torch.distributed, see DESIGN.md "Multi-GPU");
 *   - `*_dev` variants take DEVICE pointers (e.g. torch tensors' data_ptr()) and do no copies:
 *     they are what bench.py times as the HBM-resident `value`.
 *   - non-finite inputs are rejected with SD_ERR_NONFINITE (documented divergence: the reference
 *     silently skips NaNs through pandas min/max, _containment.py:68-69).
 */
#ifndef STATDEPTH_B200_H
#define STATDEPTH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_ABI_VERSION 1

typedef struct sd_ctx sd_ctx;

enum sd_status {
    SD_OK = 0,
    SD_ERR_INVALID = 1,     /* bad argument (null pointer, negative size, unsupported J/d/layout) */
    SD_ERR_CUDA = 2,        /* a CUDA runtime call failed; text in sd_last_error() */
    SD_ERR_NO_DEVICE = 3,   /* no usable sm_100 device */
    SD_ERR_NONFINITE = 4,   /* NaN or +-inf in the input */
    SD_ERR_OVERFLOW = 5,    /* integer numerator would exceed int64 for this (n, T, J) */
    SD_ERR_UNSUPPORTED = 6  /* configuration the engine does not implement (no CPU fallback) */
};

/* layout of the univariate matrix handed to sd_band_depth_f64 */
enum sd_layout {
    SD_LAYOUT_TN = 0, /* X[t*ld + c]: T time rows, n curves, ld >= n  (== DataFrame.to_numpy(), C order) */
    SD_LAYOUT_NT = 1  /* X[c*ld + t]: n curve rows, T points, ld >= T (== F-ordered DataFrame values)   */
};

/* strict band-depth kernel selection (sd_set_option(ctx, SD_OPT_BD_IMPL, ...)) */
enum sd_bd_impl {
    SD_BD_AUTO = 0,  /* sign-vector matching when it applies; queries it cannot certify (ties with the query,
                        n > 8193) go to the bit kernel, or to the dense Gram if a probe of 8 queries shows that
                        more than 2 % of all pairs survive the first mask word */
    SD_BD_BITS = 1,  /* bit-packed AND + early exit on CUDA cores                      */
    SD_BD_GEMM = 2,  /* int8 violation Gram V = Sb Sb^T + Sa Sa^T on tcgen05 / TMEM    */
    SD_BD_MATCH = 3  /* complementary sign-vector matching (O(nT) per query) + enumeration of uncertified queries */
};

enum sd_option {
    SD_OPT_BD_IMPL = 1,       /* enum sd_bd_impl */
    SD_OPT_MBD_FORCE_FALLBACK = 2, /* 1: rank every row with the generic (slow) path; testing aid */
    SD_OPT_PROFILE = 3,           /* 1: bracket every kernel phase with CUDA events (sd_get_phase_ns) */
    SD_OPT_SIMPLICIAL_IMPL = 4,   /* enum sd_simplicial_impl: 2-D triangle counting algorithm */
    SD_OPT_ASYNC_DEVICE = 5       /* 1: sd_band_depth_f64_dev (relax = 1) only ENQUEUES its work on sd_stream() and
                                     returns; status, timings and the result are valid after sd_sync().  Lets a caller
                                     queue a collective behind the kernels without a host round trip. */
};

/* how 2-D simplicial counts (point clouds, relaxed multivariate simplex depth) are obtained */
enum sd_simplicial_impl {
    SD_SIMPLICIAL_AUTO = 0,      /* enumerate samples of up to 64 points / curves, count larger ones (d = 2); same semantics */
    SD_SIMPLICIAL_ENUMERATE = 1, /* all (d+1)-subsets, closed simplex test with tolerance `tol` */
    SD_SIMPLICIAL_COUNT = 2      /* O(n log n) angular counting, d = 2: dist(p, triangle) <= tol; tol = 0: exact closed triangles */
};

/* phases of the modified-band-depth pipeline reported by sd_get_phase_ns */
enum sd_phase {
    SD_PHASE_MBD_SPLITTERS = 0,
    SD_PHASE_MBD_PARTITION = 1,
    SD_PHASE_MBD_RANK = 2,
    SD_PHASE_MBD_GENERIC = 3,
    SD_PHASE_BD_MASKS = 4,
    SD_PHASE_BD_PAIRS = 5,
    SD_PHASE_MBD_SLAB_HIST = 6, /* slab path: table + hist kernels (range, code table, per-value codes, bin starts) */
    SD_PHASE_MBD_SLAB_RANK = 7, /* slab path: rank kernel */
    SD_PHASE_COUNT = 8
};

/* nanoseconds of the LAST call on this context, from CUDA events on the context's stream */
typedef struct sd_timings {
    int64_t h2d_ns;      /* host->device copies                */
    int64_t kernel_ns;   /* all kernels of the call            */
    int64_t d2h_ns;      /* device->host copy of the result    */
    int64_t launches;    /* number of kernel launches          */
    int64_t fallback_rows; /* MBD: time rows ranked by the generic path (0 on well-spread data) */
    int64_t bd_impl_used;  /* strict band depth: enum sd_bd_impl that did (most of) the work, 0 if n/a */
} sd_timings;

typedef struct sd_devinfo {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    int64_t total_mem;
    char name[128];
} sd_devinfo;

int sd_abi_version(void);
const char *sd_last_error(void);

/* create / destroy a context on CUDA device `device` (creates a stream and timing events). */
int sd_init(int device, sd_ctx **out);
int sd_destroy(sd_ctx *ctx);
int sd_device_info(sd_ctx *ctx, sd_devinfo *out);
int sd_set_option(sd_ctx *ctx, int option, int64_t value);
int sd_get_timings(sd_ctx *ctx, sd_timings *out);
/* with SD_OPT_PROFILE on: device nanoseconds per phase (enum sd_phase) of the LAST call, summed over
 * its launches of that phase; out must hold SD_PHASE_COUNT entries. */
int sd_get_phase_ns(sd_ctx *ctx, int64_t *out);
/* micro-benchmark: sustained tcgen05.mma kind::i8 rate of this GPU in int8 ops/s (2 per MAC), the measured
 * denominator for the Gram kernel's roofline fraction (takes ~10 ms) */
int sd_probe_int8_peak(sd_ctx *ctx, double *ops_per_s);
/* the context's cudaStream_t (as void*), so a caller can order its own work / events on it */
void *sd_stream(sd_ctx *ctx);
/* waits for everything queued on sd_stream(); with SD_OPT_ASYNC_DEVICE it completes the pending
 * sd_band_depth_f64_dev call (returns its status, makes sd_get_timings valid) */
int sd_sync(sd_ctx *ctx);
/* Which rank pipeline relaxed depth takes for rows of n curves with leading dimension ld, and its geometry (host
 * arithmetic only: no context, no GPU).  out6[0] = 1: slab path (csrc/mbd_slab.cuh), 0: part pipeline; for the slab
 * path out6[1] = rank CTAs per row, [2] = bins per CTA, [3] = entries one CTA may hold, [4] / [5] = dynamic shared
 * memory of the rank / hist kernel in bytes.  The device pointer is assumed 16-byte aligned; the SD_MBD_* environment
 * variables apply as they do to the depth calls. */
int sd_mbd_plan(int64_t n, int64_t ld, int64_t *out6);

/*
 * Univariate band depth numerators.  Replaces _univariate_band_depth (_functional.py:198-255)
 * + _r2_containment (_containment.py:45-80) + _subsequences (_helper.py:16-32) for one j.
 *
 *   relax == 0 : count_out[q] = #{j-subsets S of the other n-1 curves : the band of S contains
 *                curve query_idx[q] at ALL T points}              (closed interval, J = 2 or 3)
 *   relax != 0 : count_out[q] = sum_t #{j-subsets whose band contains the curve at time t}
 *                             = sum_t [C(n-1,j) - C(b_t,j) - C(a_t,j)]
 * The host forms  depth = sum_{j=2..J} count_j [/ T] / C(n, j)   (_functional.py:229,253).
 *
 * query_idx == NULL means all n curves (nq must then equal n).  T-sharding for multi-GPU MBD:
 * pass a block of time rows; relaxed counts are additive over disjoint row blocks.
 */
int sd_band_depth_f64(sd_ctx *ctx, const double *X, int64_t T, int64_t n, int64_t ld, int layout,
                      const int64_t *query_idx, int64_t nq, int j, int relax, int64_t *count_out);

/* same, X / query_idx / count_out are DEVICE pointers, layout SD_LAYOUT_TN only */
int sd_band_depth_f64_dev(sd_ctx *ctx, const double *dX, int64_t T, int64_t n, int64_t ld,
                          const int64_t *d_query_idx, int64_t nq, int j, int relax,
                          int64_t *d_count_out);

/*
 * Per-(time, curve) strict ranks: below[t*n + c] = #curves strictly below curve c at time t,
 * above[...] likewise (int32).  The integer intermediates SURVEY 8a calls "rank/count
 * intermediates"; also what relaxed depth for J >= 4 is assembled from on the host.
 */
int sd_band_ranks_f64(sd_ctx *ctx, const double *X, int64_t T, int64_t n, int64_t ld, int layout,
                      int32_t *below_out, int32_t *above_out);

/*
 * Multivariate functional simplex depth numerators.  Replaces _simplex_depth
 * (_functional.py:257-286) + _simplex_containment / _is_in_simplex (_containment.py:105-176).
 * F[(i*T + t)*d + c]: N curves, T rows, d in {1,2,3} channels.  tol = absolute tolerance band
 * of the closed simplex test (the reference's LP accepts ~1e-7; pass 0 for exact sign tests).
 *   relax == 0 : #{(d+1)-subsets of the other N-1 curves containing the query at all T rows}
 *   relax != 0 : sum over subsets of #rows contained
 */
int sd_simplex_depth_f64(sd_ctx *ctx, const double *F, int64_t N, int64_t T, int d,
                         const int64_t *query_idx, int64_t nq, int relax, double tol,
                         int64_t *count_out);

/* Point cloud P[i*d + c], n points.  Replaces the 'simplex' branch of _pointwisedepth
 * (_pointcloud.py:44-56): #{(d+1)-subsets of the other points whose closed simplex contains p}. */
int sd_pointcloud_simplicial_f64(sd_ctx *ctx, const double *P, int64_t n, int d,
                                 const int64_t *query_idx, int64_t nq, double tol,
                                 int64_t *count_out);

/* Replaces _L1_depth (_pointcloud.py:125-150): 1 - ||sum_{o!=p} (x_o-x_p)/||x_p-x_o|| || / n. */
int sd_pointcloud_l1_f64(sd_ctx *ctx, const double *P, int64_t n, int d, const int64_t *query_idx,
                         int64_t nq, double *depth_out);

/* Replaces _oja_depth (_pointcloud.py:176-204): sum over d-subsets S of pool\{p} of
 * vol(conv(S u {p})) / hull_volume.  pool == NULL means all n points. */
int sd_pointcloud_oja_f64(sd_ctx *ctx, const double *P, int64_t n, int d, const int64_t *query_idx,
                          int64_t nq, const int64_t *pool, int64_t npool, double hull_volume,
                          double *out);

/*
 * B small sub-clouds of one cloud with ONE query each, in one launch: what K-sampled point-cloud depth consists
 * of (_samplepointwisedepth, _pointcloud.py:97-123).  Block b holds the points member[block_off[b] .. block_off[b+1])
 * (ids into P, in the reference's order: the L1 sum runs in that order, bit for bit), its query is the member at
 * position query_pos[b].  kind 0: simplicial count (returned as a double, exact below 2^53), 1: L1 depth,
 * 2: Oja sum / hull_volume[b] (d in {2,3}).  At most 4096 / d members per block (they are enumerated).
 */
int sd_pointcloud_blocks_f64(sd_ctx *ctx, const double *P, int64_t n, int d, const int64_t *member,
                             const int64_t *block_off, const int64_t *query_pos, int64_t B, int kind, double tol,
                             const double *hull_volume, double *out);

/*
 * Batched band depth over sub-populations of one matrix (K-sampled blocks, homogeneity
 * permutations): membership[b*n + c] != 0 selects the curves of batch b; queries[b*nqb + i] is
 * the i-th query curve of batch b (must be a member).  count_out[b*nqb + i] as in
 * sd_band_depth_f64 with n replaced by the batch's member count.  One launch sequence for all
 * batches.  X host, layout SD_LAYOUT_TN.
 */
int sd_band_depth_batched_f64(sd_ctx *ctx, const double *X, int64_t T, int64_t n, int64_t ld,
                              const uint8_t *membership, int64_t B, const int64_t *queries,
                              int64_t nqb, int j, int relax, int64_t *count_out);

#ifdef __cplusplus
}
#endif
#endif /* STATDEPTH_B200_H */
