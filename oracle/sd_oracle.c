/*
 * sd_oracle.c -- CPU restatement of statdepth's depth hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library; the product (statdepth_b200/) never does and fails loudly without its CUDA
 * extension.  Parity status: PINNED for band depth / modified band depth / L1 / Oja against
 * (i) the reference's two documented vectors (docs/index.md:20-42, :98-112) and (ii) outputs of
 * the unmodified Python reference generated in the build container by
 * tests/golden/make_golden.py (fixtures committed under tests/golden/).  Simplex containment is
 * pinned only up to the reference's third-party LP tolerance (scipy.optimize.linprog, see
 * DESIGN.md "Simplex tie band"): "parity unpinned" inside that band.
 *
 * Two families live here:
 *   *_enum   : the reference's own algorithm transcribed to C (enumerate J-subsets of the other
 *              curves, closed-interval min/max test per time point).  O(n^J T) per query.
 *   the rest : closed-form restatements (ranks / bitmasks) that give the SAME integer counts and
 *              run at BASELINE sizes, validated against *_enum and the Python reference.
 *
 * Layouts: univariate X[t*ld + j] (T time rows, n curves, row stride ld >= n), float64.
 *          multivariate F[(i*T + t)*d + c] (N curves, T rows, d channels).
 *          point cloud  P[i*d + c].
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define SDO_EXPORT __attribute__((visibility("default")))

typedef int64_t i64;
typedef uint64_t u64;

static inline i64 comb2(i64 m) { return m * (m - 1) / 2; }
static inline i64 comb3x(i64 m) {
    if (m < 3) return 0;
    __int128 v = (__int128)m * (m - 1) * (m - 2) / 6;
    return (i64)v;
}

/* --- tiny pthread parallel-for (libgomp is not in this image) ------------------------------ */
static int g_threads = 0;

SDO_EXPORT int sdo_num_threads(void) {
    if (g_threads > 0) return g_threads;
    long k = sysconf(_SC_NPROCESSORS_ONLN);
    return k > 0 ? (int)k : 1;
}

SDO_EXPORT void sdo_set_num_threads(int k) { g_threads = k > 0 ? k : 0; }

typedef void (*pf_body)(i64 i, void *ctx, void *scratch);
typedef struct {
    atomic_llong next;
    i64 n, chunk;
    pf_body body;
    void *ctx;
    size_t scratch_bytes;
    atomic_int err;
} pf_job;

static void *pf_worker(void *arg) {
    pf_job *job = (pf_job *)arg;
    void *scratch = NULL;
    if (job->scratch_bytes) {
        scratch = calloc(1, job->scratch_bytes);
        if (!scratch) { atomic_store(&job->err, -2); return NULL; }
    }
    for (;;) {
        const i64 lo = atomic_fetch_add(&job->next, job->chunk);
        if (lo >= job->n) break;
        const i64 hi = lo + job->chunk < job->n ? lo + job->chunk : job->n;
        for (i64 i = lo; i < hi; ++i) job->body(i, job->ctx, scratch);
    }
    free(scratch);
    return NULL;
}

/* runs body(i) for i in [0,n) on sdo_num_threads() threads; each thread gets a zeroed scratch */
static int parallel_for(i64 n, i64 chunk, pf_body body, void *ctx, size_t scratch_bytes) {
    pf_job job;
    atomic_init(&job.next, 0);
    atomic_init(&job.err, 0);
    job.n = n; job.chunk = chunk > 0 ? chunk : 1; job.body = body; job.ctx = ctx;
    job.scratch_bytes = scratch_bytes;
    int k = sdo_num_threads();
    if ((i64)k > n) k = (int)(n > 0 ? n : 1);
    if (k <= 1) { pf_worker(&job); return atomic_load(&job.err); }
    pthread_t *th = (pthread_t *)malloc((size_t)k * sizeof(pthread_t));
    if (!th) return -2;
    int started = 0;
    for (int i = 0; i < k; ++i)
        if (pthread_create(&th[i], NULL, pf_worker, &job) == 0) ++started; else break;
    if (started == 0) pf_worker(&job);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    free(th);
    return atomic_load(&job.err);
}

/* ------------------------------------------------------------------------------------------
 * (1) The reference algorithm itself: _univariate_band_depth (_functional.py:238-253) calling
 *     _r2_containment (_containment.py:68-80) for every J-subset of the other curves.
 *     Returns the integer numerator:
 *        strict : S_nj  = #subsets whose band contains the query at ALL T points (cnt // T)
 *        relaxed: sum over subsets of cnt (the reference sums cnt/T as floats; cnt is exact)
 *     for j = J only (the caller loops j = 2..J like _functional.py:238).
 * ------------------------------------------------------------------------------------------ */
SDO_EXPORT int sdo_band_counts_enum(const double *X, i64 T, i64 n, i64 ld, const i64 *q, i64 nq,
                                    int j, int relax, i64 *out) {
    if (j != 2 && j != 3) return -1;
    for (i64 qi = 0; qi < nq; ++qi) {
        const i64 c = q[qi];
        i64 acc = 0;
        for (i64 a = 0; a < n; ++a) {
            if (a == c) continue;
            for (i64 b = a + 1; b < n; ++b) {
                if (b == c) continue;
                if (j == 2) {
                    i64 cnt = 0;
                    for (i64 t = 0; t < T; ++t) {
                        const double xa = X[t * ld + a], xb = X[t * ld + b], xc = X[t * ld + c];
                        const double lo = xa < xb ? xa : xb, hi = xa < xb ? xb : xa;
                        if (lo <= xc && xc <= hi) ++cnt; /* closed interval, _containment.py:76 */
                    }
                    acc += relax ? cnt : (cnt / T); /* _containment.py:80 */
                } else {
                    for (i64 e = b + 1; e < n; ++e) {
                        if (e == c) continue;
                        i64 cnt = 0;
                        for (i64 t = 0; t < T; ++t) {
                            const double xa = X[t * ld + a], xb = X[t * ld + b], xe = X[t * ld + e];
                            const double xc = X[t * ld + c];
                            double lo = xa < xb ? xa : xb, hi = xa < xb ? xb : xa;
                            if (xe < lo) lo = xe;
                            if (xe > hi) hi = xe;
                            if (lo <= xc && xc <= hi) ++cnt;
                        }
                        acc += relax ? cnt : (cnt / T);
                    }
                }
            }
        }
        out[qi] = acc;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * (2) Closed form, relaxed: per time row, b = #others strictly below,
 *     a = #others strictly above;  term_j(t) = C(n-1,j) - C(b,j) - C(a,j).
 *     out_all[c] (length n) receives sum_t term for EVERY curve c.
 *     Optional rank output: ranks_b/ranks_a [T*n] int32 (may be NULL).
 * ------------------------------------------------------------------------------------------ */
typedef struct { double v; i64 i; } kv_t;
static int kv_cmp(const void *pa, const void *pb) {
    const double a = ((const kv_t *)pa)->v, b = ((const kv_t *)pb)->v;
    return (a > b) - (a < b);
}

typedef struct {
    const double *X; i64 T, n, ld; int j; i64 full; i64 *out_all; int32_t *rb, *ra;
    pthread_mutex_t mu;
} mbd_ctx;

/* one chunk = a block of time rows; per-thread scratch holds the sort buffer and local sums,
 * merged into out_all under a mutex at the end of every chunk (integer adds: order-free). */
static void mbd_body(i64 blk, void *vctx, void *scratch) {
    mbd_ctx *c = (mbd_ctx *)vctx;
    const i64 n = c->n, ROWS = 8;
    kv_t *row = (kv_t *)scratch;
    i64 *loc = (i64 *)((char *)scratch + (size_t)n * sizeof(kv_t));
    memset(loc, 0, (size_t)n * sizeof(i64));
    const i64 t0 = blk * ROWS, t1 = t0 + ROWS < c->T ? t0 + ROWS : c->T;
    for (i64 t = t0; t < t1; ++t) {
        for (i64 k = 0; k < n; ++k) { row[k].v = c->X[t * c->ld + k]; row[k].i = k; }
        qsort(row, (size_t)n, sizeof(kv_t), kv_cmp);
        i64 s = 0;
        while (s < n) {
            i64 e = s + 1;
            while (e < n && row[e].v == row[s].v) ++e; /* tie run [s,e) */
            const i64 b = s, a = n - e;
            const i64 term = (c->j == 2) ? c->full - comb2(b) - comb2(a) : c->full - comb3x(b) - comb3x(a);
            for (i64 k = s; k < e; ++k) {
                loc[row[k].i] += term;
                if (c->rb) c->rb[t * n + row[k].i] = (int32_t)b;
                if (c->ra) c->ra[t * n + row[k].i] = (int32_t)a;
            }
            s = e;
        }
    }
    pthread_mutex_lock(&c->mu);
    for (i64 k = 0; k < n; ++k) c->out_all[k] += loc[k];
    pthread_mutex_unlock(&c->mu);
}

SDO_EXPORT int sdo_mbd_counts_all(const double *X, i64 T, i64 n, i64 ld, int j, i64 *out_all,
                                  int32_t *ranks_b, int32_t *ranks_a) {
    if (j != 2 && j != 3) return -1;
    memset(out_all, 0, (size_t)n * sizeof(i64));
    mbd_ctx c = {X, T, n, ld, j, (j == 2) ? comb2(n - 1) : comb3x(n - 1), out_all, ranks_b, ranks_a,
                 PTHREAD_MUTEX_INITIALIZER};
    return parallel_for((T + 7) / 8, 1, mbd_body, &c, (size_t)n * (sizeof(kv_t) + sizeof(i64)));
}

/* ------------------------------------------------------------------------------------------
 * (3) Closed form, strict: Sb[c'][t] = [X[t,c'] < X[t,c]], Sa = [>];
 *     a J-subset violates at t iff all members are below or all are above; S_nj = #subsets that
 *     never violate.  Bit-packed over time, early exit.  Same integer as sdo_band_counts_enum.
 * ------------------------------------------------------------------------------------------ */
typedef struct { const double *X; i64 T, n, ld, W; const i64 *q; int j; i64 *out; } bd_ctx;

static void bd_body(i64 qi, void *vctx, void *scratch) {
    bd_ctx *x = (bd_ctx *)vctx;
    const i64 n = x->n, W = x->W, c = x->q[qi];
    u64 *Sb = (u64 *)scratch, *Sa = Sb + n * W;
    memset(Sb, 0, (size_t)(2 * n * W) * sizeof(u64));
    for (i64 t = 0; t < x->T; ++t) {
        const double xc = x->X[t * x->ld + c];
        const u64 bit = (u64)1 << (t & 63);
        const i64 w = t >> 6;
        for (i64 o = 0; o < n; ++o) {
            const double xo = x->X[t * x->ld + o];
            if (xo < xc) Sb[o * W + w] |= bit;
            if (xo > xc) Sa[o * W + w] |= bit;
        }
    }
    i64 cnt = 0;
    for (i64 a = 0; a < n; ++a) {
        if (a == c) continue;
        for (i64 b = a + 1; b < n; ++b) {
            if (b == c) continue;
            if (x->j == 2) {
                int ok = 1;
                for (i64 w = 0; w < W; ++w)
                    if ((Sb[a * W + w] & Sb[b * W + w]) | (Sa[a * W + w] & Sa[b * W + w])) { ok = 0; break; }
                cnt += ok;
            } else {
                for (i64 e = b + 1; e < n; ++e) {
                    if (e == c) continue;
                    int ok = 1;
                    for (i64 w = 0; w < W; ++w)
                        if ((Sb[a * W + w] & Sb[b * W + w] & Sb[e * W + w]) |
                            (Sa[a * W + w] & Sa[b * W + w] & Sa[e * W + w])) { ok = 0; break; }
                    cnt += ok;
                }
            }
        }
    }
    x->out[qi] = cnt;
}

SDO_EXPORT int sdo_bd_counts(const double *X, i64 T, i64 n, i64 ld, const i64 *q, i64 nq, int j,
                             i64 *out) {
    if (j != 2 && j != 3) return -1;
    bd_ctx c = {X, T, n, ld, (T + 63) / 64, q, j, out};
    return parallel_for(nq, 1, bd_body, &c, (size_t)(2 * n * c.W) * sizeof(u64));
}

/* ------------------------------------------------------------------------------------------
 * (4) L1 depth, _pointcloud.py:125-150: 1 - || sum_{o != p} (x_o - x_p)/||x_p - x_o|| || / n,
 *     sequential float64 accumulation in index order (as the reference's Python loop does).
 * ------------------------------------------------------------------------------------------ */
typedef struct { const double *P; i64 n, d; const i64 *q; double *out; } l1_ctx;

static void l1_body(i64 qi, void *vctx, void *scratch) {
    (void)scratch;
    l1_ctx *x = (l1_ctx *)vctx;
    const double *P = x->P;
    const i64 n = x->n, d = x->d, p = x->q[qi];
    double s[16];
    for (i64 c = 0; c < d; ++c) s[c] = 0.0;
    for (i64 o = 0; o < n; ++o) {
        if (o == p) continue;
        double nrm2 = 0.0, diff[16];
        for (i64 c = 0; c < d; ++c) {
            diff[c] = P[o * d + c] - P[p * d + c];
            const double back = P[p * d + c] - P[o * d + c]; /* norm(vec - other), _pointcloud.py:146 */
            nrm2 += back * back;
        }
        const double nrm = sqrt(nrm2);
        for (i64 c = 0; c < d; ++c) s[c] += diff[c] / nrm;
    }
    double tot = 0.0;
    for (i64 c = 0; c < d; ++c) tot += s[c] * s[c];
    x->out[qi] = 1.0 - sqrt(tot) / (double)n;
}

SDO_EXPORT int sdo_l1_depth(const double *P, i64 n, i64 d, const i64 *q, i64 nq, double *out) {
    if (d < 1 || d > 16) return -1;
    l1_ctx c = {P, n, d, q, out};
    return parallel_for(nq, 16, l1_body, &c, 0);
}

/* ------------------------------------------------------------------------------------------
 * (5) Closed-simplex membership with an absolute tolerance band.
 *     Reference: _is_in_simplex (_containment.py:138-176) = LP feasibility via scipy linprog
 *     (third-party, HiGHS in scipy 1.18.1; primal feasibility ~1e-7 absolute).  Restated as
 *        inside  <=>  p in conv(V) exactly (sign test)   OR   dist(p, conv(V)) <= tol
 *     where dist is the exact Euclidean distance to the hull of the d+1 vertices, obtained as
 *     the minimum over all vertex subsets S whose affine projection of p has non-negative
 *     barycentric coordinates (the closest point of a polytope lies in the relative interior
 *     of one of its faces).  Degenerate simplices (reference allows them) reduce to their
 *     lower-dimensional hulls automatically because near-singular subsets are skipped.
 *     The CUDA kernels implement the same sequence of IEEE operations (compiled -fmad=false).
 * ------------------------------------------------------------------------------------------ */
#define SDO_DEG_EPS 1e-12 /* relative Gram-determinant threshold below which a subset is "dependent" */

/* squared distance from p to the affine hull of (k+1) points v[0..k] (k<=3) in R^d (d<=3) when
 * the projection has all barycentric coordinates >= 0; returns +inf otherwise / if dependent. */
static double sub_dist2(const double *v[4], int k, const double *p, int d) {
    double e[3][3], r[3], G[3][3], g[3], mu[3];
    for (int c = 0; c < d; ++c) r[c] = p[c] - v[0][c];
    if (k == 0) {
        double s = 0.0;
        for (int c = 0; c < d; ++c) s += r[c] * r[c];
        return s;
    }
    for (int a = 0; a < k; ++a)
        for (int c = 0; c < d; ++c) e[a][c] = v[a + 1][c] - v[0][c];
    for (int a = 0; a < k; ++a) {
        g[a] = 0.0;
        for (int c = 0; c < d; ++c) g[a] += e[a][c] * r[c];
        for (int b = 0; b < k; ++b) {
            G[a][b] = 0.0;
            for (int c = 0; c < d; ++c) G[a][b] += e[a][c] * e[b][c];
        }
    }
    double det, scale;
    if (k == 1) {
        det = G[0][0];
        scale = G[0][0];
        if (!(det > 0.0)) return INFINITY;
        mu[0] = g[0] / det;
    } else if (k == 2) {
        det = G[0][0] * G[1][1] - G[0][1] * G[1][0];
        scale = G[0][0] * G[1][1];
        if (!(det > SDO_DEG_EPS * scale)) return INFINITY;
        mu[0] = (g[0] * G[1][1] - G[0][1] * g[1]) / det;
        mu[1] = (G[0][0] * g[1] - g[0] * G[1][0]) / det;
    } else {
        const double c00 = G[1][1] * G[2][2] - G[1][2] * G[2][1];
        const double c01 = G[1][0] * G[2][2] - G[1][2] * G[2][0];
        const double c02 = G[1][0] * G[2][1] - G[1][1] * G[2][0];
        det = G[0][0] * c00 - G[0][1] * c01 + G[0][2] * c02;
        scale = G[0][0] * G[1][1] * G[2][2];
        if (!(det > SDO_DEG_EPS * scale)) return INFINITY;
        const double d0 = g[0] * c00 - G[0][1] * (g[1] * G[2][2] - G[1][2] * g[2]) + G[0][2] * (g[1] * G[2][1] - G[1][1] * g[2]);
        const double d1 = G[0][0] * (g[1] * G[2][2] - G[1][2] * g[2]) - g[0] * c01 + G[0][2] * (G[1][0] * g[2] - g[1] * G[2][0]);
        const double d2 = G[0][0] * (G[1][1] * g[2] - g[1] * G[2][1]) - G[0][1] * (G[1][0] * g[2] - g[1] * G[2][0]) + g[0] * c02;
        mu[0] = d0 / det; mu[1] = d1 / det; mu[2] = d2 / det;
    }
    double l0 = 1.0;
    for (int a = 0; a < k; ++a) {
        if (!(mu[a] >= 0.0)) return INFINITY;
        l0 -= mu[a];
    }
    if (!(l0 >= 0.0)) return INFINITY;
    double s = 0.0;
    for (int c = 0; c < d; ++c) {
        double res = r[c];
        for (int a = 0; a < k; ++a) res -= mu[a] * e[a][c];
        s += res * res;
    }
    return s;
}

/* min over all non-empty vertex subsets of a (d+1)-vertex simplex */
static double hull_dist2(const double *V, int d, const double *p) {
    const int m = d + 1;
    double best = INFINITY;
    for (int mask = 1; mask < (1 << m); ++mask) {
        const double *v[4];
        int k = 0;
        for (int i = 0; i < m; ++i)
            if (mask & (1 << i)) v[k++] = V + i * d;
        if (k - 1 > d) continue;
        const double s = sub_dist2(v, k - 1, p, d);
        if (s < best) best = s;
    }
    return best;
}

static inline double orient2(const double *a, const double *b, const double *c) {
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]);
}
static inline double orient3(const double *a, const double *b, const double *c, const double *e) {
    const double ax = a[0] - e[0], ay = a[1] - e[1], az = a[2] - e[2];
    const double bx = b[0] - e[0], by = b[1] - e[1], bz = b[2] - e[2];
    const double cx = c[0] - e[0], cy = c[1] - e[1], cz = c[2] - e[2];
    return ax * (by * cz - bz * cy) - ay * (bx * cz - bz * cx) + az * (bx * cy - by * cx);
}

/* V: (d+1) x d row-major vertices.  d in {1,2,3}. */
SDO_EXPORT int sdo_in_simplex(const double *V, int d, const double *p, double tol) {
    if (d == 1) {
        const double lo = V[0] < V[1] ? V[0] : V[1], hi = V[0] < V[1] ? V[1] : V[0];
        return (p[0] >= lo - tol) && (p[0] <= hi + tol);
    }
    if (d == 2) {
        const double *a = V, *b = V + 2, *c = V + 4;
        const double D = orient2(a, b, c);
        const double lab = (b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]);
        const double lbc = (c[0] - b[0]) * (c[0] - b[0]) + (c[1] - b[1]) * (c[1] - b[1]);
        const double lca = (a[0] - c[0]) * (a[0] - c[0]) + (a[1] - c[1]) * (a[1] - c[1]);
        double lmax = lab > lbc ? lab : lbc;
        if (lca > lmax) lmax = lca;
        /* thickness |D|/Lmax > tol  <=>  D^2 > tol^2 * Lmax  : sign test is meaningful */
        if (D != 0.0 && D * D > tol * tol * lmax) {
            const double s = D > 0.0 ? 1.0 : -1.0;
            const double ea = s * orient2(p, b, c), eb = s * orient2(a, p, c), ec = s * orient2(a, b, p);
            if (ea >= 0.0 && eb >= 0.0 && ec >= 0.0) return 1;
            /* farther than tol outside one edge line => farther than tol from the triangle */
            if (ea < 0.0 && ea * ea > tol * tol * lbc) return 0;
            if (eb < 0.0 && eb * eb > tol * tol * lca) return 0;
            if (ec < 0.0 && ec * ec > tol * tol * lab) return 0;
        }
        return hull_dist2(V, 2, p) <= tol * tol;
    }
    if (d == 3) {
        const double *a = V, *b = V + 3, *c = V + 6, *e = V + 9;
        const double D = orient3(a, b, c, e);
        /* face areas^2*4 via cross products, used as the scale of each sub-determinant */
        const double *F[4][3] = {{b, c, e}, {a, c, e}, {a, b, e}, {a, b, c}};
        double A2[4], amax = 0.0;
        for (int f = 0; f < 4; ++f) {
            const double ux = F[f][1][0] - F[f][0][0], uy = F[f][1][1] - F[f][0][1], uz = F[f][1][2] - F[f][0][2];
            const double vx = F[f][2][0] - F[f][0][0], vy = F[f][2][1] - F[f][0][1], vz = F[f][2][2] - F[f][0][2];
            const double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
            A2[f] = cx * cx + cy * cy + cz * cz;
            if (A2[f] > amax) amax = A2[f];
        }
        /* longest squared edge: |D| > tol * L^2 means every height exceeds ~tol, so the face
         * areas are well conditioned and the sign test is meaningful (needle / sliver simplices,
         * e.g. the reference's own collinear fixture, go to the distance path instead) */
        const double *E[6][2] = {{a, b}, {a, c}, {a, e}, {b, c}, {b, e}, {c, e}};
        double l2max = 0.0;
        for (int k = 0; k < 6; ++k) {
            const double dx = E[k][0][0] - E[k][1][0], dy = E[k][0][1] - E[k][1][1], dz = E[k][0][2] - E[k][1][2];
            const double l2 = dx * dx + dy * dy + dz * dz;
            if (l2 > l2max) l2max = l2;
        }
        (void)amax;
        if (D != 0.0 && D * D > tol * tol * l2max * l2max) {
            const double s = D > 0.0 ? 1.0 : -1.0;
            const double e0 = s * orient3(p, b, c, e), e1 = s * orient3(a, p, c, e);
            const double e2 = s * orient3(a, b, p, e), e3 = s * orient3(a, b, c, p);
            if (e0 >= 0.0 && e1 >= 0.0 && e2 >= 0.0 && e3 >= 0.0) return 1;
            if (e0 < 0.0 && e0 * e0 > tol * tol * A2[0]) return 0;
            if (e1 < 0.0 && e1 * e1 > tol * tol * A2[1]) return 0;
            if (e2 < 0.0 && e2 * e2 > tol * tol * A2[2]) return 0;
            if (e3 < 0.0 && e3 * e3 > tol * tol * A2[3]) return 0;
        }
        return hull_dist2(V, 3, p) <= tol * tol;
    }
    return -1;
}

/* pointcloud simplicial depth numerator, _pointcloud.py:44-56: #(d+1)-subsets of the OTHER
 * points whose closed simplex contains p. */
typedef struct { const double *P; i64 n; int d; const i64 *q; double tol; i64 *out; } sc_ctx;

static void sc_body(i64 qi, void *vctx, void *scratch) {
    (void)scratch;
    sc_ctx *x = (sc_ctx *)vctx;
    const double *P = x->P;
    const i64 n = x->n, p = x->q[qi];
    const int d = x->d;
    const double tol = x->tol;
    i64 cnt = 0;
    double V[12];
    for (i64 a = 0; a < n; ++a) {
        if (a == p) continue;
        for (i64 b = a + 1; b < n; ++b) {
            if (b == p) continue;
            if (d == 1) {
                V[0] = P[a]; V[1] = P[b];
                cnt += sdo_in_simplex(V, 1, P + p, tol);
                continue;
            }
            for (i64 c = b + 1; c < n; ++c) {
                if (c == p) continue;
                if (d == 2) {
                    memcpy(V, P + a * 2, 16); memcpy(V + 2, P + b * 2, 16); memcpy(V + 4, P + c * 2, 16);
                    cnt += sdo_in_simplex(V, 2, P + p * 2, tol);
                    continue;
                }
                for (i64 e = c + 1; e < n; ++e) {
                    if (e == p) continue;
                    memcpy(V, P + a * 3, 24); memcpy(V + 3, P + b * 3, 24);
                    memcpy(V + 6, P + c * 3, 24); memcpy(V + 9, P + e * 3, 24);
                    cnt += sdo_in_simplex(V, 3, P + p * 3, tol);
                }
            }
        }
    }
    x->out[qi] = cnt;
}

SDO_EXPORT int sdo_simplicial_counts(const double *P, i64 n, int d, const i64 *q, i64 nq, double tol,
                                     i64 *out) {
    if (d < 1 || d > 3) return -1;
    sc_ctx c = {P, n, d, q, tol, out};
    return parallel_for(nq, 1, sc_body, &c, 0);
}

/* multivariate functional simplex depth numerator, _functional.py:257-286 with
 * _simplex_containment (_containment.py:105-136): over (d+1)-subsets S of the N-1 OTHER curves,
 *   strict : #S with the query inside the simplex at ALL T rows
 *   relaxed: sum over S of #rows inside     (reference: sum of cnt/T as floats). */
typedef struct { const double *F; i64 N, T; int d; const i64 *q; int relax; double tol; i64 *out; } sx_ctx;

static void sx_body(i64 qi, void *vctx, void *scratch) {
    (void)scratch;
    sx_ctx *x = (sx_ctx *)vctx;
    const double *F = x->F;
    const i64 N = x->N, T = x->T, c = x->q[qi];
    const int d = x->d, m = d + 1;
    i64 acc = 0, idx[4];
    double V[12];
    if (N - 1 < m) { x->out[qi] = 0; return; }
    for (int k = 0; k < m; ++k) idx[k] = k;
    for (;;) { /* m-subsets of the "others" list (positions skip c) in lexicographic order */
        i64 cnt = 0;
        for (i64 t = 0; t < T; ++t) {
            for (int k = 0; k < m; ++k) {
                const i64 o = idx[k] + (idx[k] >= c ? 1 : 0);
                memcpy(V + k * d, F + (o * T + t) * d, (size_t)d * sizeof(double));
            }
            if (sdo_in_simplex(V, d, F + (c * T + t) * d, x->tol)) ++cnt;
            else if (!x->relax) break;
        }
        acc += x->relax ? cnt : (cnt == T ? 1 : 0);
        int k = m - 1;
        while (k >= 0 && idx[k] == (N - 1) - m + k) --k;
        if (k < 0) break;
        ++idx[k];
        for (int r = k + 1; r < m; ++r) idx[r] = idx[r - 1] + 1;
    }
    x->out[qi] = acc;
}

SDO_EXPORT int sdo_simplex_depth_counts(const double *F, i64 N, i64 T, int d, const i64 *q, i64 nq,
                                        int relax, double tol, i64 *out) {
    if (d < 1 || d > 3) return -1;
    sx_ctx c = {F, N, T, d, q, relax, tol, out};
    return parallel_for(nq, 1, sx_body, &c, 0);
}

/* Oja "depth", _pointcloud.py:176-204: sum over d-subsets S of the pool (minus p) of
 * vol(conv(S + {p})) divided by hull_volume (computed by the caller: Qhull in the reference).
 * vol of a d-simplex = |det| / d!.  pool = indices enumerated (the reference enumerates
 * `to_compute` when given, _pointcloud.py:182-183,191-193). */
typedef struct { const double *P; int d; const i64 *q; const i64 *pool; i64 npool; double hv; double *out; } oja_ctx;

static void oja_body(i64 qi, void *vctx, void *scratch) {
    (void)scratch;
    oja_ctx *x = (oja_ctx *)vctx;
    const double *P = x->P;
    const i64 p = x->q[qi];
    double acc = 0.0;
    for (i64 ia = 0; ia < x->npool; ++ia) {
        const i64 a = x->pool[ia];
        if (a == p) continue;
        for (i64 ib = ia + 1; ib < x->npool; ++ib) {
            const i64 b = x->pool[ib];
            if (b == p) continue;
            if (x->d == 2) {
                acc += fabs(orient2(P + a * 2, P + b * 2, P + p * 2)) / 2.0;
            } else {
                for (i64 ic = ib + 1; ic < x->npool; ++ic) {
                    const i64 c = x->pool[ic];
                    if (c == p) continue;
                    acc += fabs(orient3(P + a * 3, P + b * 3, P + c * 3, P + p * 3)) / 6.0;
                }
            }
        }
    }
    x->out[qi] = acc / x->hv;
}

SDO_EXPORT int sdo_oja(const double *P, i64 n, int d, const i64 *q, i64 nq, const i64 *pool, i64 npool,
                       double hull_volume, double *out) {
    (void)n;
    if (d != 2 && d != 3) return -1;
    oja_ctx c = {P, d, q, pool, npool, hull_volume, out};
    return parallel_for(nq, 1, oja_body, &c, 0);
}

/* ---------------------------------------------------------------------------------------------
 * 2-D triangle counting with the tolerance band, O(n^2) per (query, time point): an independent
 * restatement of "dist(p, triangle) <= tol" (the predicate sdo_in_simplex applies to every subset,
 * which itself restates the LP of _is_in_simplex, _containment.py:138-176) used to pin counts at
 * sizes the enumeration cannot reach (BASELINE configs 4 and 5: C(4999,3), C(49999,3) subsets).
 * A triangle misses the closed disk D(p, tol) iff the open arcs of tangent directions
 *     A_i = (phi_i - alpha_i, phi_i + alpha_i),   alpha_i = acos(tol / |x_i - p|)
 * of its three vertices have a common point; the common arc starts at the start of exactly one class
 * of arcs with equal start, so  #missing = sum_classes C(g + c, 3) - C(c, 3), c = #arcs that contain
 * the class's start.  c is found here by a direct circular test against every other arc (no ranks, no
 * wrap bookkeeping).  Points within tol of p have no arc.  tol = 0: arcs are open half-turns, i.e. exact
 * closed triangles up to the rounding of atan2.
 * ------------------------------------------------------------------------------------------- */
#define SDO_TWO_PI 6.283185307179586476925286766559

static i64 arcs_count_one(const double *pts, i64 n, i64 stride, i64 off, i64 p, double tol, double *s, double *len) {
    const double px = pts[p * stride + off], py = pts[p * stride + off + 1];
    i64 m = 0; /* far points */
    for (i64 j = 0; j < n; ++j) {
        if (j == p) continue;
        const double dx = pts[j * stride + off] - px, dy = pts[j * stride + off + 1] - py;
        const double r = hypot(dx, dy);
        if (r <= tol || (dx == 0.0 && dy == 0.0)) continue;
        double phi = atan2(dy, dx);
        if (phi < 0.0) phi += SDO_TWO_PI;
        const double alpha = acos(tol / r);
        double st = phi - alpha;
        if (st < 0.0) st += SDO_TWO_PI;
        s[m] = st;
        len[m] = 2.0 * alpha;
        ++m;
    }
    i64 missing = 0;
    for (i64 i = 0; i < m; ++i) {
        i64 g = 0, c = 0;
        int rep = 1;
        for (i64 j = 0; j < m; ++j) {
            if (s[j] == s[i]) {
                ++g;
                if (j < i) rep = 0;
                continue;
            }
            double delta = s[i] - s[j];
            if (delta < 0.0) delta += SDO_TWO_PI;
            if (delta > 0.0 && delta < len[j]) ++c;
        }
        if (rep) missing += comb3x(g + c) - comb3x(c);
    }
    return comb3x(n - 1) - missing;
}

typedef struct { const double *pts; i64 n, stride, T; const i64 *q; double tol; i64 *out; } arcs_ctx;

static void arcs_body(i64 qi, void *vctx, void *scratch) {
    arcs_ctx *x = (arcs_ctx *)vctx;
    double *s = (double *)scratch, *len = s + x->n;
    i64 acc = 0;
    for (i64 t = 0; t < x->T; ++t) acc += arcs_count_one(x->pts, x->n, x->stride, 2 * t, x->q[qi], x->tol, s, len);
    x->out[qi] = acc;
}

/* point j of time point t at pts[j*stride + 2*t + {0,1}]: a point cloud is (stride = 2, T = 1), the relaxed
 * multivariate simplex depth numerator of F[N][T][2] is (stride = 2*T, T). */
SDO_EXPORT int sdo_triangle_counts_arcs(const double *pts, i64 n, i64 stride, i64 T, const i64 *q, i64 nq,
                                        double tol, i64 *out) {
    if (n < 1 || T < 1 || !(tol >= 0.0)) return -1;
    arcs_ctx c = {pts, n, stride, T, q, tol, out};
    return parallel_for(nq, 1, arcs_body, &c, (size_t)(2 * n) * sizeof(double));
}

/* ---------------------------------------------------------------------------------------------
 * Strict multivariate simplex depth numerator, d = 2, at sizes the plain transcription (sx_body) cannot reach
 * (BASELINE config 4: C(4999,3) triples per query).  Same definition -- #triples of other curves whose closed
 * triangle (tolerance band tol, sdo_in_simplex) contains the query at ALL T rows -- with a cheap exact pre-test
 * per (triple, row): the three edge functions of the triangle as seen from the query; all clearly of one sign ->
 * inside; one clearly negative beyond tol * (upper bound of the edge length) -> outside; anything else is decided
 * by sdo_in_simplex itself.  Checked against sdo_simplex_depth_counts in tests/test_oracle.py.
 * Work is split over (query, first vertex) so that a few queries use all host threads.
 * ------------------------------------------------------------------------------------------- */
typedef struct { const double *F; i64 N, T; const i64 *q; i64 nq; double tol; atomic_llong *acc; } sxf_ctx;

static inline int sxf_classify(double e0, double e1, double e2, double ra, double rb, double rc, double tol) {
    const double kE = 1e-12, pab = ra * rb, pbc = rb * rc, pca = rc * ra;
    if ((e0 > kE * pab && e1 > kE * pbc && e2 > kE * pca) || (e0 < -kE * pab && e1 < -kE * pbc && e2 < -kE * pca)) return 1;
    const double D = (e0 + e1) + e2, big = kE * (pab + pbc + pca);
    const double m0 = tol * (ra + rb) * (1.0 + 1e-9) + kE * pab, m1 = tol * (rb + rc) * (1.0 + 1e-9) + kE * pbc,
                 m2 = tol * (rc + ra) * (1.0 + 1e-9) + kE * pca;
    if (D > big && (e0 < -m0 || e1 < -m1 || e2 < -m2)) return -1;
    if (D < -big && (e0 > m0 || e1 > m1 || e2 > m2)) return -1;
    return 0;
}

static int sxf_inside(const double *F, i64 T, i64 t, i64 a, i64 b, i64 c, i64 qc, double tol) {
    const double *pa = F + (a * T + t) * 2, *pb = F + (b * T + t) * 2, *pc = F + (c * T + t) * 2, *pp = F + (qc * T + t) * 2;
    const double vax = pa[0] - pp[0], vay = pa[1] - pp[1], vbx = pb[0] - pp[0], vby = pb[1] - pp[1],
                 vcx = pc[0] - pp[0], vcy = pc[1] - pp[1];
    const double e0 = vax * vby - vay * vbx, e1 = vbx * vcy - vby * vcx, e2 = vcx * vay - vcy * vax;
    const double up = 1.0 + 1e-7;
    const int cls = sxf_classify(e0, e1, e2, sqrt(vax * vax + vay * vay) * up, sqrt(vbx * vbx + vby * vby) * up,
                                 sqrt(vcx * vcx + vcy * vcy) * up, tol);
    if (cls) return cls > 0;
    double V[6] = {pa[0], pa[1], pb[0], pb[1], pc[0], pc[1]};
    return sdo_in_simplex(V, 2, pp, tol);
}

static void sxf_body(i64 job, void *vctx, void *scratch) {
    (void)scratch;
    sxf_ctx *x = (sxf_ctx *)vctx;
    const i64 N = x->N, T = x->T, qi = job / N, a = job % N, qc = x->q[qi];
    if (a == qc) return;
    i64 cnt = 0;
    for (i64 b = a + 1; b < N; ++b) {
        if (b == qc) continue;
        for (i64 c = b + 1; c < N; ++c) {
            if (c == qc) continue;
            int all = 1;
            for (i64 t = 0; t < T && all; ++t) all = sxf_inside(x->F, T, t, a, b, c, qc, x->tol);
            cnt += all;
        }
    }
    atomic_fetch_add(&x->acc[qi], cnt);
}

SDO_EXPORT int sdo_simplex2_strict_fast(const double *F, i64 N, i64 T, const i64 *q, i64 nq, double tol, i64 *out) {
    if (N < 1 || T < 1 || !(tol >= 0.0)) return -1;
    atomic_llong *acc = (atomic_llong *)calloc((size_t)(nq > 0 ? nq : 1), sizeof(atomic_llong));
    if (!acc) return -2;
    sxf_ctx c = {F, N, T, q, nq, tol, acc};
    const int rc = parallel_for(nq * N, 1, sxf_body, &c, 0);
    for (i64 i = 0; i < nq; ++i) out[i] = (i64)atomic_load(&acc[i]);
    free(acc);
    return rc;
}
