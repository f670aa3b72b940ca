// simplex_pred.cuh -- closed-simplex membership with an absolute tolerance band (device side).
//
// Replaces _is_in_simplex (statdepth/depth/calculations/_containment.py:138-176), which asks
// scipy.optimize.linprog (third-party; HiGHS in scipy 1.18, primal feasibility ~1e-7 absolute)
// whether  lambda >= 0, sum lambda = 1, P^T lambda = x  is feasible.  Restated as
//     inside  <=>  x in conv(V) by exact sign tests   OR   dist(x, conv(V)) <= tol
// with dist the Euclidean distance to the hull of the d+1 vertices: the minimum, over all vertex
// subsets whose affine projection of x has non-negative barycentric coordinates, of the residual
// norm.  Degenerate simplices (collinear / coplanar / repeated vertices -- the reference's own
// multivariate fixture is 100% degenerate, statdepth/testing/_generating.py:94-96) fall out
// naturally: near-singular subsets are skipped and their hull is covered by smaller subsets.
//
// The translation units that include this header are compiled with -fmad=false so that every
// operation is the same IEEE-754 double operation, in the same order, as oracle/sd_oracle.c
// (built with -ffp-contract=off): decisions are bit-for-bit those of the CPU oracle.
#pragma once
#include <math.h>

namespace sd {

#define SD_DEG_EPS 1e-12

// squared distance from p to aff(v[0..k]) if the projection has barycentrics >= 0, else +inf
template <int D>
__device__ __forceinline__ double sub_dist2(const double *const (&v)[4], const int k, const double *p) {
    double e[3][3], r[3], G[3][3], g[3], mu[3];
#pragma unroll
    for (int c = 0; c < D; ++c) r[c] = p[c] - v[0][c];
    if (k == 0) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) s += r[c] * r[c];
        return s;
    }
    for (int a = 0; a < k; ++a)
#pragma unroll
        for (int c = 0; c < D; ++c) e[a][c] = v[a + 1][c] - v[0][c];
    for (int a = 0; a < k; ++a) {
        g[a] = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) g[a] += e[a][c] * r[c];
        for (int b = 0; b < k; ++b) {
            G[a][b] = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) G[a][b] += e[a][c] * e[b][c];
        }
    }
    double det, scale;
    if (k == 1) {
        det = G[0][0];
        if (!(det > 0.0)) return INFINITY;
        mu[0] = g[0] / det;
    } else if (k == 2) {
        det = G[0][0] * G[1][1] - G[0][1] * G[1][0];
        scale = G[0][0] * G[1][1];
        if (!(det > SD_DEG_EPS * scale)) return INFINITY;
        mu[0] = (g[0] * G[1][1] - G[0][1] * g[1]) / det;
        mu[1] = (G[0][0] * g[1] - g[0] * G[1][0]) / det;
    } else {
        const double c00 = G[1][1] * G[2][2] - G[1][2] * G[2][1];
        const double c01 = G[1][0] * G[2][2] - G[1][2] * G[2][0];
        const double c02 = G[1][0] * G[2][1] - G[1][1] * G[2][0];
        det = G[0][0] * c00 - G[0][1] * c01 + G[0][2] * c02;
        scale = G[0][0] * G[1][1] * G[2][2];
        if (!(det > SD_DEG_EPS * scale)) return INFINITY;
        const double d0 = g[0] * c00 - G[0][1] * (g[1] * G[2][2] - G[1][2] * g[2]) +
                          G[0][2] * (g[1] * G[2][1] - G[1][1] * g[2]);
        const double d1 = G[0][0] * (g[1] * G[2][2] - G[1][2] * g[2]) - g[0] * c01 +
                          G[0][2] * (G[1][0] * g[2] - g[1] * G[2][0]);
        const double d2 = G[0][0] * (G[1][1] * g[2] - g[1] * G[2][1]) -
                          G[0][1] * (G[1][0] * g[2] - g[1] * G[2][0]) + g[0] * c02;
        mu[0] = d0 / det;
        mu[1] = d1 / det;
        mu[2] = d2 / det;
    }
    double l0 = 1.0;
    for (int a = 0; a < k; ++a) {
        if (!(mu[a] >= 0.0)) return INFINITY;
        l0 -= mu[a];
    }
    if (!(l0 >= 0.0)) return INFINITY;
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        double res = r[c];
        for (int a = 0; a < k; ++a) res -= mu[a] * e[a][c];
        s += res * res;
    }
    return s;
}

// V: (D+1) vertices of D doubles each
template <int D>
__device__ __noinline__ double hull_dist2(const double *V, const double *p) {
    constexpr int m = D + 1;
    double best = INFINITY;
    for (int mask = 1; mask < (1 << m); ++mask) {
        const double *v[4] = {V, V, V, V};
        int k = 0;
        for (int i = 0; i < m; ++i)
            if (mask & (1 << i)) v[k++] = V + i * D;
        if (k - 1 > D) continue;
        const double s = sub_dist2<D>(v, k - 1, p);
        if (s < best) best = s;
    }
    return best;
}

__device__ __forceinline__ double orient2(const double *a, const double *b, const double *c) {
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]);
}

__device__ __forceinline__ double orient3(const double *a, const double *b, const double *c, const double *e) {
    const double ax = a[0] - e[0], ay = a[1] - e[1], az = a[2] - e[2];
    const double bx = b[0] - e[0], by = b[1] - e[1], bz = b[2] - e[2];
    const double cx = c[0] - e[0], cy = c[1] - e[1], cz = c[2] - e[2];
    return ax * (by * cz - bz * cy) - ay * (bx * cz - bz * cx) + az * (bx * cy - by * cx);
}

__device__ __forceinline__ bool in_simplex1(const double *V, const double *p, const double tol) {
    const double lo = V[0] < V[1] ? V[0] : V[1], hi = V[0] < V[1] ? V[1] : V[0];
    return (p[0] >= lo - tol) && (p[0] <= hi + tol);
}

__device__ __forceinline__ bool in_simplex2(const double *V, const double *p, const double tol) {
    const double *a = V, *b = V + 2, *c = V + 4;
    const double D = orient2(a, b, c);
    const double lab = (b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]);
    const double lbc = (c[0] - b[0]) * (c[0] - b[0]) + (c[1] - b[1]) * (c[1] - b[1]);
    const double lca = (a[0] - c[0]) * (a[0] - c[0]) + (a[1] - c[1]) * (a[1] - c[1]);
    double lmax = lab > lbc ? lab : lbc;
    if (lca > lmax) lmax = lca;
    if (D != 0.0 && D * D > tol * tol * lmax) {
        const double s = D > 0.0 ? 1.0 : -1.0;
        const double ea = s * orient2(p, b, c), eb = s * orient2(a, p, c), ec = s * orient2(a, b, p);
        if (ea >= 0.0 && eb >= 0.0 && ec >= 0.0) return true;
        if (ea < 0.0 && ea * ea > tol * tol * lbc) return false;
        if (eb < 0.0 && eb * eb > tol * tol * lca) return false;
        if (ec < 0.0 && ec * ec > tol * tol * lab) return false;
    }
    return hull_dist2<2>(V, p) <= tol * tol;
}

__device__ __forceinline__ bool in_simplex3(const double *V, const double *p, const double tol) {
    const double *a = V, *b = V + 3, *c = V + 6, *e = V + 9;
    const double D = orient3(a, b, c, e);
    const double *F[4][3] = {{b, c, e}, {a, c, e}, {a, b, e}, {a, b, c}};
    double A2[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        const double ux = F[f][1][0] - F[f][0][0], uy = F[f][1][1] - F[f][0][1], uz = F[f][1][2] - F[f][0][2];
        const double vx = F[f][2][0] - F[f][0][0], vy = F[f][2][1] - F[f][0][1], vz = F[f][2][2] - F[f][0][2];
        const double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
        A2[f] = cx * cx + cy * cy + cz * cz;
    }
    const double *E[6][2] = {{a, b}, {a, c}, {a, e}, {b, c}, {b, e}, {c, e}};
    double l2max = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double dx = E[k][0][0] - E[k][1][0], dy = E[k][0][1] - E[k][1][1], dz = E[k][0][2] - E[k][1][2];
        const double l2 = dx * dx + dy * dy + dz * dz;
        if (l2 > l2max) l2max = l2;
    }
    if (D != 0.0 && D * D > tol * tol * l2max * l2max) {
        const double s = D > 0.0 ? 1.0 : -1.0;
        const double e0 = s * orient3(p, b, c, e), e1 = s * orient3(a, p, c, e);
        const double e2 = s * orient3(a, b, p, e), e3 = s * orient3(a, b, c, p);
        if (e0 >= 0.0 && e1 >= 0.0 && e2 >= 0.0 && e3 >= 0.0) return true;
        if (e0 < 0.0 && e0 * e0 > tol * tol * A2[0]) return false;
        if (e1 < 0.0 && e1 * e1 > tol * tol * A2[1]) return false;
        if (e2 < 0.0 && e2 * e2 > tol * tol * A2[2]) return false;
        if (e3 < 0.0 && e3 * e3 > tol * tol * A2[3]) return false;
    }
    return hull_dist2<3>(V, p) <= tol * tol;
}

template <int D>
__device__ __forceinline__ bool in_simplex(const double *V, const double *p, const double tol) {
    if (D == 1) return in_simplex1(V, p, tol);
    if (D == 2) return in_simplex2(V, p, tol);
    return in_simplex3(V, p, tol);
}

}  // namespace sd
