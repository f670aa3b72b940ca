// angular_key.cuh -- exact-monotone angular keys of 2-D directions, shared by the triangle counting
// (simplicial_count.cu) and the Oja sums (oja_count.cu).  See simplicial_count.cu for the derivation.
#pragma once
#include "common.cuh"

namespace sd {

constexpr double SC_ZERO = 8.0 + 9.0;   // direction is the zero vector (point coincides with the query)
constexpr double SC_SELF = 8.0 + 10.0;  // the query's own entry

// instance -> (query index, offset of its coordinates); mode 0: point cloud, mode 1: functional (q, t)
struct ScGeom {
    const double *pts;   // point j of instance i at pts[j * stride_j + inst_off(i) + {0,1}]
    i64 stride_j;
    const i64 *qidx;     // query ids (may be null = identity)
    i64 T;               // functional: instances per query (time points); point cloud: 1
    i64 inst0;           // first instance of this batch
};

__device__ __forceinline__ i64 sc_query(const ScGeom &g, i64 inst) {
    const i64 qi = inst / g.T;
    return g.qidx ? g.qidx[qi] : qi;
}
__device__ __forceinline__ i64 sc_off(const ScGeom &g, i64 inst) { return (inst % g.T) * 2; }

// exact-monotone angular key of a non-zero direction, in [8, 16)
__device__ __forceinline__ double sc_key(double dx, double dy) {
    double add = 8.0;
    if (dy < 0.0 || (dy == 0.0 && dx < 0.0)) {  // lower half-plane (and the negative x axis): rotate by pi
        dx = -dx;
        dy = -dy;
        add = 12.0;
    }
    // now dy > 0, or dy == 0 and dx > 0: angle in [0, pi)
    double oct, f;
    if (dx > 0.0) {
        if (dy < dx) { oct = 0.0; f = dy / dx; }            // [0, pi/4)
        else         { oct = 1.0; f = 1.0 - dx / dy; }      // [pi/4, pi/2)
    } else {
        const double ax = -dx;
        if (ax < dy) { oct = 2.0; f = ax / dy; }            // [pi/2, 3pi/4)
        else         { oct = 3.0; f = 1.0 - dy / ax; }      // [3pi/4, pi)
    }
    // f in [0,1): one rounding to the [8,16) binade, monotone.  A direction a hair below the +x axis
    // (f = 1 - tiny rounds to 1, or 15 + f rounds up) would land on 16.0, the value range of the sentinels:
    // angle 2 pi IS angle 0, so it folds back onto 8.0 (its antipode is then exactly 12.0).
    const double k = (add + oct) + f;
    return k >= 16.0 ? 8.0 : k;
}

}  // namespace sd
