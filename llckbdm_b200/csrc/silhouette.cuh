// Silhouette coefficients of the pooled, clustered line-list features -- replaces sklearn.metrics.silhouette_samples as called
// by reference llckbdm/llckbdm.py:291 (O(n^2) pairwise Euclidean distances per clustering, 4-dimensional features
// (Re mu, Im mu, A, 0), llckbdm.py:202-230).  Every label value (including HDBSCAN's noise label) is a cluster, as in sklearn.
//
// The host sorts the points by label; thread i then walks all points cluster by cluster (tiles staged in shared memory,
// segment boundaries are uniform across the CTA) and keeps only O(1) state: the running distance sum of the current cluster,
// a_i = sum / (n_own - 1) for its own cluster and b_i = min over the others of sum / n_c.
//   s_i = (b_i - a_i) / max(a_i, b_i),  0 for singleton clusters and when max(a, b) = 0   (sklearn semantics)
// Several clusterings of the same points are scored in one launch (blockIdx.y).
#pragma once
#include "common.cuh"

#define SIL_THREADS 256
#define SIL_TILE 256

// X: [n][4] features; order[c][n]: point indices sorted by label; seg[c][n+1]: cluster start offsets into order (nseg[c] + 1 valid);
// rank_of[c][n]: cluster index (0..nseg-1) of point order[c][k] is implied by seg; out[c][n] indexed by ORIGINAL point index.
__global__ void __launch_bounds__(SIL_THREADS) silhouette_kernel(const double* __restrict__ X, int n, const int* __restrict__ order,
                                                                 const int* __restrict__ seg, const int* __restrict__ nseg,
                                                                 const int* __restrict__ cluster_of, double* __restrict__ out) {
    __shared__ double4 tile[SIL_TILE];
    const int c = blockIdx.y;
    const int* ord = order + (long long)c * n;
    const int* sg = seg + (long long)c * (n + 1);
    const int ns = nseg[c];
    const int k = blockIdx.x * SIL_THREADS + threadIdx.x;      // position in sorted order
    const bool live = k < n;
    const int me = live ? ord[k] : 0;
    const double4 xi = live ? reinterpret_cast<const double4*>(X)[me] : make_double4(0, 0, 0, 0);
    const int myc = live ? cluster_of[(long long)c * n + k] : -1;
    double a = 0.0, b = INFINITY, sum = 0.0;
    int own = 1, s = 0;
    int next_end = (ns > 0) ? sg[1] : 0;
    // tiles of SIL_TILE consecutive sorted points; cluster boundaries fall anywhere inside a tile and are uniform across the CTA
    auto close_segment = [&]() {
        const int cnt_s = next_end - sg[s];
        if (s == myc) { own = cnt_s; a = (cnt_s > 1) ? sum / (double)(cnt_s - 1) : 0.0; }
        else if (cnt_s > 0) b = fmin(b, sum / (double)cnt_s);
        sum = 0.0;
        ++s;
        next_end = (s < ns) ? sg[s + 1] : 0x7fffffff;
    };
    for (int t0 = 0; t0 < n; t0 += SIL_TILE) {
        __syncthreads();
        if (t0 + threadIdx.x < n) tile[threadIdx.x] = reinterpret_cast<const double4*>(X)[ord[t0 + threadIdx.x]];
        __syncthreads();
        const int cnt = min(SIL_TILE, n - t0);
        for (int q = 0; q < cnt; ++q) {
            while (t0 + q == next_end && s < ns) close_segment();
            const double4 y = tile[q];
            const double dx = xi.x - y.x, dy = xi.y - y.y, dz = xi.z - y.z, dw = xi.w - y.w;
            sum += sqrt(fma(dx, dx, fma(dy, dy, fma(dz, dz, dw * dw))));
        }
    }
    while (s < ns) close_segment();
    if (live) {
        double sres = 0.0;
        if (own > 1 && ns > 1) {
            const double mx = fmax(a, b);
            sres = (mx > 0.0 && isfinite(mx)) ? (b - a) / mx : 0.0;
        }
        out[(long long)c * n + me] = sres;
    }
}
