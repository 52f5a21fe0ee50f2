// Pooled, filtered line lists and their clustering features straight from the solver's output buffer -- replaces, for the LLC-KBDM
// driver, the host sequence of reference llckbdm/llckbdm.py:94-98:  np.concatenate(line_lists) -> filter_samples (sampling.py:75-97:
// keep rows with A > tol and T2 > 0) -> _transform_line_lists (llckbdm.py:202-230: (A, T2, F, PH) -> (Re mu, Im mu, A, 0) with
// mu = exp(i dwell (2 pi F + i/T2))).  One CTA per member; rows keep their order (stable compaction by warp ballots); member b
// writes at row offset[b] (exclusive prefix sum of the n_valid counts the solve epilogue produced).
#pragma once
#include "common.cuh"

__global__ void __launch_bounds__(256) pool_features_kernel(const double* __restrict__ ll, long long ll_stride, const int* __restrict__ nrows,
                                                            const long long* __restrict__ offset, double dwell, double amp_tol,
                                                            double* __restrict__ samples, double* __restrict__ features) {
    __shared__ int wcount[8];
    __shared__ int base;
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rows = nrows[b];
    const double* L = ll + (long long)b * ll_stride;
    if (tid == 0) base = 0;
    __syncthreads();
    const double twopi = 6.283185307179586476925286766559;
    for (int r0 = 0; r0 < rows; r0 += 256) {
        const int r = r0 + tid;
        double A = 0.0, T2 = 0.0, F = 0.0, PH = 0.0;
        bool keep = false;
        if (r < rows) {
            A = L[4 * r]; T2 = L[4 * r + 1]; F = L[4 * r + 2]; PH = L[4 * r + 3];
            keep = (A > amp_tol) && (T2 > 0.0);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wcount[warp] = __popc(bal);
        __syncthreads();
        int pre = base;
        for (int w = 0; w < warp; ++w) pre += wcount[w];
        if (keep) {
            const long long o = (offset[b] + pre + __popc(bal & ((1u << lane) - 1u))) * 4;
            samples[o] = A; samples[o + 1] = T2; samples[o + 2] = F; samples[o + 3] = PH;
            double sn, cs;
            sincos(dwell * twopi * F, &sn, &cs);
            const double mag = exp(-dwell / T2);
            features[o] = mag * cs; features[o + 1] = mag * sn; features[o + 2] = A; features[o + 3] = 0.0;
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += wcount[w]; base += t; }
        __syncthreads();
    }
}
