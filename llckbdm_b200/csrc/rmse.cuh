// Frequency-domain RMSE scoring of candidate line lists -- replaces the per-candidate loop of reference
// llckbdm/min_rmse_kbdm.py:33-41 and llckbdm/metrics.py:7-17 (multi_fid synthesis, two FFTs, RMSE of the real parts).
//
// No FFT is needed: Re FFT(r)_k = FFT(r_e)_k with r_e[n] = (r[n] + conj(r[(N-n) mod N])) / 2, so by Parseval
//   mean_k (Re FFT(r)_k / sqrt N)^2 = (1/N) sum_n |r_e[n]|^2,      r = data - model.
// One CTA per candidate: the model FID sum_k A_k exp(-t/T2_k) exp(i(2 pi F_k t + PH_k)) is synthesised in shared memory
// (thread t evaluates exp/sincos once per component at n = t and walks n = t + 256 j with the complex ratio mu^256),
// the row filter of sampling.py:75-97 (A > tol and T2 > 0) is applied on the fly, then one block reduction.
#pragma once
#include "common.cuh"

#define RMSE_THREADS 256

__global__ void __launch_bounds__(RMSE_THREADS) rmse_kernel(const cplx* __restrict__ data, int N, double dwell,
                                                            const double* __restrict__ ll, long long ll_stride,
                                                            const int* __restrict__ nrows, int filter, double amp_tol,
                                                            double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* model = reinterpret_cast<cplx*>(smem_raw);       // N
    __shared__ double red[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int rows = nrows[b];
    const double* L = ll + (long long)b * ll_stride;
    for (int n = tid; n < N; n += RMSE_THREADS) model[n] = mkc(0.0, 0.0);
    const double t0 = tid * dwell, ts = RMSE_THREADS * dwell;
    const double twopi = 6.283185307179586476925286766559;
    int nvalid = 0;
    for (int k = 0; k < rows; ++k) {
        const double A = L[4 * k], T2 = L[4 * k + 1], F = L[4 * k + 2], PH = L[4 * k + 3];
        if (filter && !(A > amp_tol && T2 > 0.0)) continue;
        ++nvalid;
        double sn, cs;
        sincos(twopi * F * t0 + PH, &sn, &cs);
        const double mag = A * exp(-t0 / T2);
        cplx z = mkc(mag * cs, mag * sn);
        sincos(twopi * F * ts, &sn, &cs);
        const double ms = exp(-ts / T2);
        const cplx step = mkc(ms * cs, ms * sn);
        for (int n = tid; n < N; n += RMSE_THREADS) {        // each thread owns its own n's: no conflicts
            model[n] = cadd(model[n], z);
            z = cmul(z, step);
        }
    }
    __syncthreads();
    double sum = 0.0;
    for (int n = tid; n < N; n += RMSE_THREADS) {
        const int n2 = (n == 0) ? 0 : N - n;
        const cplx r = csub(data[n], model[n]), r2 = csub(data[n2], model[n2]);
        const double x = 0.5 * (r.x + r2.x), y = 0.5 * (r.y - r2.y);
        sum = fma(x, x, fma(y, y, sum));
    }
    sum = block_sum(sum, red);
    if (tid == 0) out[b] = (nvalid > 0) ? sqrt(sum / (double)N) : INFINITY;
}

// Long FIDs (N > RMSE_SMEM_POINTS: the whole model no longer fits in shared memory): the same Parseval sum, tiled over n.
// r_e[n] and r_e[N-n] are complex conjugates, so  sum_n |r_e[n]|^2 = sum_{n=0}^{N/2} w_n |r_e[n]|^2  with w_n = 1 for the
// self-paired points (n = 0, and n = N/2 when N is even) and 2 otherwise.  Per tile of RMSE_TILE positions the model is
// synthesised for the forward range [c0, c0 + TILE) and for the mirrored range [N - c0 - TILE + 1, N - c0], both walked in
// ascending n with the mu^256 recurrence restarted from a direct exp/sincos evaluation (no growing backward recurrence).
#define RMSE_SMEM_POINTS 12800
#define RMSE_TILE 4096
__global__ void __launch_bounds__(RMSE_THREADS) rmse_tiled_kernel(const cplx* __restrict__ data, int N, double dwell,
                                                                  const double* __restrict__ ll, long long ll_stride,
                                                                  const int* __restrict__ nrows, int filter, double amp_tol,
                                                                  double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* fwd = reinterpret_cast<cplx*>(smem_raw);         // RMSE_TILE: model[c0 + i]
    cplx* mir = fwd + RMSE_TILE;                           // RMSE_TILE: model[m0 + i], m0 = N - c0 - RMSE_TILE + 1
    __shared__ double red[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int rows = nrows[b];
    const double* L = ll + (long long)b * ll_stride;
    const double twopi = 6.283185307179586476925286766559, ts = RMSE_THREADS * dwell;
    const int half = N / 2;                                // positions 0 .. half are summed
    double sum = 0.0;
    int nvalid = 0;
    for (int c0 = 0; c0 <= half; c0 += RMSE_TILE) {
        const int m0 = N - c0 - RMSE_TILE + 1;             // may be negative in the last tile: those entries are never read
        __syncthreads();
        for (int i = tid; i < RMSE_TILE; i += RMSE_THREADS) { fwd[i] = mkc(0.0, 0.0); mir[i] = mkc(0.0, 0.0); }
        nvalid = 0;
        for (int k = 0; k < rows; ++k) {
            const double A = L[4 * k], T2 = L[4 * k + 1], F = L[4 * k + 2], PH = L[4 * k + 3];
            if (filter && !(A > amp_tol && T2 > 0.0)) continue;
            ++nvalid;
            double sn, cs;
            sincos(twopi * F * ts, &sn, &cs);
            const double ms = exp(-ts / T2);
            const cplx step = mkc(ms * cs, ms * sn);
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                cplx* dst = which ? mir : fwd;
                const int base = which ? m0 : c0;
                int i = tid;
                while (base + i < 0) i += RMSE_THREADS;    // skip the unused head of the last mirrored tile
                const double t = (double)(base + i) * dwell;
                sincos(twopi * F * t + PH, &sn, &cs);
                const double mag = A * exp(-t / T2);
                cplx z = mkc(mag * cs, mag * sn);
                for (; i < RMSE_TILE; i += RMSE_THREADS) {
                    dst[i] = cadd(dst[i], z);
                    z = cmul(z, step);
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < RMSE_TILE; i += RMSE_THREADS) {
            const int n = c0 + i;
            if (n > half) break;
            const int n2 = (n == 0) ? 0 : N - n;
            const cplx mod2 = (n == 0) ? fwd[0] : mir[RMSE_TILE - 1 - i];
            const cplx r = csub(data[n], fwd[i]), r2 = csub(data[n2], mod2);
            const double x = 0.5 * (r.x + r2.x), y = 0.5 * (r.y - r2.y);
            const double w = (n == 0 || 2 * n == N) ? 1.0 : 2.0;
            sum = fma(w, fma(x, x, y * y), sum);
        }
    }
    sum = block_sum(sum, red);
    if (tid == 0) out[b] = (nvalid > 0) ? sqrt(sum / (double)N) : INFINITY;
}

// Batched FID synthesis -- reference llckbdm/sig_gen.py:57-71 (multi_fid = sum over rows of fid(), sig_gen.py:27-54) for many
// parameter sets at once: out[b][n] = sum_k A exp(-t_n/T2) exp(i(2 pi F t_n + PH)),  t_n = n * dwell.  Every term is evaluated
// directly (no recurrence) in the reference's summation order, so the result agrees with numpy to the last few ulps.
// grid (ceil(N/256), batch).
__global__ void __launch_bounds__(256) multi_fid_kernel(const double* __restrict__ params, long long pstride, const int* __restrict__ nrows,
                                                        int N, double dwell, cplx* __restrict__ out) {
    const int b = blockIdx.y, n = blockIdx.x * 256 + threadIdx.x;
    if (n >= N) return;
    const double* P = params + (long long)b * pstride;
    const int rows = nrows[b];
    const double t = n * dwell, twopi = 6.283185307179586476925286766559;
    cplx acc = mkc(0.0, 0.0);
    for (int k = 0; k < rows; ++k) {
        const double A = P[4 * k], T2 = P[4 * k + 1], F = P[4 * k + 2], PH = P[4 * k + 3];
        double sn, cs;
        sincos(twopi * F * t + PH, &sn, &cs);
        const double mag = A * exp(-t / T2);
        acc = cadd(acc, mkc(mag * cs, mag * sn));
    }
    out[(long long)b * N + n] = acc;
}
