// SVD of the REAL upper-bidiagonal B (from bidiag.cuh) by block one-sided Jacobi in real FP64 arithmetic -- second half
// of the replacement of scipy.linalg.svd (reference llckbdm/kbdm.py:166), used for the members the divide-and-conquer
// solver (bdc.cuh) flags as numerically rank deficient.  Round-robin pairs of
// 32-column blocks: Gram on DMMA -> two-sided Jacobi eigen-solve of the 64x64 Gram in shared memory (relative-accuracy
// preserving, Demmel-Veselic) -> DMMA update of the X and V panels; one real DMMA per 8x8x4 tile product.
//   X <- B, V <- I;  on exit X = L_b Sigma, V = R_b  with  B = L_b Sigma R_b^T;  then U0 = (Q L_b) Sigma (P R_b)^H.
#pragma once
#include "common.cuh"

#define J_B 32                            // columns per Jacobi block

// ---- working copy of the Hankel matrix: X = U^{shift} (zero padded to ld x mp) -----------------------------------
__global__ void hankel_init_kernel(cplx* X, long long stride, int ld, const int* mv, const int* nbv,
                                   const cplx* sig, const long long* sig_off, int shift) {
    const int b = blockIdx.y;
    const int m = mv[b], mp = nbv[b] * J_B;
    const cplx* c = sig + sig_off[b] + shift;
    cplx* Xb = X + (long long)b * stride;
    const long long total = (long long)ld * mp;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int i = (int)(idx % ld), j = (int)(idx / ld);
        Xb[idx] = (i < m && j < m) ? c[i + j] : mkc(0.0, 0.0);
    }
}

// round-robin pairing of n (even) players, round r in [0, n-1), pair q in [0, n/2)
__device__ __forceinline__ void rr_pair(int n, int r, int q, int& a, int& b) {
    if (q == 0) { a = n - 1; b = r; }
    else {
        a = (r + q) % (n - 1);
        b = (r - q + (n - 1)) % (n - 1);
    }
    if (a > b) { int t = a; a = b; b = t; }
}

__device__ __forceinline__ int pk(int r, int c) { return ((c * (c + 1)) >> 1) + r; }      // packed upper triangle, r <= c

// ---- end of sweep: convergence bookkeeping (device side only; the host never reads it back) ----------------------
__global__ void jacobi_sweep_end_kernel(unsigned long long* sweep_off, int* done, int batch, double conv2) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    if (!done[b]) {
        double off2 = __longlong_as_double((long long)sweep_off[b]);
        if (off2 < conv2) done[b] = 1;
    }
    sweep_off[b] = 0ull;
}

#define RJ_RC 64                          // rows per staged chunk
#define RJ_LDT 68                         // tile ld in doubles (= 4 mod 16: conflict-free LDS.64 fragment reads)
#define RJ_TILE (RJ_LDT * 64)             // doubles per tile buffer
#define RJ_LDJ 68
#define RJ_GRAM_SMEM (2 * RJ_TILE * 8)
#define RJ_UPD_SMEM ((RJ_LDJ * 64 + 2 * RJ_TILE) * 8)
#define RJE_THREADS 512
#define RJE_SMEM ((2080 + 65 * 64) * 8 + 63 * 32 * 2 + 528 * 2 + 2048)

struct RJacobiParams {
    double* X; double* V; long long stride; int ld;      // stride and ld in doubles
    const int* mv; const int* nbv;
    int round;
    unsigned long long* sweep_off;
    const int* done;
    double tol2;
    int inner_sweeps;
    double* Jws;          // [batch][pairs_max][64*64] sorted J per pair (column-major, ld 64)
    double* Gws;          // [batch][pairs_max][2080] packed upper triangle of the pair's Gram matrix
    double* offws;        // [batch][pairs_max] scaled off-diagonal^2 of the pair
    int* skip;            // [batch][pairs_max]
    int pairs_max;
};

__global__ void rsvd_init_kernel(double* X, double* V, long long stride, int ld, const int* mv, const int* nbv,
                                 const double* dws, const double* ews, const int* only) {
    const int b = blockIdx.y;
    if (only && !only[b]) return;          // member already solved by the divide-and-conquer path
    const int m = mv[b], mp = nbv[b] * J_B;
    double* Xb = X + (long long)b * stride;
    double* Vb = V + (long long)b * stride;
    const double* d = dws + (long long)b * ld;
    const double* e = ews + (long long)b * ld;
    const long long total = (long long)ld * mp;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int i = (int)(idx % ld), j = (int)(idx / ld);
        double x = 0.0, v = 0.0;
        if (i < m && j < m) {
            if (i == j) { x = d[j]; v = 1.0; }
            else if (j == i + 1) x = e[i];
        }
        Xb[idx] = x;
        Vb[idx] = v;
    }
}

// 16-byte cp.async of two consecutive rows (row even) with partial zero fill at the m boundary
__device__ __forceinline__ void cp_async16_rows(void* smem_dst, const double* gsrc, int row, int m, const double* safe) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    int sz = (row + 1 < m) ? 16 : ((row < m) ? 8 : 0);
    const void* src = sz ? (const void*)gsrc : (const void*)safe;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

__global__ void __launch_bounds__(256, 2) rjacobi_gram_kernel(RJacobiParams p) {
    const int b = blockIdx.y;
    if (p.done[b]) return;
    const int nb = p.nbv[b];
    if (p.round >= nb - 1 || (int)blockIdx.x >= nb / 2) return;
    const int m = p.mv[b];
    int bi, bj;
    rr_pair(nb, p.round, blockIdx.x, bi, bj);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    __shared__ double red[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int ld = p.ld;
    const double* Xb = p.X + (long long)b * p.stride;
    const long long colI = (long long)ld * (bi * J_B), colJ = (long long)ld * (bj * J_B);
    auto col_base = [&](int c) -> long long { return (c < J_B) ? colI + (long long)ld * c : colJ + (long long)ld * (c - J_B); };
    const long long pairidx = (long long)b * p.pairs_max + blockIdx.x;
    auto load_tile = [&](int buf, int r0) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int idx = tid + 256 * r;            // 32 row pairs x 64 cols
            int i2 = idx & 31, c = idx >> 5;
            int row = r0 + 2 * i2;
            cp_async16_rows(&tiles[buf * RJ_TILE + 2 * i2 + RJ_LDT * c], Xb + col_base(c) + row, row, m, Xb);
        }
        cp_async_commit();
    };
    const int nchunks = (m + RJ_RC - 1) / RJ_RC;
    int tR[5], tC[5];
    const int ntile = (warp < 4) ? 5 : 4;
    {
        int first = (warp < 4) ? 5 * warp : 20 + 4 * (warp - 4);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            int f = first + ((q < ntile) ? q : 0);
            int R = 0, rowlen = 8;
            while (f >= rowlen) { f -= rowlen; ++R; --rowlen; }
            tR[q] = R; tC[q] = R + f;
        }
    }
    double acc[5][2];
#pragma unroll
    for (int q = 0; q < 5; ++q) { acc[q][0] = 0.0; acc[q][1] = 0.0; }
    load_tile(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) { load_tile(buf ^ 1, (ch + 1) * RJ_RC); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const double* T = tiles + buf * RJ_TILE;
#pragma unroll 4
        for (int k = 0; k < RJ_RC; k += 4) {
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                if (q < ntile) {
                    const double a = T[RJ_LDT * (8 * tR[q] + g) + k + t];
                    const double bb = T[RJ_LDT * (8 * tC[q] + g) + k + t];
                    dmma(acc[q][0], acc[q][1], a, bb);
                }
            }
        }
        __syncthreads();
    }
    double* G = tiles;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        if (q < ntile) {
            int r = 8 * tR[q] + g, c = 8 * tC[q] + 2 * t;
            if (r <= c) G[pk(r, c)] = acc[q][0];
            if (r <= c + 1) G[pk(r, c + 1)] = acc[q][1];
        }
    }
    __syncthreads();
    double mx = 0.0;
    for (int idx = tid; idx < 64 * 64; idx += 256) {
        int r = idx & 63, c = idx >> 6;
        if (r < c) {
            double dd = G[pk(r, r)] * G[pk(c, c)];
            double o = G[pk(r, c)];
            double o2 = o * o;
            if (dd > 0.0) mx = fmax(mx, o2 / dd);
            else if (o2 > 0.0) mx = fmax(mx, 1.0);
        }
    }
    mx = block_max(mx, red);
    if (tid == 0) {
        atomicMax(&p.sweep_off[b], (unsigned long long)__double_as_longlong(mx));
        p.skip[pairidx] = (mx < p.tol2) ? 1 : 0;
        p.offws[pairidx] = mx;
    }
    if (mx < p.tol2) return;
    double* Gout = p.Gws + pairidx * 2080;
    for (int idx = tid; idx < 2080; idx += 256) Gout[idx] = G[idx];
}

__global__ void __launch_bounds__(RJE_THREADS, 2) rjacobi_eig_kernel(RJacobiParams p) {
    const int b = blockIdx.y;
    if (p.done[b]) return;
    const int nb = p.nbv[b];
    if (p.round >= nb - 1 || (int)blockIdx.x >= nb / 2) return;
    const long long pairidx = (long long)b * p.pairs_max + blockIdx.x;
    if (p.skip[pairidx]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* G = reinterpret_cast<double*>(smem_raw);                   // packed upper, 2080
    double* Jm = G + 2080;                                              // 64 x 64, ld 65
    double* rc = Jm + 65 * 64;                                          // 32
    double* rs = rc + 32;                                               // 32
    double* wv = rs + 32;                                               // 64
    int* perm = reinterpret_cast<int*>(wv + 64);                        // 64
    int* flags = perm + 64;                                             // 4
    unsigned char* ptab = reinterpret_cast<unsigned char*>(flags + 4);  // [63][32][2]
    unsigned char* btab = ptab + 63 * 32 * 2;                           // [528][2]
    const int tid = threadIdx.x;
    const double* Gin = p.Gws + pairidx * 2080;
    for (int idx = tid; idx < 2080; idx += RJE_THREADS) G[idx] = Gin[idx];
    for (int idx = tid; idx < 64 * 64; idx += RJE_THREADS) {
        int r = idx & 63, c = idx >> 6;
        Jm[r + 65 * c] = (r == c) ? 1.0 : 0.0;
    }
    for (int idx = tid; idx < 63 * 32; idx += RJE_THREADS) {
        int a, bb;
        rr_pair(64, idx >> 5, idx & 31, a, bb);
        ptab[2 * idx] = (unsigned char)a; ptab[2 * idx + 1] = (unsigned char)bb;
    }
    for (int blk = tid; blk < 528; blk += RJE_THREADS) {
        int bq = (int)((sqrtf(8.0f * blk + 1.0f) - 1.0f) * 0.5f);
        while (((bq + 1) * (bq + 2)) / 2 <= blk) ++bq;
        while ((bq * (bq + 1)) / 2 > blk) --bq;
        btab[2 * blk] = (unsigned char)(blk - (bq * (bq + 1)) / 2);
        btab[2 * blk + 1] = (unsigned char)bq;
    }
    const double pair_off2 = p.offws[pairidx];
    const int inner_cap = (pair_off2 < 1e-6) ? max(p.inner_sweeps, 2) : p.inner_sweeps;
    const double tol_in2 = 4e-30;
    __syncthreads();
    for (int sweep = 0; sweep < inner_cap; ++sweep) {
        if (tid == 0) flags[0] = 0;
        for (int step = 0; step < 63; ++step) {
            const unsigned char* pt = ptab + step * 64;
            __syncthreads();
            if (tid < 32) {
                const int pa = pt[2 * tid], pb = pt[2 * tid + 1];
                const double gpp = G[pk(pa, pa)], gqq = G[pk(pb, pb)], gpq = G[pk(pa, pb)];
                const double ab2 = gpq * gpq;
                double c = 1.0, s = 0.0;
                if (ab2 > tol_in2 * fabs(gpp * gqq) && ab2 > 0.0) {
                    const double d = gqq - gpp;
                    const double r = sqrt(fma(d, d, 4.0 * ab2));
                    const double u = ((d >= 0.0) ? 2.0 : -2.0) / (fabs(d) + r);
                    c = rsqrt(fma(u * u, ab2, 1.0));
                    s = c * u * gpq;
                    flags[0] = 1;
                }
                rc[tid] = c; rs[tid] = s;
            }
            __syncthreads();
            for (int blk = tid; blk < 528; blk += RJE_THREADS) {
                const int a = btab[2 * blk], bq = btab[2 * blk + 1];
                const int p1 = pt[2 * a], q1 = pt[2 * a + 1], p2 = pt[2 * bq], q2 = pt[2 * bq + 1];
                const double ca = rc[a], cb = rc[bq], sa = rs[a], sb = rs[bq];
                auto gi = [&](int r, int c) { return (r <= c) ? pk(r, c) : pk(c, r); };
                const int i00 = gi(p1, p2), i01 = gi(p1, q2), i10 = gi(q1, p2), i11 = gi(q1, q2);
                const double m00 = G[i00], m01 = G[i01], m10 = G[i10], m11 = G[i11];
                // columns: new_p = c*p - s*q ; new_q = s*p + c*q   (real rotation), then rows likewise
                const double n00 = cb * m00 - sb * m01, n01 = sb * m00 + cb * m01;
                const double n10 = cb * m10 - sb * m11, n11 = sb * m10 + cb * m11;
                double o00 = ca * n00 - sa * n10, o01 = ca * n01 - sa * n11;
                const double o10 = sa * n00 + ca * n10;
                double o11 = sa * n01 + ca * n11;
                if (a == bq) {
                    const bool rot = (ca != 1.0) || (sa != 0.0);
                    if (rot) o01 = 0.0;
                    G[pk(p1, p1)] = o00; G[pk(q1, q1)] = o11; G[pk(p1, q1)] = o01;
                } else {
                    G[i00] = o00; G[i01] = o01; G[i10] = o10; G[i11] = o11;
                }
            }
#pragma unroll
            for (int r = 0; r < 2048 / RJE_THREADS; ++r) {
                const int idx = tid + RJE_THREADS * r;
                const int row = idx & 63, bq = idx >> 6;
                const int p2 = pt[2 * bq], q2 = pt[2 * bq + 1];
                const double cb = rc[bq], sb = rs[bq];
                const double x = Jm[row + 65 * p2], y = Jm[row + 65 * q2];
                Jm[row + 65 * p2] = cb * x - sb * y;
                Jm[row + 65 * q2] = sb * x + cb * y;
            }
        }
        __syncthreads();
        const int rotated = flags[0];
        __syncthreads();
        if (!rotated) break;
    }
    if (tid < 64) wv[tid] = G[pk(tid, tid)];
    __syncthreads();
    if (tid < 64) {
        const double w = wv[tid];
        int rank = 0;
        for (int j = 0; j < 64; ++j) {
            const double wj = wv[j];
            rank += (wj > w) || (wj == w && j < tid);
        }
        perm[rank] = tid;
    }
    __syncthreads();
    double* Jout = p.Jws + pairidx * 4096;
    for (int idx = tid; idx < 64 * 64; idx += RJE_THREADS) {
        int r = idx & 63, c = idx >> 6;
        Jout[idx] = Jm[r + 65 * perm[c]];
    }
}

// blockIdx.z: 0 -> X panel, 1 -> V panel.   Mp <- Mp * J   (64-row chunks, real DMMA)
__global__ void __launch_bounds__(256, 2) rjacobi_update_kernel(RJacobiParams p) {
    const int b = blockIdx.y;
    if (p.done[b]) return;
    const int nb = p.nbv[b];
    if (p.round >= nb - 1 || (int)blockIdx.x >= nb / 2) return;
    if (p.skip[(long long)b * p.pairs_max + blockIdx.x]) return;
    const int m = p.mv[b];
    int bi, bj;
    rr_pair(nb, p.round, blockIdx.x, bi, bj);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Js = reinterpret_cast<double*>(smem_raw);                  // 64 x 64, ld 68
    double* tiles = Js + RJ_LDJ * 64;                                   // 2 x RJ_TILE
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int ld = p.ld;
    double* Mb = (blockIdx.z ? p.V : p.X) + (long long)b * p.stride;
    const long long colI = (long long)ld * (bi * J_B), colJ = (long long)ld * (bj * J_B);
    auto col_base = [&](int c) -> long long { return (c < J_B) ? colI + (long long)ld * c : colJ + (long long)ld * (c - J_B); };
    const double* Jin = p.Jws + ((long long)b * p.pairs_max + blockIdx.x) * 4096;
    for (int idx = tid; idx < 2048; idx += 256) {          // 16-byte copies of two rows
        int r2 = idx & 31, c = idx >> 5;
        cp_async16(&Js[2 * r2 + RJ_LDJ * c], Jin + 2 * r2 + 64 * c, true);
    }
    cp_async_commit();
    auto load_tile = [&](int buf, int r0) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int idx = tid + 256 * r;
            int i2 = idx & 31, c = idx >> 5;
            int row = r0 + 2 * i2;
            cp_async16_rows(&tiles[buf * RJ_TILE + 2 * i2 + RJ_LDT * c], Mb + col_base(c) + row, row, m, Mb);
        }
        cp_async_commit();
    };
    const int nchunks = (m + RJ_RC - 1) / RJ_RC;
    const int wr = warp >> 1, wc = warp & 1;          // warp tile: 16 rows (2 row tiles) x 32 cols (4 col tiles)
    load_tile(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) { load_tile(buf ^ 1, (ch + 1) * RJ_RC); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const double* T = tiles + buf * RJ_TILE;
        double acc[2][4][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
        const double* ap = T + (16 * wr + g) + RJ_LDT * t;
        const double* bp = Js + t + RJ_LDJ * (32 * wc + g);
#pragma unroll 4
        for (int k = 0; k < 64; k += 4) {
            double a[2], bb[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = ap[8 * i + RJ_LDT * k];
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = bp[k + RJ_LDJ * 8 * j];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], bb[j]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int row = ch * RJ_RC + 16 * wr + 8 * i + g;
            if (row < m) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int c = 32 * wc + 8 * j + 2 * t;
                    Mb[col_base(c) + row] = acc[i][j][0];
                    Mb[col_base(c + 1) + row] = acc[i][j][1];
                }
            }
        }
        __syncthreads();
    }
}

// column norms of X -> singular values sorted descending + permutation 
__global__ void __launch_bounds__(256) rsvd_finalize_kernel(const double* X, long long stride, int ld, const int* mv, const int* nbv,
                                                            double* sing_vals, long long sv_stride, int* perm_out, int npow2,
                                                            const int* only = nullptr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* key = reinterpret_cast<double*>(smem_raw);
    int* val = reinterpret_cast<int*>(key + npow2);
    const int b = blockIdx.x, m = mv[b], mp = nbv[b] * J_B;
    if (only && !only[b]) return;          // member already solved by the divide-and-conquer path
    const double* Xb = X + (long long)b * stride;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int j = warp; j < npow2; j += 8) {
        double s = 0.0;
        if (j < mp) {
            const double* col = Xb + (long long)ld * j;
            for (int i = lane; i < m; i += 32) s = fma(col[i], col[i], s);
            s = warp_sum(s);
        } else s = -1.0;
        if (lane == 0) { key[j] = s; val[j] = j; }
    }
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npow2; i += 256) {
                int ixj = i ^ j;
                if (ixj > i) {
                    bool desc = ((i & k) == 0);
                    double a = key[i], c = key[ixj];
                    bool sw = desc ? (a < c) : (a > c);
                    if (sw) { key[i] = c; key[ixj] = a; int tv = val[i]; val[i] = val[ixj]; val[ixj] = tv; }
                }
            }
            __syncthreads();
        }
    }
    for (int k = tid; k < m; k += 256) {
        sing_vals[(long long)b * sv_stride + k] = sqrt(fmax(key[k], 0.0));
        perm_out[(long long)b * ld + k] = val[k];
    }
}

// complex, scaled, permuted copies for the back-multiplication by Q and P:
//   Lpre[:,k] = X[:,perm k] * (dsqi_k / sigma_k)   (so that  Lt = Q * Lpre),   Rpre[:,k] = V[:,perm k] * dsqi_k  (Rs = P * Rpre)
// With scale_mode = 1 the columns are left unscaled except X / 1 (debug: X_dbg = Q * X, V_dbg = P * V).
__global__ void rsvd_gather_kernel(const double* X, const double* V, long long rstride, int ld, const int* mv, const int* lv,
                                   const double* sing_vals, long long sv_stride, const int* perm, double q,
                                   cplx* Lpre, cplx* Rpre, long long cstride, int* status, int scale_mode,
                                   const int* only = nullptr) {
    const int b = blockIdx.y, k = blockIdx.x;
    const int m = mv[b], l = scale_mode ? m : lv[b];
    if (k >= l) return;
    if (only && !only[b]) return;
    const int src = perm[(long long)b * ld + k];
    double fx = 1.0, fv = 1.0;
    if (!scale_mode) {
        const double s = sing_vals[(long long)b * sv_stride + k];
        const double gq = (q > 0.0) ? (s + q * q / s) : s;
        if (!(gq > 0.0) || !isfinite(gq)) {
            if (threadIdx.x == 0) atomicMax(&status[b], 2);
            fx = 0.0; fv = 0.0;
        } else {
            fv = 1.0 / sqrt(gq);
            fx = fv / s;
        }
    }
    const double* xs = X + (long long)b * rstride + (long long)ld * src;
    const double* vs = V + (long long)b * rstride + (long long)ld * src;
    cplx* ldst = Lpre + (long long)b * cstride + (long long)ld * k;
    cplx* rdst = Rpre + (long long)b * cstride + (long long)ld * k;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        ldst[i] = mkc(xs[i] * fx, 0.0);
        rdst[i] = mkc(vs[i] * fv, 0.0);
    }
}
