// Batched complex-FP64 SVD of the Hankel matrix U^{p-1} by block one-sided Jacobi (Hestenes) --
// replaces scipy.linalg.svd / LAPACK zgesdd at reference llckbdm/kbdm.py:166.
//
//   X <- U^{p-1} (working copy, m x mp, column-major), V <- I.
//   One "round" = nb/2 disjoint pairs of 32-column blocks (round-robin tournament); one CTA per
//   (member, pair):   G = Xp^H Xp (64x64, DMMA)  ->  two-sided Jacobi eigen-solve of G in shared memory
//   (J, relative-accuracy preserving: Demmel-Veselic)  ->  Xp <- Xp J, Vp <- Vp J (DMMA).
//   nb-1 rounds = one sweep; sweeps repeat until every pair is orthogonal to 1e-14 (scaled).
//   On exit  X = L*Sigma (columns), V = R; singular values = column norms, sorted descending.
#pragma once
#include "common.cuh"

#define J_B 32
#define J_P 64
#define J_RC 32
#define J_LDT_G 36
#define J_LDT_U 34
#define J_LDJ 68
#define J_TILE_ELEMS (36 * 64)
#define J_MAT_ELEMS (68 * 64)
#define J_SMEM_BYTES ((2 * J_TILE_ELEMS + 2 * J_MAT_ELEMS) * 16 + 2048)

struct JacobiParams {
    cplx* X; cplx* V; long long stride; int ld;
    const int* mv;        // rows (= m) per member
    const int* nbv;       // number of 32-column blocks per member (even, >= 2)
    int round;            // round index within the sweep
    unsigned long long* sweep_off;   // per member: max scaled off-diagonal^2 seen this sweep (double bits)
    const int* done;      // per member: converged flag
    double tol2;          // skip a pair when off^2 < tol2
    int inner_sweeps;     // cap on the two-sided Jacobi sweeps of the 64x64 Gram eigen-solve
};

// ---- init: X = Hankel(U^{shift}), V = I -----------------------------------------------------------
__global__ void svd_init_kernel(cplx* X, cplx* V, long long stride, int ld, const int* mv, const int* nbv,
                                const cplx* sig, const long long* sig_off, int shift) {
    const int b = blockIdx.y;
    const int m = mv[b], mp = nbv[b] * J_B;
    const cplx* c = sig + sig_off[b] + shift;
    cplx* Xb = X + (long long)b * stride;
    cplx* Vb = V + (long long)b * stride;
    const long long total = (long long)ld * mp;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int i = (int)(idx % ld), j = (int)(idx / ld);
        cplx x = mkc(0.0, 0.0), v = mkc(0.0, 0.0);
        if (i < m && j < m) x = c[i + j];
        if (i == j && i < m) v = mkc(1.0, 0.0);
        Xb[idx] = x;
        Vb[idx] = v;
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

__global__ void __launch_bounds__(256, 1) jacobi_step_kernel(JacobiParams p) {
    const int b = blockIdx.y;
    if (p.done[b]) return;
    const int nb = p.nbv[b];
    if (p.round >= nb - 1 || (int)blockIdx.x >= nb / 2) return;
    const int m = p.mv[b];
    int bi, bj;
    rr_pair(nb, p.round, blockIdx.x, bi, bj);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* tiles = reinterpret_cast<cplx*>(smem_raw);                // 2 x J_TILE_ELEMS
    cplx* G = tiles + 2 * J_TILE_ELEMS;                              // J_MAT_ELEMS (ld 68), later Jp
    cplx* Jm = G + J_MAT_ELEMS;                                      // J_MAT_ELEMS (ld 68)
    double* rc = reinterpret_cast<double*>(Jm + J_MAT_ELEMS);        // 32 c
    cplx* rs = reinterpret_cast<cplx*>(rc + 32);                     // 32 s
    double* wv = reinterpret_cast<double*>(rs + 32);                 // 64 eigenvalues
    int* perm = reinterpret_cast<int*>(wv + 64);                     // 64
    int* flags = perm + 64;                                          // [0] rotated flag
    double* red = reinterpret_cast<double*>(flags + 4);              // 32 scratch

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int ld = p.ld;
    cplx* Xb = p.X + (long long)b * p.stride;
    cplx* Vb = p.V + (long long)b * p.stride;
    const long long colI = (long long)ld * (bi * J_B), colJ = (long long)ld * (bj * J_B);

    auto col_base = [&](int c) -> long long { return (c < J_B) ? colI + (long long)ld * c : colJ + (long long)ld * (c - J_B); };

    // ---------------- phase 1: G = Xp^H Xp ----------------
    auto load_tile = [&](const cplx* M, int buf, int r0, int ldt) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int idx = tid + 256 * r;
            int i = idx & 31, c = idx >> 5;
            bool ok = (r0 + i) < m;
            const cplx* src = ok ? (M + col_base(c) + r0 + i) : M;
            cp_async16(&tiles[buf * J_TILE_ELEMS + i + ldt * c], src, ok);
        }
        cp_async_commit();
    };
    const int nchunks = (m + J_RC - 1) / J_RC;
    {
        const int wr = warp >> 1, wc = warp & 1;
        double acc[2][4][4];
        zero_acc<2, 4>(acc);
        load_tile(Xb, 0, 0, J_LDT_G);
        for (int ch = 0; ch < nchunks; ++ch) {
            const int buf = ch & 1;
            if (ch + 1 < nchunks) { load_tile(Xb, buf ^ 1, (ch + 1) * J_RC, J_LDT_G); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncthreads();
            const cplx* T = tiles + buf * J_TILE_ELEMS;
            warp_zmma<2, 4, true, false>(acc, T + J_LDT_G * (16 * wr), J_LDT_G, 1, T + J_LDT_G * (32 * wc), 1, J_LDT_G, J_RC);
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int r = 16 * wr + 8 * i + g, c = 32 * wc + 8 * j + 2 * t;
                G[r + J_LDJ * c] = mkc(acc[i][j][0], acc[i][j][2]);
                G[r + J_LDJ * (c + 1)] = mkc(acc[i][j][1], acc[i][j][3]);
            }
    }
    __syncthreads();
    // ---------------- phase 2: scaled off-diagonal measure ----------------
    {
        double mx = 0.0;
        for (int idx = tid; idx < 64 * 64; idx += 256) {
            int r = idx & 63, c = idx >> 6;
            if (r < c) {
                double dd = G[r + J_LDJ * r].x * G[c + J_LDJ * c].x;
                double o2 = cabs2(G[r + J_LDJ * c]);
                if (dd > 0.0) mx = fmax(mx, o2 / dd);
                else if (o2 > 0.0) mx = fmax(mx, 1.0);
            }
        }
        mx = block_max(mx, red);
        if (tid == 0) atomicMax(&p.sweep_off[b], (unsigned long long)__double_as_longlong(mx));
        if (mx < p.tol2) return;   // uniform: already orthogonal
    }
    // ---------------- phase 3: two-sided cyclic Jacobi on G, accumulate J ----------------
    for (int idx = tid; idx < 64 * 64; idx += 256) {
        int r = idx & 63, c = idx >> 6;
        Jm[r + J_LDJ * c] = mkc(r == c ? 1.0 : 0.0, 0.0);
    }
    if (tid < 64) G[tid + J_LDJ * tid].y = 0.0;
    __syncthreads();
    const double tol_in2 = 4e-30;   // (2e-15)^2
    for (int sweep = 0; sweep < p.inner_sweeps; ++sweep) {
        if (tid == 0) flags[0] = 0;
        for (int step = 0; step < 63; ++step) {
            __syncthreads();
            if (tid < 32) {
                int pa, pb;
                rr_pair(64, step, tid, pa, pb);
                double gpp = G[pa + J_LDJ * pa].x, gqq = G[pb + J_LDJ * pb].x;
                cplx gpq = G[pa + J_LDJ * pb];
                double ab2 = cabs2(gpq);
                double c = 1.0; cplx s = mkc(0.0, 0.0);
                if (ab2 > tol_in2 * fabs(gpp * gqq) && ab2 > 0.0) {
                    double ab = sqrt(ab2);
                    double zeta = (gqq - gpp) / (2.0 * ab);
                    double tt;
                    if (fabs(zeta) > 1e150) tt = 0.5 / zeta;
                    else if (zeta == 0.0) tt = 1.0;
                    else tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    c = 1.0 / sqrt(1.0 + tt * tt);
                    double f = c * tt / ab;
                    s = mkc(gpq.x * f, gpq.y * f);
                    flags[0] = 1;
                }
                rc[tid] = c; rs[tid] = s;
            }
            __syncthreads();
            // G <- R^H G R by 2x2 blocks (a = row pair, bq = col pair)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int idx = tid + 256 * r;
                int a = idx & 31, bq = idx >> 5;
                int p1, q1, p2, q2;
                rr_pair(64, step, a, p1, q1);
                rr_pair(64, step, bq, p2, q2);
                double ca = rc[a], cb = rc[bq];
                cplx sa = rs[a], sb = rs[bq];
                cplx m00 = G[p1 + J_LDJ * p2], m01 = G[p1 + J_LDJ * q2], m10 = G[q1 + J_LDJ * p2], m11 = G[q1 + J_LDJ * q2];
                // columns: new_p = c*p - conj(s)*q ; new_q = s*p + c*q
                cplx csb = cconj(sb);
                cplx n00 = csub(cscale(m00, cb), cmul(csb, m01));
                cplx n01 = cadd(cmul(sb, m00), cscale(m01, cb));
                cplx n10 = csub(cscale(m10, cb), cmul(csb, m11));
                cplx n11 = cadd(cmul(sb, m10), cscale(m11, cb));
                // rows: new_p = c*p - s*q ; new_q = conj(s)*p + c*q
                cplx csa = cconj(sa);
                cplx o00 = csub(cscale(n00, ca), cmul(sa, n10));
                cplx o01 = csub(cscale(n01, ca), cmul(sa, n11));
                cplx o10 = cadd(cmul(csa, n00), cscale(n10, ca));
                cplx o11 = cadd(cmul(csa, n01), cscale(n11, ca));
                if (a == bq) {
                    bool rot = (ca != 1.0) || (sa.x != 0.0) || (sa.y != 0.0);
                    if (rot) { o01 = mkc(0.0, 0.0); o10 = mkc(0.0, 0.0); }
                    o00.y = 0.0; o11.y = 0.0;
                }
                G[p1 + J_LDJ * p2] = o00; G[p1 + J_LDJ * q2] = o01; G[q1 + J_LDJ * p2] = o10; G[q1 + J_LDJ * q2] = o11;
            }
            // J <- J R (columns)
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                int idx = tid + 256 * r;
                int row = idx & 63, bq = idx >> 6;
                int p2, q2;
                rr_pair(64, step, bq, p2, q2);
                double cb = rc[bq];
                cplx sb = rs[bq];
                cplx x = Jm[row + J_LDJ * p2], y = Jm[row + J_LDJ * q2];
                Jm[row + J_LDJ * p2] = csub(cscale(x, cb), cmul(cconj(sb), y));
                Jm[row + J_LDJ * q2] = cadd(cmul(sb, x), cscale(y, cb));
            }
        }
        __syncthreads();
        int rotated = flags[0];
        __syncthreads();
        if (!rotated) break;
    }
    // ---------------- phase 4: sort eigenvalues descending, Jp = J[:, perm] (into G storage) ----------------
    if (tid < 64) wv[tid] = G[tid + J_LDJ * tid].x;
    __syncthreads();
    if (tid < 64) {
        double w = wv[tid];
        int rank = 0;
        for (int j = 0; j < 64; ++j) {
            double wj = wv[j];
            rank += (wj > w) || (wj == w && j < tid);
        }
        perm[rank] = tid;
    }
    __syncthreads();
    for (int idx = tid; idx < 64 * 64; idx += 256) {
        int r = idx & 63, c = idx >> 6;
        G[r + J_LDJ * c] = Jm[r + J_LDJ * perm[c]];
    }
    __syncthreads();
    // ---------------- phase 5: Xp <- Xp Jp, Vp <- Vp Jp ----------------
    {
        const int wr = warp >> 1, wc = warp & 1;
        for (int which = 0; which < 2; ++which) {
            cplx* Mb = which ? Vb : Xb;
            load_tile(Mb, 0, 0, J_LDT_U);
            for (int ch = 0; ch < nchunks; ++ch) {
                const int buf = ch & 1;
                if (ch + 1 < nchunks) { load_tile(Mb, buf ^ 1, (ch + 1) * J_RC, J_LDT_U); cp_async_wait<1>(); }
                else cp_async_wait<0>();
                __syncthreads();
                const cplx* T = tiles + buf * J_TILE_ELEMS;
                double acc[1][4][4];
                zero_acc<1, 4>(acc);
                warp_zmma<1, 4, false, false>(acc, T + 8 * wr, 1, J_LDT_U, G + J_LDJ * (32 * wc), 1, J_LDJ, J_P);
                const int row = ch * J_RC + 8 * wr + g;
                if (row < m) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        int c = 32 * wc + 8 * j + 2 * t;
                        Mb[col_base(c) + row] = mkc(acc[0][j][0], acc[0][j][2]);
                        Mb[col_base(c + 1) + row] = mkc(acc[0][j][1], acc[0][j][3]);
                    }
                }
                __syncthreads();
            }
        }
    }
}

// =================================================================================================
// Split variant (default): two kernels per round, each < 113 KB of shared memory so that TWO CTAs
// co-reside per SM and the shared-memory-bound inner eigen-solve of one pair overlaps the DMMA work of
// another.
//   jacobi_gram_kernel   : G = Xp^H Xp (DMMA, 16-row chunks) -> packed upper-triangular G in smem ->
//                          one two-sided cyclic Jacobi sweep exploiting Hermitian symmetry (only the
//                          2x2 blocks on/above the diagonal are updated) -> sorted J written to HBM.
//   jacobi_update_kernel : Xp <- Xp J or Vp <- Vp J (blockIdx.z selects the panel), DMMA, 16-row chunks.
// =================================================================================================
#define JS_RC 16
#define JS_LDT_G 20                       // Gram tile ld (16 rows + 4)
#define JS_LDT_U 18                       // update tile ld (= 2 mod 8)
#define JS_GTILE (JS_LDT_G * 64)
#define JS_UTILE (JS_LDT_U * 64)
#define JS_LDJI 65                        // ld of J inside the Gram/inner kernel (not an MMA operand there)
#define JS_GRAM_SMEM ((2 * JS_GTILE + JS_LDJI * 64) * 16 + 2048)
#define JS_UPD_SMEM ((J_MAT_ELEMS + 2 * JS_UTILE) * 16)

struct JacobiSplitParams {
    cplx* X; cplx* V; long long stride; int ld;
    const int* mv; const int* nbv;
    int round;
    unsigned long long* sweep_off;
    const int* done;
    double tol2;
    int inner_sweeps;
    cplx* Jws;            // [batch][pairs_max][64*64] sorted J per pair (column-major, ld 64)
    int* skip;            // [batch][pairs_max]
    int pairs_max;
};

__device__ __forceinline__ int pk(int r, int c) { return ((c * (c + 1)) >> 1) + r; }      // r <= c
__device__ __forceinline__ cplx pk_get(const cplx* G, int r, int c) {
    return (r <= c) ? G[pk(r, c)] : cconj(G[pk(c, r)]);
}
__device__ __forceinline__ void pk_set(cplx* G, int r, int c, cplx v) {
    if (r <= c) G[pk(r, c)] = v; else G[pk(c, r)] = cconj(v);
}

__global__ void __launch_bounds__(256, 2) jacobi_gram_kernel(JacobiSplitParams p) {
    const int b = blockIdx.y;
    if (p.done[b]) return;
    const int nb = p.nbv[b];
    if (p.round >= nb - 1 || (int)blockIdx.x >= nb / 2) return;
    const int m = p.mv[b];
    int bi, bj;
    rr_pair(nb, p.round, blockIdx.x, bi, bj);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* tiles = reinterpret_cast<cplx*>(smem_raw);                 // 2 x JS_GTILE ; later packed G (2080)
    cplx* G = tiles;
    cplx* Jm = tiles + 2 * JS_GTILE;                                  // 64 x 64, ld JS_LDJI
    double* rc = reinterpret_cast<double*>(Jm + JS_LDJI * 64);
    cplx* rs = reinterpret_cast<cplx*>(rc + 32);
    double* wv = reinterpret_cast<double*>(rs + 32);
    int* perm = reinterpret_cast<int*>(wv + 64);
    int* flags = perm + 64;
    double* red = reinterpret_cast<double*>(flags + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int ld = p.ld;
    const cplx* Xb = p.X + (long long)b * p.stride;
    const long long colI = (long long)ld * (bi * J_B), colJ = (long long)ld * (bj * J_B);
    auto col_base = [&](int c) -> long long { return (c < J_B) ? colI + (long long)ld * c : colJ + (long long)ld * (c - J_B); };
    int* skipflag = p.skip + (long long)b * p.pairs_max + blockIdx.x;
    cplx* Jout = p.Jws + ((long long)b * p.pairs_max + blockIdx.x) * 4096;

    auto load_tile = [&](int buf, int r0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int idx = tid + 256 * r;            // 16 x 64
            int i = idx & 15, c = idx >> 4;
            bool ok = (r0 + i) < m;
            const cplx* src = ok ? (Xb + col_base(c) + r0 + i) : Xb;
            cp_async16(&tiles[buf * JS_GTILE + i + JS_LDT_G * c], src, ok);
        }
        cp_async_commit();
    };
    const int nchunks = (m + JS_RC - 1) / JS_RC;
    {
        const int wr = warp >> 1, wc = warp & 1;
        double acc[2][4][4];
        zero_acc<2, 4>(acc);
        load_tile(0, 0);
        for (int ch = 0; ch < nchunks; ++ch) {
            const int buf = ch & 1;
            if (ch + 1 < nchunks) { load_tile(buf ^ 1, (ch + 1) * JS_RC); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncthreads();
            const cplx* T = tiles + buf * JS_GTILE;
            warp_zmma<2, 4, true, false>(acc, T + JS_LDT_G * (16 * wr), JS_LDT_G, 1, T + JS_LDT_G * (32 * wc), 1, JS_LDT_G, JS_RC);
            __syncthreads();
        }
        // packed upper triangle of G (tiles are dead now)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int r = 16 * wr + 8 * i + g, c = 32 * wc + 8 * j + 2 * t;
                if (r <= c) G[pk(r, c)] = mkc(acc[i][j][0], (r == c) ? 0.0 : acc[i][j][2]);
                if (r <= c + 1) G[pk(r, c + 1)] = mkc(acc[i][j][1], (r == c + 1) ? 0.0 : acc[i][j][3]);
            }
    }
    __syncthreads();
    double pair_off2 = 1.0;
    {
        double mx = 0.0;
        for (int idx = tid; idx < 64 * 64; idx += 256) {
            int r = idx & 63, c = idx >> 6;
            if (r < c) {
                double dd = G[pk(r, r)].x * G[pk(c, c)].x;
                double o2 = cabs2(G[pk(r, c)]);
                if (dd > 0.0) mx = fmax(mx, o2 / dd);
                else if (o2 > 0.0) mx = fmax(mx, 1.0);
            }
        }
        mx = block_max(mx, red);
        if (tid == 0) {
            atomicMax(&p.sweep_off[b], (unsigned long long)__double_as_longlong(mx));
            *skipflag = (mx < p.tol2) ? 1 : 0;
        }
        if (mx < p.tol2) return;
        pair_off2 = mx;
    }
    for (int idx = tid; idx < 64 * 64; idx += 256) {
        int r = idx & 63, c = idx >> 6;
        Jm[r + JS_LDJI * c] = mkc(r == c ? 1.0 : 0.0, 0.0);
    }
    __syncthreads();
    const double tol_in2 = 4e-30;
    // far from convergence one cyclic sweep per visit is the best trade (measured); once the pair is nearly
    // orthogonal the eigen-solve is made accurate (2 sweeps, early exit) so that the outer iteration stays quadratic
    const int inner_cap = (pair_off2 < 1e-6) ? max(p.inner_sweeps, 2) : p.inner_sweeps;
    for (int sweep = 0; sweep < inner_cap; ++sweep) {
        if (tid == 0) flags[0] = 0;
        for (int step = 0; step < 63; ++step) {
            __syncthreads();
            if (tid < 32) {
                int pa, pb;
                rr_pair(64, step, tid, pa, pb);
                double gpp = G[pk(pa, pa)].x, gqq = G[pk(pb, pb)].x;
                cplx gpq = G[pk(pa, pb)];
                double ab2 = cabs2(gpq);
                double c = 1.0; cplx s = mkc(0.0, 0.0);
                if (ab2 > tol_in2 * fabs(gpp * gqq) && ab2 > 0.0) {
                    double ab = sqrt(ab2);
                    double zeta = (gqq - gpp) / (2.0 * ab);
                    double tt;
                    if (fabs(zeta) > 1e150) tt = 0.5 / zeta;
                    else if (zeta == 0.0) tt = 1.0;
                    else tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    c = 1.0 / sqrt(1.0 + tt * tt);
                    double f = c * tt / ab;
                    s = mkc(gpq.x * f, gpq.y * f);
                    flags[0] = 1;
                }
                rc[tid] = c; rs[tid] = s;
            }
            __syncthreads();
            // G <- R^H G R on the 2x2 blocks (a <= bq) only; the mirrored blocks are implied by Hermitian symmetry
            for (int idx = tid; idx < 528; idx += 256) {
                // unrank idx -> (a, bq) with a <= bq, column-major over bq
                int bq = (int)((sqrtf(8.0f * idx + 1.0f) - 1.0f) * 0.5f);
                while (((bq + 1) * (bq + 2)) / 2 <= idx) ++bq;
                while ((bq * (bq + 1)) / 2 > idx) --bq;
                int a = idx - (bq * (bq + 1)) / 2;
                int p1, q1, p2, q2;
                rr_pair(64, step, a, p1, q1);
                rr_pair(64, step, bq, p2, q2);
                double ca = rc[a], cb = rc[bq];
                cplx sa = rs[a], sb = rs[bq];
                cplx m00 = pk_get(G, p1, p2), m01 = pk_get(G, p1, q2), m10 = pk_get(G, q1, p2), m11 = pk_get(G, q1, q2);
                cplx csb = cconj(sb);
                cplx n00 = csub(cscale(m00, cb), cmul(csb, m01));
                cplx n01 = cadd(cmul(sb, m00), cscale(m01, cb));
                cplx n10 = csub(cscale(m10, cb), cmul(csb, m11));
                cplx n11 = cadd(cmul(sb, m10), cscale(m11, cb));
                cplx csa = cconj(sa);
                cplx o00 = csub(cscale(n00, ca), cmul(sa, n10));
                cplx o01 = csub(cscale(n01, ca), cmul(sa, n11));
                cplx o10 = cadd(cmul(csa, n00), cscale(n10, ca));
                cplx o11 = cadd(cmul(csa, n01), cscale(n11, ca));
                if (a == bq) {
                    bool rot = (ca != 1.0) || (sa.x != 0.0) || (sa.y != 0.0);
                    if (rot) o01 = mkc(0.0, 0.0);
                    o00.y = 0.0; o11.y = 0.0;
                    G[pk(p1, p1)] = o00; G[pk(q1, q1)] = o11; G[pk(p1, q1)] = o01;      // p1 < q1
                } else {
                    pk_set(G, p1, p2, o00); pk_set(G, p1, q2, o01); pk_set(G, q1, p2, o10); pk_set(G, q1, q2, o11);
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                int idx = tid + 256 * r;
                int row = idx & 63, bq = idx >> 6;
                int p2, q2;
                rr_pair(64, step, bq, p2, q2);
                double cb = rc[bq];
                cplx sb = rs[bq];
                cplx x = Jm[row + JS_LDJI * p2], y = Jm[row + JS_LDJI * q2];
                Jm[row + JS_LDJI * p2] = csub(cscale(x, cb), cmul(cconj(sb), y));
                Jm[row + JS_LDJI * q2] = cadd(cmul(sb, x), cscale(y, cb));
            }
        }
        __syncthreads();
        int rotated = flags[0];
        __syncthreads();
        if (!rotated) break;
    }
    if (tid < 64) wv[tid] = G[pk(tid, tid)].x;
    __syncthreads();
    if (tid < 64) {
        double w = wv[tid];
        int rank = 0;
        for (int j = 0; j < 64; ++j) {
            double wj = wv[j];
            rank += (wj > w) || (wj == w && j < tid);
        }
        perm[rank] = tid;
    }
    __syncthreads();
    for (int idx = tid; idx < 64 * 64; idx += 256) {
        int r = idx & 63, c = idx >> 6;
        Jout[idx] = Jm[r + JS_LDJI * perm[c]];
    }
}

__global__ void __launch_bounds__(256, 2) jacobi_update_kernel(JacobiSplitParams p) {
    const int b = blockIdx.y;
    if (p.done[b]) return;
    const int nb = p.nbv[b];
    if (p.round >= nb - 1 || (int)blockIdx.x >= nb / 2) return;
    if (p.skip[(long long)b * p.pairs_max + blockIdx.x]) return;
    const int m = p.mv[b];
    int bi, bj;
    rr_pair(nb, p.round, blockIdx.x, bi, bj);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* Js = reinterpret_cast<cplx*>(smem_raw);                    // 64 x 64, ld 68
    cplx* tiles = Js + J_MAT_ELEMS;                                   // 2 x JS_UTILE
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int ld = p.ld;
    cplx* Mb = (blockIdx.z ? p.V : p.X) + (long long)b * p.stride;
    const long long colI = (long long)ld * (bi * J_B), colJ = (long long)ld * (bj * J_B);
    auto col_base = [&](int c) -> long long { return (c < J_B) ? colI + (long long)ld * c : colJ + (long long)ld * (c - J_B); };
    const cplx* Jin = p.Jws + ((long long)b * p.pairs_max + blockIdx.x) * 4096;
    for (int idx = tid; idx < 4096; idx += 256) {
        int r = idx & 63, c = idx >> 6;
        cp_async16(&Js[r + J_LDJ * c], Jin + idx, true);
    }
    cp_async_commit();
    auto load_tile = [&](int buf, int r0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int idx = tid + 256 * r;
            int i = idx & 15, c = idx >> 4;
            bool ok = (r0 + i) < m;
            const cplx* src = ok ? (Mb + col_base(c) + r0 + i) : Mb;
            cp_async16(&tiles[buf * JS_UTILE + i + JS_LDT_U * c], src, ok);
        }
        cp_async_commit();
    };
    const int nchunks = (m + JS_RC - 1) / JS_RC;
    const int wr = warp >> 2, wc = warp & 3;          // rows 8*wr (2 row tiles), cols 16*wc
    load_tile(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) { load_tile(buf ^ 1, (ch + 1) * JS_RC); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const cplx* T = tiles + buf * JS_UTILE;
        double acc[1][2][4];
        zero_acc<1, 2>(acc);
        warp_zmma<1, 2, false, false>(acc, T + 8 * wr, 1, JS_LDT_U, Js + J_LDJ * (16 * wc), 1, J_LDJ, J_P);
        const int row = ch * JS_RC + 8 * wr + g;
        if (row < m) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                int c = 16 * wc + 8 * j + 2 * t;
                Mb[col_base(c) + row] = mkc(acc[0][j][0], acc[0][j][2]);
                Mb[col_base(c + 1) + row] = mkc(acc[0][j][1], acc[0][j][3]);
            }
        }
        __syncthreads();
    }
}

// =================================================================================================
// Three-kernel variant (default):
//   jacobi_gram3_kernel  : pure DMMA Gram; writes the packed upper triangle of G (2080 complex) + the scaled
//                          off-diagonal measure of the pair to HBM.
//   jacobi_eig_kernel    : the 64x64 Hermitian eigen-solve, 1024 threads (one 2x2 block per thread), two CTAs per
//                          SM, pair/block index tables in shared memory, rotation from 1 sqrt + 1 div + 1 rsqrt.
//   jacobi_update_kernel : as above.
// =================================================================================================
#define JE_THREADS 512
#define JE_SMEM ((2080 + JS_LDJI * 64) * 16 + 63 * 32 * 2 + 528 * 2 + 2048)

__global__ void __launch_bounds__(256, 2) jacobi_gram3_kernel(JacobiSplitParams p, cplx* Gws, double* offws) {
    const int b = blockIdx.y;
    if (p.done[b]) return;
    const int nb = p.nbv[b];
    if (p.round >= nb - 1 || (int)blockIdx.x >= nb / 2) return;
    const int m = p.mv[b];
    int bi, bj;
    rr_pair(nb, p.round, blockIdx.x, bi, bj);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* tiles = reinterpret_cast<cplx*>(smem_raw);                 // 2 x JS_GTILE, later packed G
    __shared__ double red[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int ld = p.ld;
    const cplx* Xb = p.X + (long long)b * p.stride;
    const long long colI = (long long)ld * (bi * J_B), colJ = (long long)ld * (bj * J_B);
    auto col_base = [&](int c) -> long long { return (c < J_B) ? colI + (long long)ld * c : colJ + (long long)ld * (c - J_B); };
    const long long pairidx = (long long)b * p.pairs_max + blockIdx.x;
    auto load_tile = [&](int buf, int r0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int idx = tid + 256 * r;
            int i = idx & 15, c = idx >> 4;
            bool ok = (r0 + i) < m;
            const cplx* src = ok ? (Xb + col_base(c) + r0 + i) : Xb;
            cp_async16(&tiles[buf * JS_GTILE + i + JS_LDT_G * c], src, ok);
        }
        cp_async_commit();
    };
    const int nchunks = (m + JS_RC - 1) / JS_RC;
    // G is Hermitian: only the 36 upper-triangular 8x8 tiles (tile row R <= tile column C) are computed.
    // They are flattened row-major (R=0: C=0..7, R=1: C=1..7, ...) and dealt out 5,5,5,5,4,4,4,4 to the 8 warps.
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
    double acc[1][5][4];
    zero_acc<1, 5>(acc);
    load_tile(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) { load_tile(buf ^ 1, (ch + 1) * JS_RC); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const cplx* T = tiles + buf * JS_GTILE;
#pragma unroll
        for (int k = 0; k < JS_RC; k += 4) {
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                if (q < ntile) {
                    cplx a = T[JS_LDT_G * (8 * tR[q] + g) + k + t];
                    cplx bb = T[JS_LDT_G * (8 * tC[q] + g) + k + t];
                    // conj(a)*b : (ar - i ai)(br + i bi)
                    dmma(acc[0][q][0], acc[0][q][1], a.x, bb.x);
                    dmma(acc[0][q][2], acc[0][q][3], a.x, bb.y);
                    dmma(acc[0][q][0], acc[0][q][1], a.y, bb.y);
                    dmma(acc[0][q][2], acc[0][q][3], -a.y, bb.x);
                }
            }
        }
        __syncthreads();
    }
    cplx* G = tiles;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        if (q < ntile) {
            int r = 8 * tR[q] + g, c = 8 * tC[q] + 2 * t;
            if (r <= c) G[pk(r, c)] = mkc(acc[0][q][0], (r == c) ? 0.0 : acc[0][q][2]);
            if (r <= c + 1) G[pk(r, c + 1)] = mkc(acc[0][q][1], (r == c + 1) ? 0.0 : acc[0][q][3]);
        }
    }
    __syncthreads();
    double mx = 0.0;
    for (int idx = tid; idx < 64 * 64; idx += 256) {
        int r = idx & 63, c = idx >> 6;
        if (r < c) {
            double dd = G[pk(r, r)].x * G[pk(c, c)].x;
            double o2 = cabs2(G[pk(r, c)]);
            if (dd > 0.0) mx = fmax(mx, o2 / dd);
            else if (o2 > 0.0) mx = fmax(mx, 1.0);
        }
    }
    mx = block_max(mx, red);
    if (tid == 0) {
        atomicMax(&p.sweep_off[b], (unsigned long long)__double_as_longlong(mx));
        p.skip[pairidx] = (mx < p.tol2) ? 1 : 0;
        offws[pairidx] = mx;
    }
    if (mx < p.tol2) return;
    cplx* Gout = Gws + pairidx * 2080;
    for (int idx = tid; idx < 2080; idx += 256) Gout[idx] = G[idx];
}

__global__ void __launch_bounds__(JE_THREADS, 2) jacobi_eig_kernel(JacobiSplitParams p, const cplx* Gws, const double* offws) {
    const int b = blockIdx.y;
    if (p.done[b]) return;
    const int nb = p.nbv[b];
    if (p.round >= nb - 1 || (int)blockIdx.x >= nb / 2) return;
    const long long pairidx = (long long)b * p.pairs_max + blockIdx.x;
    if (p.skip[pairidx]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* G = reinterpret_cast<cplx*>(smem_raw);                      // packed upper, 2080
    cplx* Jm = G + 2080;                                               // 64 x 64, ld JS_LDJI
    double* rc = reinterpret_cast<double*>(Jm + JS_LDJI * 64);        // 32
    cplx* rs = reinterpret_cast<cplx*>(rc + 32);                       // 32
    double* wv = reinterpret_cast<double*>(rs + 32);                   // 64
    int* perm = reinterpret_cast<int*>(wv + 64);                       // 64
    int* flags = perm + 64;                                            // 4
    unsigned char* ptab = reinterpret_cast<unsigned char*>(flags + 4); // [63][32][2] pair table
    unsigned char* btab = ptab + 63 * 32 * 2;                          // [528][2] block table (a <= bq)
    const int tid = threadIdx.x;
    const cplx* Gin = Gws + pairidx * 2080;
    for (int idx = tid; idx < 2080; idx += JE_THREADS) G[idx] = Gin[idx];
    for (int idx = tid; idx < 64 * 64; idx += JE_THREADS) {
        int r = idx & 63, c = idx >> 6;
        Jm[r + JS_LDJI * c] = mkc(r == c ? 1.0 : 0.0, 0.0);
    }
    for (int idx = tid; idx < 63 * 32; idx += JE_THREADS) {
        int a, bb;
        rr_pair(64, idx >> 5, idx & 31, a, bb);
        ptab[2 * idx] = (unsigned char)a; ptab[2 * idx + 1] = (unsigned char)bb;
    }
    for (int blk = tid; blk < 528; blk += JE_THREADS) {
        int bq = (int)((sqrtf(8.0f * blk + 1.0f) - 1.0f) * 0.5f);
        while (((bq + 1) * (bq + 2)) / 2 <= blk) ++bq;
        while ((bq * (bq + 1)) / 2 > blk) --bq;
        btab[2 * blk] = (unsigned char)(blk - (bq * (bq + 1)) / 2);
        btab[2 * blk + 1] = (unsigned char)bq;
    }
    const double pair_off2 = offws[pairidx];
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
                const double gpp = G[pk(pa, pa)].x, gqq = G[pk(pb, pb)].x;
                const cplx gpq = G[pk(pa, pb)];
                const double ab2 = cabs2(gpq);
                double c = 1.0; cplx s = mkc(0.0, 0.0);
                if (ab2 > tol_in2 * fabs(gpp * gqq) && ab2 > 0.0) {
                    // tan(theta) = sign(d) 2|g| / (|d| + sqrt(d^2 + 4|g|^2)),  s = c * tan(theta) * g/|g|
                    const double d = gqq - gpp;
                    const double r = sqrt(fma(d, d, 4.0 * ab2));
                    const double u = ((d >= 0.0) ? 2.0 : -2.0) / (fabs(d) + r);
                    c = rsqrt(fma(u * u, ab2, 1.0));
                    const double f = c * u;
                    s = mkc(gpq.x * f, gpq.y * f);
                    flags[0] = 1;
                }
                rc[tid] = c; rs[tid] = s;
            }
            __syncthreads();
            for (int blk = tid; blk < 528; blk += JE_THREADS) {
                const int a = btab[2 * blk], bq = btab[2 * blk + 1];
                const int p1 = pt[2 * a], q1 = pt[2 * a + 1], p2 = pt[2 * bq], q2 = pt[2 * bq + 1];
                const double ca = rc[a], cb = rc[bq];
                const cplx sa = rs[a], sb = rs[bq];
                const cplx m00 = pk_get(G, p1, p2), m01 = pk_get(G, p1, q2), m10 = pk_get(G, q1, p2), m11 = pk_get(G, q1, q2);
                const cplx csb = cconj(sb);
                const cplx n00 = csub(cscale(m00, cb), cmul(csb, m01));
                const cplx n01 = cadd(cmul(sb, m00), cscale(m01, cb));
                const cplx n10 = csub(cscale(m10, cb), cmul(csb, m11));
                const cplx n11 = cadd(cmul(sb, m10), cscale(m11, cb));
                const cplx csa = cconj(sa);
                cplx o00 = csub(cscale(n00, ca), cmul(sa, n10));
                cplx o01 = csub(cscale(n01, ca), cmul(sa, n11));
                const cplx o10 = cadd(cmul(csa, n00), cscale(n10, ca));
                cplx o11 = cadd(cmul(csa, n01), cscale(n11, ca));
                if (a == bq) {
                    const bool rot = (ca != 1.0) || (sa.x != 0.0) || (sa.y != 0.0);
                    if (rot) o01 = mkc(0.0, 0.0);
                    o00.y = 0.0; o11.y = 0.0;
                    G[pk(p1, p1)] = o00; G[pk(q1, q1)] = o11; G[pk(p1, q1)] = o01;
                } else {
                    pk_set(G, p1, p2, o00); pk_set(G, p1, q2, o01); pk_set(G, q1, p2, o10); pk_set(G, q1, q2, o11);
                }
            }
#pragma unroll
            for (int r = 0; r < 2048 / JE_THREADS; ++r) {
                const int idx = tid + JE_THREADS * r;
                const int row = idx & 63, bq = idx >> 6;
                const int p2 = pt[2 * bq], q2 = pt[2 * bq + 1];
                const double cb = rc[bq];
                const cplx sb = rs[bq];
                const cplx x = Jm[row + JS_LDJI * p2], y = Jm[row + JS_LDJI * q2];
                Jm[row + JS_LDJI * p2] = csub(cscale(x, cb), cmul(cconj(sb), y));
                Jm[row + JS_LDJI * q2] = cadd(cmul(sb, x), cscale(y, cb));
            }
        }
        __syncthreads();
        const int rotated = flags[0];
        __syncthreads();
        if (!rotated) break;
    }
    if (tid < 64) wv[tid] = G[pk(tid, tid)].x;
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
    cplx* Jout = p.Jws + pairidx * 4096;
    for (int idx = tid; idx < 64 * 64; idx += JE_THREADS) {
        int r = idx & 63, c = idx >> 6;
        Jout[idx] = Jm[r + JS_LDJI * perm[c]];
    }
}

// ---- end of sweep: convergence bookkeeping -------------------------------------------------------
__global__ void jacobi_sweep_end_kernel(unsigned long long* sweep_off, int* done, int* n_active, int batch, double conv2) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    if (!done[b]) {
        double off2 = __longlong_as_double((long long)sweep_off[b]);
        if (off2 < conv2) done[b] = 1;
        else atomicAdd(n_active, 1);
    }
    sweep_off[b] = 0ull;
}

// ---- finalize: column norms -> singular values (sorted descending) + permutation ------------------
// one CTA (256 threads) per member; dynamic smem: npow2 * (8 + 4) bytes
__global__ void __launch_bounds__(256) svd_finalize_kernel(const cplx* X, long long stride, int ld, const int* mv, const int* nbv,
                                                           double* sing_vals, long long sv_stride, int* perm_out, int npow2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* key = reinterpret_cast<double*>(smem_raw);
    int* val = reinterpret_cast<int*>(key + npow2);
    const int b = blockIdx.x, m = mv[b], mp = nbv[b] * J_B;
    const cplx* Xb = X + (long long)b * stride;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int j = warp; j < npow2; j += 8) {
        double s = 0.0;
        if (j < mp) {
            const cplx* col = Xb + (long long)ld * j;
            for (int i = lane; i < m; i += 32) s += cabs2(col[i]);
            s = warp_sum(s);
        } else s = -1.0;    // padding sorts last
        if (lane == 0) { key[j] = s; val[j] = j; }
    }
    __syncthreads();
    // bitonic sort, descending by key
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

// ---- gather/scale: Rs[:,k] = V[:,perm k] * dsqi_k ; Lt[:,k] = X[:,perm k] * dsqi_k / sigma_k  (k < l) ---
// (reference kbdm.py:171-186: truncation to l, Tikhonov g = s + q^2/s, Dsqi = g^{-1/2})
__global__ void svd_gather_kernel(const cplx* X, const cplx* V, long long stride, int ld, const int* mv, const int* lv,
                                  const double* sing_vals, long long sv_stride, const int* perm, double q,
                                  cplx* Rs, cplx* Lt, int* status) {
    const int b = blockIdx.y, k = blockIdx.x;
    const int m = mv[b], l = lv[b];
    if (k >= l) return;
    const int src = perm[(long long)b * ld + k];
    const double s = sing_vals[(long long)b * sv_stride + k];
    double gq = (q > 0.0) ? (s + q * q / s) : s;
    double dsqi = 0.0, ls = 0.0;
    if (!(gq > 0.0) || !isfinite(gq)) {
        if (threadIdx.x == 0) atomicMax(&status[b], 2);
    } else {
        dsqi = 1.0 / sqrt(gq);
        ls = dsqi / s;
    }
    const cplx* xs = X + (long long)b * stride + (long long)ld * src;
    const cplx* vs = V + (long long)b * stride + (long long)ld * src;
    cplx* rd = Rs + (long long)b * stride + (long long)ld * k;
    cplx* ldst = Lt + (long long)b * stride + (long long)ld * k;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        rd[i] = cscale(vs[i], dsqi);
        ldst[i] = cscale(xs[i], ls);
    }
}
