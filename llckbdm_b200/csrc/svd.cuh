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
