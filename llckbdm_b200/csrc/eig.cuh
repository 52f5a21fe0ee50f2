// Batched complex-FP64 nonsymmetric eigensolver for the reduced operator U_red (l x l) -- replaces
// scipy.linalg.eig / LAPACK zgeev at reference llckbdm/kbdm.py:192.  One CTA per ensemble member:
//   hess_panel_kernel : blocked (compact-WY) Householder reduction H = Q^H U Q; trailing updates and the backward
//                       accumulation of Q are batched DMMA GEMMs issued by the host driver
//   hqr_kernel        : small-bulge multishift QR (single-shift complex bulges, spacing 2, chased in
//                       lockstep by one warp each inside a 64x64 shared-memory window; the accumulated
//                       unitary W is applied to the off-window strips of H and to Z with DMMA GEMMs)
//   trevc_diag_kernel : eigenvectors of the triangular Schur factor by block back substitution (+ one DMMA GEMM per block)
// Eigenvector scale/phase is irrelevant downstream, so LAPACK's normalisation is not reproduced.
#pragma once
#include "common.cuh"

#define E_THREADS 512
#define E_NWARPS 16
#define E_W 64          // window size
#define E_NB 16         // shifts (bulges) per multishift sweep
#define E_LDW 68        // ld of Ww in smem (= 4 mod 8: conflict-free DMMA fragment reads)
#define E_LDH 65        // ld of Hw in smem (odd: row rotations, one lane per COLUMN, hit 8 distinct 16-byte bank groups)
#define E_MAT (E_LDW * E_W)
#define E_TILE (68 * 32)   // strip tile: [68 x 32] (row strips) or [34 x 64] (col strips)
#define E_LDS 17        // ld of the shift scratch matrix
#define HQR_SMEM_BYTES ((2 * E_MAT + 2 * E_TILE + E_LDS * E_NB + 64) * 16 + 1024)

// Small batches run the one-CTA-per-member kernels as a thread-block cluster of csize CTAs per member: every CTA executes the cheap
// O(n) bookkeeping redundantly (bit-identical, so redundant global writes are benign) and the O(n^2) pass is split across the cluster.
// Barrier with release/acquire ordering of global-memory writes inside the cluster.
__device__ __forceinline__ void cluster_barrier(int csize) {
    if (csize > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    }
}


// ---------------------------------------------------------------------------------------------
// Blocked (compact-WY) Hessenberg reduction: panel kernel.
// For the panel of HB_NB columns starting at k0 it generates reflectors v_j (P_j = I - tau_j v_j v_j^H), the upper
// triangular T (Q = P_0..P_{nb-1} = I - V T V^H) and Y = A V T, reading the trailing matrix ONCE per column (the gemv
// A[:, c+1:] v) and never writing it; the trailing updates  A[:, e:] -= Y V[e:,:]^H  and  A[k0+1:, e:] -= V (V T)^H A
// are done afterwards by the batched DMMA GEMM.  Panel columns are finished in-panel:
//   b = A[:,c] - Y V[c,:]^H ;  b -= V T^H (V^H b).
// Outputs (per member, leading dimension ld, HB_NB columns, zero-filled where unused):
//   Vp explicit V (unit diagonal, zeros above), Yp, VTp = V*T, Tws[panel] = T; v is also kept below the subdiagonal of H.
// dynamic smem: (2*ld + HB_NB*HB_NB + 4*HB_NB) * 16 + 512
// ---------------------------------------------------------------------------------------------
#define HB_NB 32
// MINB = 2: two CTAs per SM for multi-wave batches (see bidiag_panel_kernel)
template <int MINB>
__global__ void __launch_bounds__(E_THREADS, MINB) hess_panel_kernel(cplx* H, long long stride, int ld, const int* lv, int k0,
                                                                  cplx* Vp, cplx* Yp, cplx* VTp, long long pstride,
                                                                  cplx* Tws, long long tstride, cplx* ypart, int csize) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* bvec = reinterpret_cast<cplx*>(smem_raw);       // ld
    cplx* vvec = bvec + ld;                                // ld
    cplx* Tsm = vvec + ld;                                 // HB_NB x HB_NB (col-major, ld HB_NB)
    cplx* w1 = Tsm + HB_NB * HB_NB;                        // HB_NB
    cplx* w2 = w1 + HB_NB;                                 // HB_NB
    cplx* zv = w2 + HB_NB;                                 // HB_NB
    cplx* vrow = zv + HB_NB;                               // HB_NB : conj(V[c, :j])
    double* red = reinterpret_cast<double*>(vrow + HB_NB);
    // csize > 1: cluster of CTAs per member -- the gemv over the trailing matrix is split by COLUMNS (every thread keeps its row),
    // partial sums go through ypart[member][parity][rank][ld] and one cluster barrier per column
    const int b = blockIdx.x / csize, crank = blockIdx.x % csize, n = lv[b];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    cplx* Hb = H + (long long)b * stride;
    cplx* Vb = Vp + (long long)b * pstride;
    cplx* Yb = Yp + (long long)b * pstride;
    cplx* VTb = VTp + (long long)b * pstride;
    cplx* ypb = (csize > 1) ? ypart + (long long)b * 2 * csize * ld : nullptr;
    if (k0 + 2 >= n) {                                     // nothing left to reduce for this member: neutral panel
        for (int idx = tid; idx < ld * HB_NB; idx += E_THREADS) { Vb[idx] = mkc(0.0, 0.0); Yb[idx] = mkc(0.0, 0.0); VTb[idx] = mkc(0.0, 0.0); }
        return;
    }
    cplx* Tb = Tws + (long long)b * tstride + (long long)(k0 / HB_NB) * HB_NB * HB_NB;
    // Rows 0..k0 of Y and of the panel's columns take no part in the panel factorisation (LAPACK zlahr2): the kernel works on rows
    // >= r0 only, which cuts the level-2 traffic from n(n-c) to (n-k0)(n-c) per column; the caller fills Y[0:r0, :] = A[0:r0, r0:] (V T)
    // and updates A[0:r0, r0:e] with two DMMA GEMMs afterwards.
    const int r0 = k0 + 1;

    for (int idx = tid; idx < HB_NB * HB_NB; idx += E_THREADS) Tsm[idx] = mkc(0.0, 0.0);
    for (int idx = tid; idx < ld * HB_NB; idx += E_THREADS) { Vb[idx] = mkc(0.0, 0.0); Yb[idx] = mkc(0.0, 0.0); }
    __syncthreads();

    for (int j = 0; j < HB_NB; ++j) {
        const int c = k0 + j;
        if (c >= n) break;
        cplx* colc = Hb + (long long)ld * c;
        // ---- b = A[:,c] - Y[:, :j] conj(V[c, :j]) ----
        if (tid < j) vrow[tid] = cconj(Vb[c + (long long)ld * tid]);
        __syncthreads();
        for (int i = r0 + tid; i < n; i += E_THREADS) {
            cplx acc = colc[i];
            for (int jj = 0; jj < j; ++jj) acc = csub(acc, cmul(Yb[i + (long long)ld * jj], vrow[jj]));
            bvec[i] = acc;
        }
        __syncthreads();
        // ---- b -= V (T^H (V^H b)) ----
        if (j > 0) {
            for (int jj = warp; jj < j; jj += E_NWARPS) {
                const cplx* vc = Vb + (long long)ld * jj;
                cplx d = mkc(0.0, 0.0);
                for (int i = k0 + jj + 1 + lane; i < n; i += 32) d = cfmac(vc[i], bvec[i], d);
                d = warp_sum(d);
                if (lane == 0) w1[jj] = d;
            }
            __syncthreads();
            if (tid < j) {        // w2 = T^H w1 : w2[r] = sum_{q<=r} conj(T[q,r]) w1[q]
                cplx a = mkc(0.0, 0.0);
                for (int q = 0; q <= tid; ++q) a = cfmac(Tsm[q + HB_NB * tid], w1[q], a);
                w2[tid] = a;
            }
            __syncthreads();
            for (int i = k0 + 1 + tid; i < n; i += E_THREADS) {
                cplx acc = bvec[i];
                for (int jj = 0; jj < j; ++jj) acc = csub(acc, cmul(Vb[i + (long long)ld * jj], w2[jj]));
                bvec[i] = acc;
            }
            __syncthreads();
        }
        if (c + 2 >= n) {
            // last two columns of the matrix: no reflector, just store the updated column
            cluster_barrier(csize);             // every CTA of the cluster has read the old column before anyone overwrites it
            for (int i = r0 + tid; i < n; i += E_THREADS) colc[i] = bvec[i];
            __syncthreads();
            continue;
        }
        // ---- reflector from b[c+1:] ----
        double part = 0.0;
        for (int i = c + 2 + tid; i < n; i += E_THREADS) part += cabs2(bvec[i]);
        const double xnorm2 = block_sum(part, red);
        const cplx alpha = bvec[c + 1];
        cplx tau = mkc(0.0, 0.0);
        double beta = alpha.x;
        cplx scale = mkc(0.0, 0.0);
        const bool trivial = (xnorm2 == 0.0 && alpha.y == 0.0);
        if (!trivial) {
            beta = -copysign(sqrt(cabs2(alpha) + xnorm2), alpha.x);
            tau = mkc((beta - alpha.x) / beta, -alpha.y / beta);
            scale = cdiv(mkc(1.0, 0.0), mkc(alpha.x - beta, alpha.y));
        }
        __syncthreads();
        for (int i = r0 + tid; i < n; i += E_THREADS) {
            cplx vv = mkc(0.0, 0.0), hv = bvec[i];
            if (i == c + 1) { vv = mkc(1.0, 0.0); hv = trivial ? alpha : mkc(beta, 0.0); }
            else if (i > c + 1) { vv = trivial ? mkc(0.0, 0.0) : cmul(bvec[i], scale); hv = vv; }
            vvec[i] = vv;
            Vb[i + (long long)ld * j] = vv;
            if (csize == 1) colc[i] = hv;       // Hessenberg column above/at the subdiagonal, v stored below it
        }                                       // (cluster: written after the barrier below -- a slower CTA may still be reading the old column)
        __syncthreads();
        // ---- z = V[:, :j]^H v ;  T[:j, j] = -tau T[:j,:j] z ; T[j,j] = tau ----
        for (int jj = warp; jj < j; jj += E_NWARPS) {
            const cplx* vc = Vb + (long long)ld * jj;
            cplx d = mkc(0.0, 0.0);
            for (int i = c + 1 + lane; i < n; i += 32) d = cfmac(vc[i], vvec[i], d);
            d = warp_sum(d);
            if (lane == 0) zv[jj] = d;
        }
        __syncthreads();
        if (tid < j) {
            cplx a = mkc(0.0, 0.0);
            for (int q = tid; q < j; ++q) a = cfma(Tsm[tid + HB_NB * q], zv[q], a);
            Tsm[tid + HB_NB * j] = cneg(cmul(tau, a));
        }
        if (tid == 0) Tsm[j + HB_NB * j] = tau;
        // ---- y = tau (A[:, c+1:] v[c+1:] - Y[:, :j] z)  (the one pass over the trailing matrix) ----
        {
            const int len = n - c - 1;
            const int q0 = (int)(((long long)len * crank) / csize), q1 = (int)(((long long)len * (crank + 1)) / csize);
            cplx* mine = (csize > 1) ? ypb + ((long long)(j & 1) * csize + crank) * ld : nullptr;
            for (int i = r0 + tid; i < n; i += E_THREADS) {
                const cplx* row = Hb + i + (long long)ld * (c + 1);
                const cplx* vv = vvec + (c + 1);
                cplx y0 = mkc(0.0, 0.0), y1 = mkc(0.0, 0.0), y2 = mkc(0.0, 0.0), y3 = mkc(0.0, 0.0);
                int q = q0;
                for (; q + 7 < q1; q += 8) {
                    cplx a0 = row[(long long)ld * q], a1 = row[(long long)ld * (q + 1)], a2 = row[(long long)ld * (q + 2)], a3 = row[(long long)ld * (q + 3)];
                    cplx a4 = row[(long long)ld * (q + 4)], a5 = row[(long long)ld * (q + 5)], a6 = row[(long long)ld * (q + 6)], a7 = row[(long long)ld * (q + 7)];
                    y0 = cfma(a0, vv[q], y0); y1 = cfma(a1, vv[q + 1], y1); y2 = cfma(a2, vv[q + 2], y2); y3 = cfma(a3, vv[q + 3], y3);
                    y0 = cfma(a4, vv[q + 4], y0); y1 = cfma(a5, vv[q + 5], y1); y2 = cfma(a6, vv[q + 6], y2); y3 = cfma(a7, vv[q + 7], y3);
                }
                for (; q < q1; ++q) y0 = cfma(row[(long long)ld * q], vv[q], y0);
                cplx y = cadd(cadd(y0, y1), cadd(y2, y3));
                if (csize > 1) { mine[i] = y; continue; }
                for (int jj = 0; jj < j; ++jj) y = csub(y, cmul(Yb[i + (long long)ld * jj], zv[jj]));
                Yb[i + (long long)ld * j] = cmul(tau, y);
            }
            if (csize > 1) {
                cluster_barrier(csize);
                const cplx* parts = ypb + (long long)(j & 1) * csize * ld;
                for (int i = r0 + tid; i < n; i += E_THREADS) {
                    cplx y = parts[i];
                    for (int r = 1; r < csize; ++r) y = cadd(y, parts[(long long)r * ld + i]);
                    for (int jj = 0; jj < j; ++jj) y = csub(y, cmul(Yb[i + (long long)ld * jj], zv[jj]));
                    Yb[i + (long long)ld * j] = cmul(tau, y);      // every CTA writes the same value
                    cplx hv = bvec[i];
                    if (i == c + 1) hv = trivial ? alpha : mkc(beta, 0.0);
                    else if (i > c + 1) hv = vvec[i];
                    colc[i] = hv;
                }
            }
        }
        __syncthreads();
    }
    // ---- VT = V * T (rows k0+1..n-1), T to global ----
    for (int idx = tid; idx < HB_NB * HB_NB; idx += E_THREADS) Tb[idx] = Tsm[idx];
    __syncthreads();
    for (int i = tid; i < ld; i += E_THREADS) {
        const bool live = (i > k0 && i < n);
        for (int jj = 0; jj < HB_NB; ++jj) {
            cplx a = mkc(0.0, 0.0);
            if (live)
                for (int q = 0; q <= jj; ++q) a = cfma(Vb[i + (long long)ld * q], Tsm[q + HB_NB * jj], a);
            VTb[i + (long long)ld * jj] = a;
        }
    }
}

// Q formation, one panel (processed in reverse order): rebuilds the explicit V of the panel from the reflectors stored
// below the subdiagonal of H and writes VTp = V * T^H, so that  Q[k0+1:, k0+1:] -= V ((V T^H)^H Q[k0+1:, k0+1:]).
__global__ void __launch_bounds__(E_THREADS, 1) hess_qpanel_kernel(const cplx* H, long long stride, int ld, const int* lv, int k0,
                                                                   cplx* Vp, cplx* VTp, long long pstride, const cplx* Tws, long long tstride) {
    __shared__ cplx Tsm[HB_NB * HB_NB];
    const int b = blockIdx.x, n = lv[b];
    const int tid = threadIdx.x;
    const cplx* Hb = H + (long long)b * stride;
    cplx* Vb = Vp + (long long)b * pstride;
    cplx* VTb = VTp + (long long)b * pstride;
    if (k0 + 2 >= n) {
        for (int idx = tid; idx < ld * HB_NB; idx += E_THREADS) { Vb[idx] = mkc(0.0, 0.0); VTb[idx] = mkc(0.0, 0.0); }
        return;
    }
    const cplx* Tb = Tws + (long long)b * tstride + (long long)(k0 / HB_NB) * HB_NB * HB_NB;
    for (int idx = tid; idx < HB_NB * HB_NB; idx += E_THREADS) Tsm[idx] = Tb[idx];
    __syncthreads();
    for (int i = tid; i < ld; i += E_THREADS) {
        for (int q = 0; q < HB_NB; ++q) {
            const int c = k0 + q;
            cplx vv = mkc(0.0, 0.0);
            if (c + 2 < n && i < n) {
                if (i == c + 1) vv = mkc(1.0, 0.0);
                else if (i > c + 1) vv = Hb[i + (long long)ld * c];
            }
            Vb[i + (long long)ld * q] = vv;
        }
        for (int jj = 0; jj < HB_NB; ++jj) {          // (V T^H)[i,jj] = sum_{q>=jj} V[i,q] conj(T[jj,q])
            cplx a = mkc(0.0, 0.0);
            for (int q = jj; q < HB_NB; ++q) a = cfma(Vb[i + (long long)ld * q], cconj(Tsm[jj + HB_NB * q]), a);
            VTb[i + (long long)ld * jj] = a;
        }
    }
}

__global__ void set_identity_kernel(cplx* Q, long long stride, int ld, const int* lv) {
    const int b = blockIdx.y, n = lv[b];
    cplx* Qb = Q + (long long)b * stride;
    const long long total = (long long)ld * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int i = (int)(idx % ld), j = (int)(idx / ld);
        Qb[idx] = mkc(i == j ? 1.0 : 0.0, 0.0);
    }
}

__global__ void clear_below_subdiag_kernel(cplx* H, long long stride, int ld, const int* lv) {
    const int b = blockIdx.y, n = lv[b];
    cplx* Hb = H + (long long)b * stride;
    const long long total = (long long)ld * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int i = (int)(idx % ld), j = (int)(idx / ld);
        if (i > j + 1 && i < n) Hb[idx] = mkc(0.0, 0.0);
    }
}

// ---------------------------------------------------------------------------------------------
// small dense Hessenberg QR in shared memory, executed by ONE warp (all lanes run the same control flow)
// Hs: n x n upper Hessenberg (ld ldh) -> upper triangular; W (optional, wrows x n, ld ldw) <- W * (rotations)
// returns number of QR sweeps, or -1 on non-convergence
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool negligible_sub(cplx sub, cplx d0, cplx d1) {
    const double h = cabs1(sub);
    if (h <= LLCK_SAFMIN / LLCK_EPS) return true;
    const double tst = cabs1(d0) + cabs1(d1);
    return h <= LLCK_EPS * tst;
}

__device__ __noinline__ int warp_small_hqr(cplx* Hs, int ldh, cplx* W, int ldw, int n, int wrows) {
    const int lane = threadIdx.x & 31;
    int ihi = n - 1, its = 0, total = 0;
    while (ihi >= 0) {
        // largest k in [1, ihi] with a negligible subdiagonal entry (lane-parallel scan, 32 candidates per ballot)
        int ilo = 0;
        for (int base = ihi; base >= 1 && ilo == 0; base -= 32) {
            const int k = base - lane;
            const bool neg = (k >= 1) && negligible_sub(Hs[k + ldh * (k - 1)], Hs[(k - 1) + ldh * (k - 1)], Hs[k + ldh * k]);
            const unsigned bal = __ballot_sync(0xffffffffu, neg);
            if (bal) ilo = base - (__ffs(bal) - 1);
        }
        if (ilo > 0) {
            if (lane == 0) Hs[ilo + ldh * (ilo - 1)] = mkc(0.0, 0.0);
            __syncwarp();
        }
        if (ilo == ihi) { --ihi; its = 0; continue; }
        ++its; ++total;
        if (its > 300) return -1;
        cplx sh;
        if (its % 10 == 0) {
            cplx d = Hs[ihi + ldh * ihi];
            sh = mkc(d.x + 0.75 * fabs(Hs[ihi + ldh * (ihi - 1)].x), d.y);
        } else {
            cplx a = Hs[(ihi - 1) + ldh * (ihi - 1)], bb = Hs[(ihi - 1) + ldh * ihi];
            cplx cc = Hs[ihi + ldh * (ihi - 1)], d = Hs[ihi + ldh * ihi];
            cplx tr2 = cscale(cadd(a, d), 0.5);
            cplx det = csub(cmul(a, d), cmul(bb, cc));
            cplx disc = csqrt_(csub(cmul(tr2, tr2), det));
            cplx e1 = cadd(tr2, disc), e2 = csub(tr2, disc);
            sh = (cabs2(csub(e1, d)) < cabs2(csub(e2, d))) ? e1 : e2;
        }
        for (int k = ilo; k < ihi; ++k) {
            cplx x, y;
            if (k == ilo) { x = csub(Hs[ilo + ldh * ilo], sh); y = Hs[(ilo + 1) + ldh * ilo]; }
            else { x = Hs[k + ldh * (k - 1)]; y = Hs[(k + 1) + ldh * (k - 1)]; }
            double c; cplx s;
            givens(x, y, c, s);
            const cplx cs = cconj(s);
            // the accumulated transform only depends on (c, s) and on this lane's own rows: issue it first so that its
            // shared-memory latency overlaps the dependent row / column rotations of the window
            if (W != nullptr) {
                for (int row = lane; row < wrows; row += 32) {
                    cplx a = W[row + ldw * k], bq = W[row + ldw * (k + 1)];
                    W[row + ldw * k] = cadd(cscale(a, c), cmul(cs, bq));
                    W[row + ldw * (k + 1)] = csub(cscale(bq, c), cmul(s, a));
                }
            }
            __syncwarp();
            const int c0 = (k > ilo) ? k - 1 : k;
            for (int col = c0 + lane; col < n; col += 32) {
                cplx a = Hs[k + ldh * col], bq = Hs[(k + 1) + ldh * col];
                Hs[k + ldh * col] = cadd(cscale(a, c), cmul(s, bq));
                Hs[(k + 1) + ldh * col] = csub(cscale(bq, c), cmul(cs, a));
            }
            __syncwarp();
            if (k > ilo && lane == 0) Hs[(k + 1) + ldh * (k - 1)] = mkc(0.0, 0.0);
            const int r1 = min(k + 2, ihi);
            for (int row = lane; row <= r1; row += 32) {
                cplx a = Hs[row + ldh * k], bq = Hs[row + ldh * (k + 1)];
                Hs[row + ldh * k] = cadd(cscale(a, c), cmul(cs, bq));
                Hs[row + ldh * (k + 1)] = csub(cscale(bq, c), cmul(s, a));
            }
            __syncwarp();
        }
    }
    return total;
}

// ---------------------------------------------------------------------------------------------
// Aggressive early deflation (Braman-Byers-Mathias), executed by ONE warp on the trailing nw x nw window held in shared
// memory: T (Hessenberg on entry), V = I on entry.  s = H[kwtop, kwtop-1] is the spike scale.
//   1. Schur-decompose the window (T <- V^H T V upper triangular).
//   2. From the bottom: eigenvalue k is deflatable if |s V[0,k]| <= eps |T[k,k]|; undeflatable ones are swapped to the top.
//   3. The ns undeflated eigenvalues become the shifts of the next multishift sweep.
//   4. If anything deflated (or s == 0): reflect the spike s*conj(V[0,0:ns]) back to a single entry, restore the leading
//      ns x ns block to Hessenberg form (transformations accumulated in V); the caller writes T back and applies V.
// Returns ns (>= 0), or -1 if the window QR failed.  out[0] = 1 if T/V must be written back.  newsub = new H[kwtop,kwtop-1].
// ---------------------------------------------------------------------------------------------
// AED window: chosen by the host driver (24 / 28 / 32 by size; llck_options.aed_window overrides), at most 48 (vbuf / shifts hold 64)
__device__ __noinline__ int warp_aed(cplx* T, cplx* V, int nw, cplx s, cplx* shifts, cplx* vbuf, int* out, cplx* newsub, long long* ap) {
    const int lane = threadIdx.x & 31;
    const int L = E_LDH, LV = E_LDW;       // T lives in Hw (ld E_LDH), V in Ww (ld E_LDW)
    long long t0 = ap ? clock64() : 0;
    if (warp_small_hqr(T, L, V, LV, nw, nw) < 0) return -1;
    __syncwarp();
    if (ap) { long long t1 = clock64(); ap[0] += t1 - t0; t0 = t1; }
    int ns = nw, ilst = 0;
    const double s1 = cabs1(s);
    const double smallnum = LLCK_SAFMIN / LLCK_EPS;
    for (int knt = 0; knt < nw; ++knt) {
        if (ilst >= ns) break;
        double foo = cabs1(T[(ns - 1) + L * (ns - 1)]);
        if (foo == 0.0) foo = s1;
        if (s1 * cabs1(V[0 + LV * (ns - 1)]) <= fmax(smallnum, LLCK_EPS * foo)) {
            --ns;
        } else {
            // move T[ns-1,ns-1] up to position ilst by adjacent swaps
            for (int k = ns - 2; k >= ilst; --k) {
                const cplx t11 = T[k + L * k], t22 = T[(k + 1) + L * (k + 1)];
                double c; cplx sn;
                givens(T[k + L * (k + 1)], csub(t22, t11), c, sn);
                const cplx csn = cconj(sn);
                __syncwarp();
                for (int col = k + 2 + lane; col < nw; col += 32) {
                    cplx a = T[k + L * col], bq = T[(k + 1) + L * col];
                    T[k + L * col] = cadd(cscale(a, c), cmul(sn, bq));
                    T[(k + 1) + L * col] = csub(cscale(bq, c), cmul(csn, a));
                }
                for (int row = lane; row < k; row += 32) {
                    cplx a = T[row + L * k], bq = T[row + L * (k + 1)];
                    T[row + L * k] = cadd(cscale(a, c), cmul(csn, bq));
                    T[row + L * (k + 1)] = csub(cscale(bq, c), cmul(sn, a));
                }
                for (int row = lane; row < nw; row += 32) {
                    cplx a = V[row + LV * k], bq = V[row + LV * (k + 1)];
                    V[row + LV * k] = cadd(cscale(a, c), cmul(csn, bq));
                    V[row + LV * (k + 1)] = csub(cscale(bq, c), cmul(sn, a));
                }
                __syncwarp();
                if (lane == 0) { T[k + L * k] = t22; T[(k + 1) + L * (k + 1)] = t11; }
                __syncwarp();
            }
            ++ilst;
        }
    }
    if (ap) { long long t1 = clock64(); ap[1] += t1 - t0; t0 = t1; }
    if (ns == 0) s = mkc(0.0, 0.0);
    for (int j = lane; j < ns; j += 32) shifts[j] = T[j + L * j];
    const bool sz = (s.x == 0.0 && s.y == 0.0);
    const int nd = nw - ns;
    if (lane == 0) out[0] = (nd > 0 || sz) ? 1 : 0;
    __syncwarp();
    if (!(nd > 0 || sz)) return ns;
    if (ns > 1 && !sz) {
        // ---- reflect the spike w = s * conj(V[0, 0:ns]) to beta * e1 ----
        double part = 0.0;
        for (int j = lane; j < ns; j += 32) {
            cplx w = cmul(s, cconj(V[0 + LV * j]));
            vbuf[j] = w;
            if (j > 0) part += cabs2(w);
        }
        const double xnorm2 = warp_sum(part);
        __syncwarp();
        const cplx alpha = vbuf[0];
        if (!(xnorm2 == 0.0 && alpha.y == 0.0)) {
            const double beta = -copysign(sqrt(cabs2(alpha) + xnorm2), alpha.x);
            const cplx tau = mkc((beta - alpha.x) / beta, -alpha.y / beta), ctau = cconj(tau);
            const cplx scale = cdiv(mkc(1.0, 0.0), mkc(alpha.x - beta, alpha.y));
            __syncwarp();
            for (int j = lane; j < ns; j += 32) vbuf[j] = (j == 0) ? mkc(1.0, 0.0) : cmul(vbuf[j], scale);
            __syncwarp();
            for (int col = lane; col < nw; col += 32) {           // T[0:ns, :] <- P^H T
                cplx d = mkc(0.0, 0.0);
                for (int j = 0; j < ns; ++j) d = cfmac(vbuf[j], T[j + L * col], d);
                const cplx f = cmul(ctau, d);
                for (int j = 0; j < ns; ++j) T[j + L * col] = csub(T[j + L * col], cmul(f, vbuf[j]));
            }
            __syncwarp();
            for (int row = lane; row < ns; row += 32) {           // T[0:ns, 0:ns] <- T P
                cplx y = mkc(0.0, 0.0);
                for (int j = 0; j < ns; ++j) y = cfma(T[row + L * j], vbuf[j], y);
                const cplx f = cmul(tau, y);
                for (int j = 0; j < ns; ++j) T[row + L * j] = csub(T[row + L * j], cmul(f, cconj(vbuf[j])));
            }
            for (int row = lane; row < nw; row += 32) {           // V[:, 0:ns] <- V P
                cplx y = mkc(0.0, 0.0);
                for (int j = 0; j < ns; ++j) y = cfma(V[row + LV * j], vbuf[j], y);
                const cplx f = cmul(tau, y);
                for (int j = 0; j < ns; ++j) V[row + LV * j] = csub(V[row + LV * j], cmul(f, cconj(vbuf[j])));
            }
            __syncwarp();
        }
        // ---- restore Hessenberg form of T[0:ns, 0:ns] (Householder), carrying T[0:ns, ns:nw] and V[:, 0:ns] along ----
        for (int k = 0; k + 2 < ns; ++k) {
            const int len = ns - k - 1;
            double pp = 0.0;
            for (int i = 1 + lane; i < len; i += 32) pp += cabs2(T[(k + 1 + i) + L * k]);
            const double xn2 = warp_sum(pp);
            const cplx al = T[(k + 1) + L * k];
            if (xn2 == 0.0 && al.y == 0.0) continue;
            const double beta = -copysign(sqrt(cabs2(al) + xn2), al.x);
            const cplx tau = mkc((beta - al.x) / beta, -al.y / beta), ctau = cconj(tau);
            const cplx scale = cdiv(mkc(1.0, 0.0), mkc(al.x - beta, al.y));
            __syncwarp();
            for (int i = lane; i < len; i += 32) {
                if (i == 0) { vbuf[0] = mkc(1.0, 0.0); T[(k + 1) + L * k] = mkc(beta, 0.0); }
                else { vbuf[i] = cmul(T[(k + 1 + i) + L * k], scale); T[(k + 1 + i) + L * k] = mkc(0.0, 0.0); }
            }
            __syncwarp();
            for (int col = k + 1 + lane; col < nw; col += 32) {   // left: rows k+1..ns-1
                cplx d = mkc(0.0, 0.0);
                for (int i = 0; i < len; ++i) d = cfmac(vbuf[i], T[(k + 1 + i) + L * col], d);
                const cplx f = cmul(ctau, d);
                for (int i = 0; i < len; ++i) T[(k + 1 + i) + L * col] = csub(T[(k + 1 + i) + L * col], cmul(f, vbuf[i]));
            }
            __syncwarp();
            for (int row = lane; row < ns; row += 32) {           // right: cols k+1..ns-1 of T[0:ns, :]
                cplx y = mkc(0.0, 0.0);
                for (int i = 0; i < len; ++i) y = cfma(T[row + L * (k + 1 + i)], vbuf[i], y);
                const cplx f = cmul(tau, y);
                for (int i = 0; i < len; ++i) T[row + L * (k + 1 + i)] = csub(T[row + L * (k + 1 + i)], cmul(f, cconj(vbuf[i])));
            }
            for (int row = lane; row < nw; row += 32) {           // right on V
                cplx y = mkc(0.0, 0.0);
                for (int i = 0; i < len; ++i) y = cfma(V[row + LV * (k + 1 + i)], vbuf[i], y);
                const cplx f = cmul(tau, y);
                for (int i = 0; i < len; ++i) V[row + LV * (k + 1 + i)] = csub(V[row + LV * (k + 1 + i)], cmul(f, cconj(vbuf[i])));
            }
            __syncwarp();
        }
    }
    if (lane == 0) *newsub = cmul(s, cconj(V[0]));
    __syncwarp();
    if (ap) ap[2] += clock64() - t0;
    return ns;
}

// ---------------------------------------------------------------------------------------------
// apply the window transform Ww (64x64 in smem, identity-padded beyond ww) to the off-window strips:
//   H[ws:we, we:n] <- Ww^H H[ws:we, we:n];  H[0:ws, ws:we] <- H[0:ws, ws:we] Ww;  Z[:, ws:we] <- Z[:, ws:we] Ww
// all E_THREADS threads participate; tiles: 2 x E_TILE double buffer
// ---------------------------------------------------------------------------------------------
// cluster-wide barrier with release/acquire ordering of the global-memory strip writes (no-op for a single CTA per member)
__device__ __forceinline__ void hqr_cluster_sync(int csize) {
    if (csize > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    } else {
        __syncthreads();
    }
}

// With csize > 1 CTAs per member (small batches), every CTA of the cluster runs the same window computation redundantly (identical
// Ww in its own shared memory) and applies it to every csize-th strip tile only; the call ends with a cluster barrier.
__device__ __noinline__ void apply_window_transform(cplx* Hb, cplx* Zb, int ld, int n, int ws, int we, const cplx* Ww, cplx* tiles, int* kr,
                                                    int crank, int csize) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int ww = we - ws;
    // W is a product of plane rotations: every column has a limited range of nonzero rows.  kr[2J], kr[2J+1] = first/last+1
    // row (multiples of 4) holding a nonzero in column tile J (8 columns); the DMMA k-loops below skip the exact zeros.
    __syncthreads();
    if (tid < 64) {
        int lo = 64, hi = 0;
        for (int k = 0; k < E_W; ++k) {
            cplx v = Ww[k + E_LDW * tid];
            if (v.x != 0.0 || v.y != 0.0) { if (k < lo) lo = k; hi = k + 1; }
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((tid & 7) == 0) {
            if (lo >= hi) { lo = 0; hi = 4; }
            kr[2 * (tid >> 3)] = lo & ~3;
            kr[2 * (tid >> 3) + 1] = (hi + 3) & ~3;
        }
    }
    __syncthreads();
    // ---- row strip ----
    {
        const int ncols = n - we;
        const int ntiles = (ncols + 31) / 32;
        auto load = [&](int buf, int tl) {
            const int c0 = we + tl * 32;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int idx = tid + E_THREADS * r;      // 64 x 32 = 2048
                int k = idx & 63, j = idx >> 6;
                bool ok = (k < ww) && (c0 + j < n);
                const cplx* src = ok ? (Hb + (ws + k) + (long long)ld * (c0 + j)) : Hb;
                cp_async16(&tiles[buf * E_TILE + k + 68 * j], src, ok);
            }
            cp_async_commit();
        };
        const int wr = warp >> 1, wc = warp & 1;
        if (crank < ntiles) load(0, crank);
        for (int tl = crank, it = 0; tl < ntiles; tl += csize, ++it) {
            const int buf = it & 1;
            if (tl + csize < ntiles) { load(buf ^ 1, tl + csize); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncthreads();
            const cplx* T = tiles + buf * E_TILE;
            double acc[1][2][4];
            zero_acc<1, 2>(acc);
            {
                const int klo = kr[2 * wr], khi = kr[2 * wr + 1];
                warp_zmma<1, 2, true, false>(acc, Ww + E_LDW * (8 * wr) + klo, E_LDW, 1, T + 68 * (16 * wc) + klo, 1, 68, khi - klo);
            }
            const int row = 8 * wr + g;
            if (row < ww) {
                const int c0 = we + tl * 32 + 16 * wc;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    int c = c0 + 8 * j + 2 * t;
                    if (c < n) Hb[(ws + row) + (long long)ld * c] = mkc(acc[0][j][0], acc[0][j][2]);
                    if (c + 1 < n) Hb[(ws + row) + (long long)ld * (c + 1)] = mkc(acc[0][j][1], acc[0][j][3]);
                }
            }
            __syncthreads();
        }
    }
    // ---- column strips: H rows [0, ws) then Z rows [0, n) ----
    for (int which = 0; which < 2; ++which) {
        cplx* Mb = which ? Zb : Hb;
        const int nrows = which ? n : ws;
        const int ntiles = (nrows + 31) / 32;
        auto load = [&](int buf, int tl) {
            const int r0 = tl * 32;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int idx = tid + E_THREADS * r;      // 32 x 64
                int i = idx & 31, k = idx >> 5;
                bool ok = (r0 + i < nrows) && (k < ww);
                const cplx* src = ok ? (Mb + (r0 + i) + (long long)ld * (ws + k)) : Mb;
                cp_async16(&tiles[buf * E_TILE + i + 34 * k], src, ok);
            }
            cp_async_commit();
        };
        // warps w, w+4, w+8, w+12 share one SM sub-partition (and its DMMA pipe): give each sub-partition all four column
        // groups (their k-ranges differ) so that the four pipes carry equal work
        const int wr = warp >> 2, wc = (warp + (warp >> 2)) & 3;
        if (crank < ntiles) load(0, crank);
        for (int tl = crank, it = 0; tl < ntiles; tl += csize, ++it) {
            const int buf = it & 1;
            if (tl + csize < ntiles) { load(buf ^ 1, tl + csize); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncthreads();
            const cplx* T = tiles + buf * E_TILE;
            double acc[1][2][4];
            zero_acc<1, 2>(acc);
            {
                const int klo = min(kr[4 * wc], kr[4 * wc + 2]), khi = max(kr[4 * wc + 1], kr[4 * wc + 3]);
                warp_zmma<1, 2, false, false>(acc, T + 8 * wr + 34 * klo, 1, 34, Ww + E_LDW * (16 * wc) + klo, 1, E_LDW, khi - klo);
            }
            const int row = tl * 32 + 8 * wr + g;
            if (row < nrows) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    int c = 16 * wc + 8 * j + 2 * t;
                    if (c < ww) Mb[row + (long long)ld * (ws + c)] = mkc(acc[0][j][0], acc[0][j][2]);
                    if (c + 1 < ww) Mb[row + (long long)ld * (ws + c + 1)] = mkc(acc[0][j][1], acc[0][j][3]);
                }
            }
            __syncthreads();
        }
    }
    hqr_cluster_sync(csize);
}

__device__ __forceinline__ int block_max_int(int v, int* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = __reduce_max_sync(0xffffffffu, v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    int r = (lane < E_NWARPS) ? scratch[lane] : -2147483647;
    r = __reduce_max_sync(0xffffffffu, r);
    return r;
}

// status: 0 ok, 1 = QR did not converge
__global__ void __launch_bounds__(E_THREADS, 1) hqr_kernel(cplx* H, cplx* Z, long long stride, int ld, const int* lv, int* status, int* sweeps_out, long long* prof, int max_trains, int aed_nw, int nibble, int csize) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* Hw = reinterpret_cast<cplx*>(smem_raw);
    cplx* Ww = Hw + E_MAT;
    cplx* tiles = Ww + E_MAT;
    cplx* Hs = tiles + 2 * E_TILE;            // E_LDS x E_NB
    cplx* shifts = Hs + E_LDS * E_NB;         // E_NB (+ padding to 64)
    int* iscr = reinterpret_cast<int*>(shifts + 64);   // 64 ints
    // csize CTAs (one thread-block cluster) per member: all of them run the window computations redundantly and share the strips
    const int b = blockIdx.x / csize, crank = blockIdx.x % csize, n = lv[b];
    cplx* Hb = H + (long long)b * stride;
    cplx* Zb = Z + (long long)b * stride;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    int ihi = n - 1, its = 0, nsweeps = 0;
    bool failed = false;
    long long tp[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // scan+shifts, window load, chase, store, strips, small blocks, AED load, AED warp work, AED strips, AED calls
    __shared__ long long aedprof[3];
    if (threadIdx.x < 3) aedprof[threadIdx.x] = 0;
    long long tc = clock64();
#define PROF(i) do { if (prof) { long long tn_ = clock64(); tp[i] += tn_ - tc; tc = tn_; } } while (0)
    while (ihi >= 0) {
        __syncthreads();
        // ---- deflation scan: largest k in [1, ihi] with negligible H[k,k-1] ----
        int cand = 0;
        for (int k = ihi - tid; k >= 1; k -= E_THREADS) {
            if (negligible_sub(Hb[k + (long long)ld * (k - 1)], Hb[(k - 1) + (long long)ld * (k - 1)], Hb[k + (long long)ld * k])) { cand = k; break; }
        }
        const int ilo = block_max_int(cand, iscr);
        if (ilo > 0 && tid == 0) Hb[ilo + (long long)ld * (ilo - 1)] = mkc(0.0, 0.0);
        if (ilo == ihi) { --ihi; its = 0; continue; }
        const int size = ihi - ilo + 1;
        if (size <= E_W) {
            // ---- whole active block fits in the window: finish it in shared memory ----
            for (int idx = tid; idx < E_W * E_W; idx += E_THREADS) {
                int r = idx & 63, c = idx >> 6;
                cplx hv = mkc(0.0, 0.0);
                if (r < size && c < size) hv = Hb[(ilo + r) + (long long)ld * (ilo + c)];
                Hw[r + E_LDH * c] = hv;
                Ww[r + E_LDW * c] = mkc(r == c ? 1.0 : 0.0, 0.0);
            }
            __syncthreads();
            if (warp == 0) {
                int r = warp_small_hqr(Hw, E_LDH, Ww, E_LDW, size, size);
                if (lane == 0) iscr[32] = r;
            }
            __syncthreads();
            if (iscr[32] < 0) { failed = true; break; }
            PROF(5);
            // strips first (they never touch the window block), cluster barrier, then the window itself: a CTA of the cluster that is
            // still loading this window must not see another CTA's post-transform values
            apply_window_transform(Hb, Zb, ld, n, ilo, ihi + 1, Ww, tiles, iscr + 40, crank, csize);
            for (int idx = tid; idx < size * size; idx += E_THREADS) {
                int r = idx % size, c = idx / size;
                Hb[(ilo + r) + (long long)ld * (ilo + c)] = Hw[r + E_LDH * c];
            }
            PROF(4);
            ihi = ilo - 1; its = 0;
            continue;
        }
        ++its;
        if (its > 60) { failed = true; break; }
        // ---- aggressive early deflation on the trailing aed_nw x aed_nw window; its undeflated eigenvalues are the shifts ----
        int nbu = E_NB, ntrains = 1, ns_all = E_NB;
        {
            const int nw = aed_nw;                    // size > E_W >= nw
            const int kwtop = ihi - nw + 1;
            const cplx spike = Hb[kwtop + (long long)ld * (kwtop - 1)];
            for (int idx = tid; idx < E_W * E_W; idx += E_THREADS) {
                int r = idx & 63, c = idx >> 6;
                cplx hv = mkc(0.0, 0.0);
                if (r < nw && c < nw && r <= c + 1) hv = Hb[(kwtop + r) + (long long)ld * (kwtop + c)];
                Hw[r + E_LDH * c] = hv;
                Ww[r + E_LDW * c] = mkc(r == c ? 1.0 : 0.0, 0.0);
            }
            __syncthreads();
            PROF(6);
            if (warp == 0) {
                int r = warp_aed(Hw, Ww, nw, spike, shifts, Hs, iscr + 33, Hs + 64, prof ? aedprof : nullptr);
                if (lane == 0) iscr[32] = r;
            }
            __syncthreads();
            PROF(7);
            if (prof) ++tp[9];
            const int ns = iscr[32];
            if (ns < 0) { failed = true; break; }
            const int nd = nw - ns;
            if (iscr[33]) {
                apply_window_transform(Hb, Zb, ld, n, kwtop, ihi + 1, Ww, tiles, iscr + 40, crank, csize);
                for (int idx = tid; idx < nw * nw; idx += E_THREADS) {
                    int r = idx % nw, c = idx / nw;
                    Hb[(kwtop + r) + (long long)ld * (kwtop + c)] = (r <= c + 1) ? Hw[r + E_LDH * c] : mkc(0.0, 0.0);
                }
                if (tid == 0) Hb[kwtop + (long long)ld * (kwtop - 1)] = Hs[64];
            }
            PROF(8);
            if (nd > 0) { ihi -= nd; its = 0; }
            // good deflation: look again before sweeping.  A sweep costs O(n^2), the single-warp AED a constant, so the threshold
            // (LAPACK's NIBBLE) is high: round 1 (window 24) measured hqr ms at nibble 14/30/60: n=512 1024/972/918, n=256 635/585/504;
            // with the round-2 windows (28 / 32) every nibble in 45..100 gives the same time at n = 768 and 1024 and beats 30
            const int nib = nibble > 0 ? nibble : 60;
            if (nd * 100 > nib * nw || ihi - ilo + 1 <= E_W) continue;
            if (ns < 2 || (its > 0 && its % 6 == 0)) {
                __syncthreads();
                if (tid < E_NB) {
                    cplx d = Hb[(ihi - tid) + (long long)ld * (ihi - tid)];
                    cplx sub = Hb[(ihi - tid) + (long long)ld * (ihi - tid - 1)];
                    shifts[tid] = mkc(d.x + 0.75 * cabs_(sub), d.y);
                }
                __syncthreads();
            } else {
                // all ns undeflated window eigenvalues are used as shifts: trains of <= E_NB bulges, one sweep per train
                ns_all = ns;
                ntrains = min(max_trains, (ns + E_NB - 1) / E_NB);
                nbu = min(ns, E_NB);
            }
        }
        // ---- one multishift sweep over [ilo, ihi] per train of shifts ----
        PROF(0);
        for (int train = 0; train < ntrains; ++train) {
        const cplx* tshifts = shifts + train * E_NB;
        if (ntrains > 1) nbu = min(ns_all - train * E_NB, E_NB);
        ++nsweeps;
        int tstep = 0;
        while (true) {
            const int p_last = ilo - 1 - 2 * (nbu - 1) + tstep;
            if (p_last > ihi - 2) break;
            const int p_top = max(ilo - 1, p_last);
            const int p0 = ilo - 1 + tstep;
            const int ws = max(ilo, p_top);
            const int we = min(ws + E_W, ihi + 1);
            const int T = (we == ihi + 1) ? ((ihi - 2) - p_last + 1) : (we - 3 - p0);
            const int ww = we - ws;
            __syncthreads();
            for (int idx = tid; idx < E_W * E_W; idx += E_THREADS) {
                int r = idx & 63, c = idx >> 6;
                cplx hv = mkc(0.0, 0.0);
                if (r < ww && c < ww) hv = Hb[(ws + r) + (long long)ld * (ws + c)];
                Hw[r + E_LDH * c] = hv;
                Ww[r + E_LDW * c] = mkc(r == c ? 1.0 : 0.0, 0.0);
            }
            __syncthreads();
            PROF(1);
            for (int step = 0; step < T; ++step) {
                const int p = ilo - 1 - 2 * warp + tstep + step;     // this warp's bulge
                const bool active = (warp < nbu) && (p >= ilo - 1) && (p <= ihi - 2);
                double c = 1.0; cplx s = mkc(0.0, 0.0);
                if (active) {
                    cplx x, y;
                    if (p == ilo - 1) { x = csub(Hw[(ilo - ws) + E_LDH * (ilo - ws)], tshifts[warp]); y = Hw[(ilo + 1 - ws) + E_LDH * (ilo - ws)]; }
                    else { x = Hw[(p + 1 - ws) + E_LDH * (p - ws)]; y = Hw[(p + 2 - ws) + E_LDH * (p - ws)]; }
                    givens(x, y, c, s);
                    const cplx cs = cconj(s);
                    const int r = p + 1 - ws;
                    const int c0 = max(p - ws, 0);
                    __syncwarp();
                    for (int col = c0 + lane; col < ww; col += 32) {
                        cplx a = Hw[r + E_LDH * col], bq = Hw[(r + 1) + E_LDH * col];
                        Hw[r + E_LDH * col] = cadd(cscale(a, c), cmul(s, bq));
                        Hw[(r + 1) + E_LDH * col] = csub(cscale(bq, c), cmul(cs, a));
                    }
                    __syncwarp();
                    if (p >= ilo && lane == 0) Hw[(r + 1) + E_LDH * (p - ws)] = mkc(0.0, 0.0);
                }
                __syncthreads();
                if (active) {
                    const cplx cs = cconj(s);
                    const int k = p + 1 - ws;
                    const int r1 = min(p + 3, ihi) - ws + 1;
                    for (int row = lane; row < r1; row += 32) {
                        cplx a = Hw[row + E_LDH * k], bq = Hw[row + E_LDH * (k + 1)];
                        Hw[row + E_LDH * k] = cadd(cscale(a, c), cmul(cs, bq));
                        Hw[row + E_LDH * (k + 1)] = csub(cscale(bq, c), cmul(s, a));
                    }
                    for (int row = lane; row < ww; row += 32) {
                        cplx a = Ww[row + E_LDW * k], bq = Ww[row + E_LDW * (k + 1)];
                        Ww[row + E_LDW * k] = cadd(cscale(a, c), cmul(cs, bq));
                        Ww[row + E_LDW * (k + 1)] = csub(cscale(bq, c), cmul(s, a));
                    }
                }
                __syncthreads();
            }
            PROF(2);
            apply_window_transform(Hb, Zb, ld, n, ws, we, Ww, tiles, iscr + 40, crank, csize);
            PROF(4);
            for (int idx = tid; idx < ww * ww; idx += E_THREADS) {
                int r = idx % ww, c = idx / ww;
                Hb[(ws + r) + (long long)ld * (ws + c)] = Hw[r + E_LDH * c];
            }
            PROF(3);
            tstep += T;
        }
        }
    }
    if (tid == 0 && crank == 0) {
        if (failed) atomicMax(&status[b], 1);
        if (sweeps_out) sweeps_out[b] = nsweeps;
        if (prof) { for (int i = 0; i < 10; ++i) prof[10 * b + i] = tp[i]; if (lane == 0) { prof[10 * b + 0] = aedprof[0]; prof[10 * b + 6] = aedprof[1]; } }
    }
#undef PROF
}

// ---------------------------------------------------------------------------------------------
// Blocked trevc: X (upper triangular eigenvector matrix of T) by block back substitution.
// For the 32-row block [j0, j0+32): trevc_diag_kernel solves, for every column k >= j0, the small shifted triangular
// system inside the block (one thread per column, block of T in shared memory); the rows above the block are then
// updated for all columns at once by the batched DMMA GEMM   X[0:j0, j0:n] -= T[0:j0, j0:j1] * X[j0:j1, j0:n].
// X must be zero-initialised (trevc_zero_kernel); the unit diagonal is written by the diag kernel.
// ---------------------------------------------------------------------------------------------
#define TV_NB 32
__global__ void trevc_zero_kernel(cplx* X, long long stride, int ld, const int* lv) {
    const int b = blockIdx.y, n = lv[b];
    cplx* Xb = X + (long long)b * stride;
    const long long total = (long long)ld * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
        Xb[idx] = mkc(0.0, 0.0);
}

__global__ void __launch_bounds__(128) trevc_diag_kernel(const cplx* Tm, cplx* X, long long stride, int ld, const int* lv, int j0) {
    __shared__ cplx Tblk[TV_NB][TV_NB + 1];
    const int b = blockIdx.y, n = lv[b];
    if (j0 >= n) return;
    const cplx* Tb = Tm + (long long)b * stride;
    cplx* Xb = X + (long long)b * stride;
    const int j1 = min(j0 + TV_NB, n), nr = j1 - j0;
    for (int idx = threadIdx.x; idx < TV_NB * TV_NB; idx += blockDim.x) {
        int r = idx % TV_NB, c = idx / TV_NB;
        Tblk[r][c] = (r < nr && c < nr && r <= c) ? Tb[(j0 + r) + (long long)ld * (j0 + c)] : mkc(0.0, 0.0);
    }
    __syncthreads();
    const int k = j0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const cplx tkk = Tb[k + (long long)ld * k];
    const double smlnum = LLCK_SAFMIN * ((double)n / LLCK_EPS);
    const double smin = fmax(LLCK_EPS * cabs1(tkk), smlnum);
    cplx* xc = Xb + (long long)ld * k + j0;
    cplx x[TV_NB];
#pragma unroll
    for (int r = 0; r < TV_NB; ++r) x[r] = (r < nr) ? xc[r] : mkc(0.0, 0.0);
    const int rtop = min(nr - 1, k - j0);      // rows below k (inside the block) stay zero
#pragma unroll
    for (int r = TV_NB - 1; r >= 0; --r) {
        if (r <= rtop) {
            cplx xr;
            if (j0 + r == k) xr = mkc(1.0, 0.0);
            else {
                cplx d = csub(Tblk[r][r], tkk);
                if (cabs1(d) < smin) d = mkc(smin, 0.0);
                xr = cdiv(x[r], d);
            }
            x[r] = xr;
#pragma unroll
            for (int q = 0; q < TV_NB; ++q)
                if (q < r) x[q] = csub(x[q], cmul(xr, Tblk[q][r]));
        }
    }
#pragma unroll
    for (int r = 0; r < TV_NB; ++r)
        if (r <= rtop) xc[r] = x[r];
}

__global__ void trevc_normalize_kernel(cplx* X, long long stride, int ld, const int* lv) {
    const int b = blockIdx.y, n = lv[b];
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= n) return;
    cplx* xc = X + (long long)b * stride + (long long)ld * k;
    double mx = 0.0;
    for (int i = lane; i <= k; i += 32) mx = fmax(mx, cabs1(xc[i]));
    mx = warp_max(mx);
    if (mx > 0.0 && isfinite(mx)) {
        const double sc = 1.0 / mx;
        for (int i = lane; i <= k; i += 32) xc[i] = cscale(xc[i], sc);
    }
}
