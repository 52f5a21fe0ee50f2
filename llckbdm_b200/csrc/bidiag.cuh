// Blocked Householder bidiagonalisation  A = Q B P^H  (B real upper bidiagonal) for the SVD front end -- first half of
// the replacement of scipy.linalg.svd (reference llckbdm/kbdm.py:166).  One CTA per ensemble member per panel.
//
// Inside a panel of BD_NB columns the trailing matrix is only READ: with V/U the left/right Householder vectors generated
// so far and Y, X the accumulated products, the current matrix is  A_i = A - V Y^H - X U^H.  Each column costs two
// passes over the trailing matrix (y = tau A_i^H v : column dots, x = pi A_i u : row dots); the rank-2*BD_NB trailing
// update  A[e:,e:] -= V Y^H + X U^H  is done afterwards by the batched DMMA GEMM.
// The reflectors stay in A (v below the diagonal, u right of the superdiagonal); T factors of the compact-WY forms
// Q_panel = I - V TQ V^H, P_panel = I - U TP U^H are kept per panel for the blocked formation of Q and P.
#pragma once
#include "common.cuh"
#include "eig.cuh"

#define BD_NB 32

// dynamic smem: (2*ld + 2*BD_NB*BD_NB + 8*BD_NB + 8) * 16 + 512
// MINB = 2: two CTAs (members) per SM for batches of more than one wave -- one member's serial vector phase overlaps the other's
// streaming gemv pass (64 registers per thread, a few spilled words); MINB = 1 for single-wave batches and cluster launches.
template <int MINB>
__global__ void __launch_bounds__(E_THREADS, MINB) bidiag_panel_kernel(cplx* A, long long stride, int ld, const int* mv, int k0,
                                                                    cplx* Vp, cplx* Yp, cplx* Xp, cplx* Up, long long pstride,
                                                                    cplx* TQws, cplx* TPws, long long tstride, double* dws, double* ews,
                                                                    cplx* ypart, int csize) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* vvec = reinterpret_cast<cplx*>(smem_raw);       // ld : column work vector, then v
    cplx* uvec = vvec + ld;                                // ld : y, then conj(row), then u
    cplx* TQ = uvec + ld;                                  // BD_NB x BD_NB
    cplx* TP = TQ + BD_NB * BD_NB;
    cplx* cy = TP + BD_NB * BD_NB;                         // conj(Y[c, :i])
    cplx* cu = cy + BD_NB;                                 // conj(U[c, :i])
    cplx* zv = cu + BD_NB;                                 // V^H v
    cplx* zx = zv + BD_NB;                                 // X^H v
    cplx* zy = zx + BD_NB;                                 // Y^H u
    cplx* zu = zy + BD_NB;                                 // U^H u
    cplx* vrow = zu + BD_NB;                               // V[c, :i+1]
    cplx* xrow = vrow + BD_NB;                             // X[c, :i]
    double* red = reinterpret_cast<double*>(xrow + BD_NB + 8);
    // csize > 1 (small batches): a thread-block cluster per member.  Every CTA runs the O(m) bookkeeping redundantly (bit-identical);
    // pass 1 is split by column pairs, pass 2 by column ranges with partial sums through ypart[member][parity][rank][ld]; the writes of
    // the reflectors into A are deferred behind the cluster barriers so that no CTA sees them while it still needs the old column / row.
    const int b = blockIdx.x / csize, crank = blockIdx.x % csize, m = mv[b];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    cplx* Ab = A + (long long)b * stride;
    cplx* Vb = Vp + (long long)b * pstride;
    cplx* Yb = Yp + (long long)b * pstride;
    cplx* Xb = Xp + (long long)b * pstride;
    cplx* Ub = Up + (long long)b * pstride;
    cplx* ypb = (csize > 1) ? ypart + (long long)b * 2 * csize * ld : nullptr;
    if (k0 >= m) {                                         // member already finished: neutral panel for the batched updates
        for (int idx = tid; idx < ld * BD_NB; idx += E_THREADS) {
            Vb[idx] = mkc(0.0, 0.0); Yb[idx] = mkc(0.0, 0.0); Xb[idx] = mkc(0.0, 0.0); Ub[idx] = mkc(0.0, 0.0);
        }
        return;
    }
    double* db = dws + (long long)b * ld;
    double* eb = ews + (long long)b * ld;
    const long long tpan = (long long)(k0 / BD_NB) * BD_NB * BD_NB;

    for (int idx = tid; idx < BD_NB * BD_NB; idx += E_THREADS) { TQ[idx] = mkc(0.0, 0.0); TP[idx] = mkc(0.0, 0.0); }
    for (int idx = tid; idx < ld * BD_NB; idx += E_THREADS) {
        Vb[idx] = mkc(0.0, 0.0); Yb[idx] = mkc(0.0, 0.0); Xb[idx] = mkc(0.0, 0.0); Ub[idx] = mkc(0.0, 0.0);
    }
    __syncthreads();

    for (int i = 0; i < BD_NB; ++i) {
        const int c = k0 + i;
        if (c >= m) break;
        // ---------- column c of A_i ----------
        if (tid < i) { cy[tid] = cconj(Yb[c + (long long)ld * tid]); cu[tid] = cconj(Ub[c + (long long)ld * tid]); }
        __syncthreads();
        for (int r = c + tid; r < m; r += E_THREADS) {
            cplx acc = Ab[r + (long long)ld * c];
            for (int jj = 0; jj < i; ++jj) {
                acc = csub(acc, cmul(Vb[r + (long long)ld * jj], cy[jj]));
                acc = csub(acc, cmul(Xb[r + (long long)ld * jj], cu[jj]));
            }
            vvec[r] = acc;
        }
        __syncthreads();
        // ---------- left reflector ----------
        double part = 0.0;
        for (int r = c + 1 + tid; r < m; r += E_THREADS) part += cabs2(vvec[r]);
        const double xnorm2 = block_sum(part, red);
        const cplx alpha = vvec[c];
        cplx tau = mkc(0.0, 0.0), scale = mkc(0.0, 0.0);
        double beta = alpha.x;
        const bool trivial = (xnorm2 == 0.0 && alpha.y == 0.0);
        if (!trivial) {
            beta = -copysign(sqrt(cabs2(alpha) + xnorm2), alpha.x);
            tau = mkc((beta - alpha.x) / beta, -alpha.y / beta);
            scale = cdiv(mkc(1.0, 0.0), mkc(alpha.x - beta, alpha.y));
        }
        __syncthreads();
        for (int r = c + tid; r < m; r += E_THREADS) {
            cplx vv;
            if (r == c) { vv = mkc(1.0, 0.0); if (csize == 1) Ab[c + (long long)ld * c] = mkc(beta, 0.0); }
            else { vv = trivial ? mkc(0.0, 0.0) : cmul(vvec[r], scale); if (csize == 1) Ab[r + (long long)ld * c] = vv; }
            vvec[r] = vv;
            Vb[r + (long long)ld * i] = vv;
        }
        if (tid == 0) db[c] = beta;
        __syncthreads();
        // ---------- zv = V^H v, zx = X^H v ; TQ column ----------
        for (int w = warp; w < 2 * i; w += E_NWARPS) {
            const int jj = w >> 1;
            const cplx* src = ((w & 1) ? Xb : Vb) + (long long)ld * jj;
            cplx d = mkc(0.0, 0.0);
            for (int r = c + lane; r < m; r += 32) d = cfmac(src[r], vvec[r], d);
            d = warp_sum(d);
            if (lane == 0) { if (w & 1) zx[jj] = d; else zv[jj] = d; }
        }
        __syncthreads();
        if (tid < i) {
            cplx a = mkc(0.0, 0.0);
            for (int q = tid; q < i; ++q) a = cfma(TQ[tid + BD_NB * q], zv[q], a);
            TQ[tid + BD_NB * i] = cneg(cmul(tau, a));
        }
        if (tid == 0) TQ[i + BD_NB * i] = tau;
        if (c + 1 >= m) {
            if (csize > 1) {                    // last column: store its reflector once every CTA of the cluster has read the old column
                cluster_barrier(csize);
                for (int r = c + tid; r < m; r += E_THREADS) Ab[r + (long long)ld * c] = (r == c) ? mkc(beta, 0.0) : vvec[r];
            }
            __syncthreads();
            continue;
        }
        // ---------- y = tau (A^H v - Y zv - U zx) on columns j > c   (pass 1 over the trailing matrix: column dots) ----------
        for (int j = c + 1 + 2 * (warp + E_NWARPS * crank); j < m; j += 2 * E_NWARPS * csize) {   // two columns per warp, 4 independent loads each
            const bool two = (j + 1 < m);
            const cplx* col0 = Ab + (long long)ld * j;
            const cplx* col1 = Ab + (long long)ld * (two ? j + 1 : j);
            cplx a0 = mkc(0.0, 0.0), a1 = a0, a2 = a0, a3 = a0, b0 = a0, b1 = a0, b2 = a0, b3 = a0;
            int r = c + lane;
            for (; r + 96 < m; r += 128) {
                const cplx v0 = vvec[r], v1 = vvec[r + 32], v2 = vvec[r + 64], v3 = vvec[r + 96];
                const cplx x0 = col0[r], x1 = col0[r + 32], x2 = col0[r + 64], x3 = col0[r + 96];
                const cplx y0 = col1[r], y1 = col1[r + 32], y2 = col1[r + 64], y3 = col1[r + 96];
                a0 = cfmac(x0, v0, a0); a1 = cfmac(x1, v1, a1); a2 = cfmac(x2, v2, a2); a3 = cfmac(x3, v3, a3);
                b0 = cfmac(y0, v0, b0); b1 = cfmac(y1, v1, b1); b2 = cfmac(y2, v2, b2); b3 = cfmac(y3, v3, b3);
            }
            for (; r < m; r += 32) { const cplx v0 = vvec[r]; a0 = cfmac(col0[r], v0, a0); b0 = cfmac(col1[r], v0, b0); }
            cplx d0 = warp_sum(cadd(cadd(a0, a1), cadd(a2, a3)));
            cplx d1 = warp_sum(cadd(cadd(b0, b1), cadd(b2, b3)));
            if (lane < 2 && (lane == 0 || two)) {
                const int jc = j + lane;
                cplx d = lane ? d1 : d0;
                for (int jj = 0; jj < i; ++jj) {
                    d = csub(d, cmul(Yb[jc + (long long)ld * jj], zv[jj]));
                    d = csub(d, cmul(Ub[jc + (long long)ld * jj], zx[jj]));
                }
                Yb[jc + (long long)ld * i] = cmul(tau, d);
            }
        }
        cluster_barrier(csize);                 // column i of Y is complete (its entries were computed by different CTAs of the cluster)
        if (tid <= i) vrow[tid] = Vb[c + (long long)ld * tid];
        if (tid < i) xrow[tid] = Xb[c + (long long)ld * tid];
        __syncthreads();
        // ---------- row c of (A_i - v y^H), conjugated, on columns j > c ----------
        for (int j = c + 1 + tid; j < m; j += E_THREADS) {
            cplx acc = Ab[c + (long long)ld * j];
            for (int jj = 0; jj <= i; ++jj) acc = csub(acc, cmul(vrow[jj], cconj(Yb[j + (long long)ld * jj])));
            for (int jj = 0; jj < i; ++jj) acc = csub(acc, cmul(xrow[jj], cconj(Ub[j + (long long)ld * jj])));
            uvec[j] = cconj(acc);
        }
        __syncthreads();
        // ---------- right reflector from z = conj(row)[c+1:] ----------
        double part2 = 0.0;
        for (int j = c + 2 + tid; j < m; j += E_THREADS) part2 += cabs2(uvec[j]);
        const double xn2 = block_sum(part2, red);
        const cplx al2 = uvec[c + 1];
        cplx pi_ = mkc(0.0, 0.0), scale2 = mkc(0.0, 0.0);
        double beta2 = al2.x;
        const bool triv2 = (xn2 == 0.0 && al2.y == 0.0);
        if (!triv2) {
            beta2 = -copysign(sqrt(cabs2(al2) + xn2), al2.x);
            pi_ = mkc((beta2 - al2.x) / beta2, -al2.y / beta2);
            scale2 = cdiv(mkc(1.0, 0.0), mkc(al2.x - beta2, al2.y));
        }
        __syncthreads();
        for (int j = c + 1 + tid; j < m; j += E_THREADS) {
            cplx uu;
            if (j == c + 1) { uu = mkc(1.0, 0.0); if (csize == 1) Ab[c + (long long)ld * j] = mkc(beta2, 0.0); }
            else { uu = triv2 ? mkc(0.0, 0.0) : cmul(uvec[j], scale2); if (csize == 1) Ab[c + (long long)ld * j] = uu; }
            uvec[j] = uu;
            Ub[j + (long long)ld * i] = uu;
        }
        if (tid == 0) eb[c] = beta2;
        __syncthreads();
        // ---------- zy = Y^H u (i+1 dots), zu = U^H u (i dots) ; TP column ----------
        for (int w = warp; w < 2 * i + 1; w += E_NWARPS) {
            const int jj = (w <= i) ? w : (w - i - 1);
            const bool isU = (w > i);
            const cplx* src = (isU ? Ub : Yb) + (long long)ld * jj;
            cplx d = mkc(0.0, 0.0);
            for (int j = c + 1 + lane; j < m; j += 32) d = cfmac(src[j], uvec[j], d);
            d = warp_sum(d);
            if (lane == 0) { if (isU) zu[jj] = d; else zy[jj] = d; }
        }
        __syncthreads();
        if (tid < i) {
            cplx a = mkc(0.0, 0.0);
            for (int q = tid; q < i; ++q) a = cfma(TP[tid + BD_NB * q], zu[q], a);
            TP[tid + BD_NB * i] = cneg(cmul(pi_, a));
        }
        if (tid == 0) TP[i + BD_NB * i] = pi_;
        // ---------- x = pi (A u - V zy - X zu) on rows r > c   (pass 2 over the trailing matrix: row dots) ----------
        {
            const int len = m - c - 1;
            const int q0 = (int)(((long long)len * crank) / csize), q1 = (int)(((long long)len * (crank + 1)) / csize);
            cplx* mine = (csize > 1) ? ypb + ((long long)(i & 1) * csize + crank) * ld : nullptr;
            for (int r = c + 1 + tid; r < m; r += E_THREADS) {
                const cplx* row = Ab + r + (long long)ld * (c + 1);
                const cplx* uu = uvec + (c + 1);
                cplx y0 = mkc(0.0, 0.0), y1 = mkc(0.0, 0.0), y2 = mkc(0.0, 0.0), y3 = mkc(0.0, 0.0);
                int q = q0;
                for (; q + 7 < q1; q += 8) {
                    cplx a0 = row[(long long)ld * q], a1 = row[(long long)ld * (q + 1)], a2 = row[(long long)ld * (q + 2)], a3 = row[(long long)ld * (q + 3)];
                    cplx a4 = row[(long long)ld * (q + 4)], a5 = row[(long long)ld * (q + 5)], a6 = row[(long long)ld * (q + 6)], a7 = row[(long long)ld * (q + 7)];
                    y0 = cfma(a0, uu[q], y0); y1 = cfma(a1, uu[q + 1], y1); y2 = cfma(a2, uu[q + 2], y2); y3 = cfma(a3, uu[q + 3], y3);
                    y0 = cfma(a4, uu[q + 4], y0); y1 = cfma(a5, uu[q + 5], y1); y2 = cfma(a6, uu[q + 6], y2); y3 = cfma(a7, uu[q + 7], y3);
                }
                for (; q < q1; ++q) y0 = cfma(row[(long long)ld * q], uu[q], y0);
                cplx x = cadd(cadd(y0, y1), cadd(y2, y3));
                if (csize > 1) { mine[r] = x; continue; }
                for (int jj = 0; jj <= i; ++jj) x = csub(x, cmul(Vb[r + (long long)ld * jj], zy[jj]));
                for (int jj = 0; jj < i; ++jj) x = csub(x, cmul(Xb[r + (long long)ld * jj], zu[jj]));
                Xb[r + (long long)ld * i] = cmul(pi_, x);
            }
            if (csize > 1) {
                cluster_barrier(csize);
                const cplx* parts = ypb + (long long)(i & 1) * csize * ld;
                for (int r = c + 1 + tid; r < m; r += E_THREADS) {
                    cplx x = parts[r];
                    for (int k = 1; k < csize; ++k) x = cadd(x, parts[(long long)k * ld + r]);
                    for (int jj = 0; jj <= i; ++jj) x = csub(x, cmul(Vb[r + (long long)ld * jj], zy[jj]));
                    for (int jj = 0; jj < i; ++jj) x = csub(x, cmul(Xb[r + (long long)ld * jj], zu[jj]));
                    Xb[r + (long long)ld * i] = cmul(pi_, x);          // every CTA writes the same value
                }
                // deferred stores of the two reflectors of this step (the whole cluster is past its reads of column c and row c of A)
                for (int r = c + tid; r < m; r += E_THREADS) Ab[r + (long long)ld * c] = (r == c) ? mkc(beta, 0.0) : vvec[r];
                for (int j = c + 1 + tid; j < m; j += E_THREADS) Ab[c + (long long)ld * j] = (j == c + 1) ? mkc(beta2, 0.0) : uvec[j];
            }
        }
        __syncthreads();
    }
    cplx* TQg = TQws + (long long)b * tstride + tpan;
    cplx* TPg = TPws + (long long)b * tstride + tpan;
    for (int idx = tid; idx < BD_NB * BD_NB; idx += E_THREADS) { TQg[idx] = TQ[idx]; TPg[idx] = TP[idx]; }
}

// Rebuilds the explicit Householder vectors of one panel from the reflectors stored in A and writes VTh = V * T^H, for
//   Qacc[o:, o:] -= V ((V T^H)^H Qacc[o:, o:])     (blocked backward accumulation; o = k0 for Q, k0+1 for P)
// right = 0: left reflectors  (column c = k0+i of A, rows >= c, unit at row c)
// right = 1: right reflectors (row c of A, columns >= c+1, unit at column c+1)
__global__ void __launch_bounds__(E_THREADS, 1) bidiag_qpanel_kernel(const cplx* A, long long stride, int ld, const int* mv, int k0, int right,
                                                                     cplx* Vp, cplx* VTp, long long pstride, const cplx* Tws, long long tstride) {
    __shared__ cplx Tsm[BD_NB * BD_NB];
    const int b = blockIdx.x, m = mv[b];
    const int tid = threadIdx.x;
    const cplx* Ab = A + (long long)b * stride;
    cplx* Vb = Vp + (long long)b * pstride;
    cplx* VTb = VTp + (long long)b * pstride;
    if (k0 >= m) {
        for (int idx = tid; idx < ld * BD_NB; idx += E_THREADS) { Vb[idx] = mkc(0.0, 0.0); VTb[idx] = mkc(0.0, 0.0); }
        return;
    }
    const cplx* Tb = Tws + (long long)b * tstride + (long long)(k0 / BD_NB) * BD_NB * BD_NB;
    for (int idx = tid; idx < BD_NB * BD_NB; idx += E_THREADS) Tsm[idx] = Tb[idx];
    __syncthreads();
    for (int r = tid; r < ld; r += E_THREADS) {
        for (int q = 0; q < BD_NB; ++q) {
            const int c = k0 + q;
            cplx vv = mkc(0.0, 0.0);
            if (r < m) {
                if (!right) {
                    if (c < m) { if (r == c) vv = mkc(1.0, 0.0); else if (r > c) vv = Ab[r + (long long)ld * c]; }
                } else {
                    if (c + 1 < m) { if (r == c + 1) vv = mkc(1.0, 0.0); else if (r > c + 1) vv = Ab[c + (long long)ld * r]; }
                }
            }
            Vb[r + (long long)ld * q] = vv;
        }
        for (int jj = 0; jj < BD_NB; ++jj) {
            cplx a = mkc(0.0, 0.0);
            for (int q = jj; q < BD_NB; ++q) a = cfma(Vb[r + (long long)ld * q], cconj(Tsm[jj + BD_NB * q]), a);
            VTb[r + (long long)ld * jj] = a;
        }
    }
}
