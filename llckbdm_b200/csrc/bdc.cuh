// SVD of the REAL upper-bidiagonal B (from bidiag.cuh) by DIVIDE AND CONQUER -- second half of the replacement of
// scipy.linalg.svd (reference llckbdm/kbdm.py:166; LAPACK zgesdd is itself bidiagonalisation + divide and conquer).
//
// Formulation: the Golub-Kahan tridiagonal T (order N = 2m, zero diagonal, off-diagonals d1,e1,d2,e2,...,dm) of the
// perfect-shuffled [[0, B^T], [B, 0]] has the eigenpairs (+-sigma_k, (v1,u1,v2,u2,...)/sqrt2).  T is torn into <= 64-wide
// leaves (Cuppen rank-one tearing), the leaves are solved by implicit QL in shared memory (one warp each) and merged level
// by level; every level is seven batched kernels over (merge, member):
//   setup    : z vector, merge-sort of the poles, deflation (tiny z / close poles with Givens rotations), K ordering
//   secular  : one warp per root of 1 + rho sum z_i^2/(d_i - lambda), origin shifted to the nearer pole, two-pole rational
//              (Bunch-Nielsen-Sorensen) iteration with a bisection safeguard
//   order    : final sorted positions of roots and deflated eigenvalues
//   zhat     : Gu-Eisenstat / Loewner recomputation of z (numerically orthogonal eigenvectors)
//   xmat     : eigenvectors of the rank-one-modified diagonal matrix, X (k x k)
//   gemm     : Q_out[:, pos] = Q_in[:, nd] * X on FP64 DMMA, K restricted to the child that owns the row tile
//   copydefl : deflated columns copied to their sorted positions
// At a member's top level only the columns of the m positive eigenvalues are produced.  Finally u, v are de-interleaved and
// normalised separately (this removes the +sigma/-sigma mixing of small singular values).  Members whose smallest singular
// value is below BDC_FALLBACK_RATIO * sigma_max (numerically rank deficient: the +-sigma clusters around 0 cannot be
// separated) are flagged and solved by the one-sided Jacobi path (svd_real.cuh) instead.
#pragma once
#include "common.cuh"

#define BDC_LEAF 32
#define BDC_FALLBACK_RATIO 1e-8

struct BdcParams {
    const double* dws; const double* ews; int ld;      // bidiagonal d (m), e (m-1) per member, stride ld
    const int* mv;                                     // m per member
    int batch;
    double* Q[2]; long long qstride; int ldq;          // ping-pong eigenvector buffers, ldq x ldq doubles per member (ldq = 2 ld)
    double* X; long long xstride;                      // secular eigenvector matrices, ldq*ldq/2 doubles per member
    // per member vectors of length ldq
    double* off; double* diag; double* D[2]; double* dd; double* zz; double* dfD; double* mu; double* zhat; double* rho;
    int* ndsrc; int* dfsrc; int* org; int* posnd; int* posdf; int* kpos; int* meta;    // meta: 8 ints per merge at [8*merge]
    int* nleaf; int* levels;                           // per member
    int level;
};

__host__ __device__ __forceinline__ int bdc_nleaf(int m) {
    int N = 2 * m, nl = 1;
    while ((N + nl - 1) / nl > BDC_LEAF) nl <<= 1;
    return nl;
}
// even block boundaries: leaf t of a member with m rows and nleaf leaves starts at 2*floor(t*m/nleaf)
__device__ __forceinline__ int bdc_bnd(int t, int m, int nleaf) { return 2 * (int)(((long long)t * m) / nleaf); }

enum { BM_K = 0, BM_K1 = 1, BM_K2 = 2, BM_J0 = 3, BM_NDF = 4 };

// ---------------------------------------------------------------------------------------------------------------------
__global__ void bdc_init_kernel(BdcParams p) {
    const int b = blockIdx.x, m = p.mv[b], N = 2 * m;
    const double* d = p.dws + (long long)b * p.ld;
    const double* e = p.ews + (long long)b * p.ld;
    double* off = p.off + (long long)b * p.ldq;
    double* dg = p.diag + (long long)b * p.ldq;
    const int nl = bdc_nleaf(m);
    if (threadIdx.x == 0) {
        p.nleaf[b] = nl;
        int lv = 0;
        while ((1 << lv) < nl) ++lv;
        p.levels[b] = lv;
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        if (i < N - 1) off[i] = (i & 1) ? e[i >> 1] : d[i >> 1];
        dg[i] = 0.0;
    }
    __syncthreads();
    // tearing: at every cut c the diagonal entries c-1 and c lose |off[c-1]|
    for (int t = 1 + threadIdx.x; t < nl; t += blockDim.x) {
        const int c = bdc_bnd(t, m, nl);
        const double a = fabs(off[c - 1]);
        dg[c - 1] -= a;        // cuts are >= 2 apart (leaves hold >= 2 rows), no two cuts touch the same entry
        dg[c] -= a;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// leaf: implicit QL with Wilkinson shift on a <= 64 x 64 symmetric tridiagonal, one warp; eigenvectors in shared memory
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) bdc_leaf_kernel(BdcParams p) {
    __shared__ double z[BDC_LEAF * BDC_LEAF];
    __shared__ double d[BDC_LEAF], e[BDC_LEAF];
    __shared__ int rank[BDC_LEAF];
    const int b = blockIdx.y, m = p.mv[b], nl = p.nleaf[b], leaf = blockIdx.x;
    if (leaf >= nl) return;
    const int lo = bdc_bnd(leaf, m, nl), hi = bdc_bnd(leaf + 1, m, nl), n = hi - lo;
    const int lane = threadIdx.x;
    const double* off = p.off + (long long)b * p.ldq;
    const double* dg = p.diag + (long long)b * p.ldq;
    for (int i = lane; i < n; i += 32) {
        d[i] = dg[lo + i];
        e[i] = (i < n - 1) ? off[lo + i] : 0.0;
    }
    for (int idx = lane; idx < n * n; idx += 32) z[idx] = ((idx % n) == (idx / n)) ? 1.0 : 0.0;
    __syncwarp();
    double an = 0.0;
    for (int i = lane; i < n; i += 32) an = fmax(an, fabs(d[i]) + fabs(e[i]) + (i > 0 ? fabs(e[i - 1]) : 0.0));
    an = warp_max(an);
    const double tol = LLCK_EPS * an;
    for (int l = 0; l < n; ++l) {
        int iter = 0;
        while (true) {
            int mm = l;
            while (mm < n - 1 && fabs(e[mm]) > tol) ++mm;
            if (mm == l) break;
            if (++iter > 80) break;
            double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
            double r = hypot(g, 1.0);
            g = d[mm] - d[l] + e[l] / (g + copysign(r, g));
            double s = 1.0, c = 1.0, pp = 0.0;
            int i;
            bool zero_r = false;
            __syncwarp();
            for (i = mm - 1; i >= l; --i) {
                const double ei = e[i], dip1 = d[i + 1], di = d[i];
                __syncwarp();                      // every lane has read step i's inputs before lane 0 overwrites them
                double f = s * ei, bb = c * ei;
                r = hypot(f, g);
                if (lane == 0) e[i + 1] = r;
                if (r == 0.0) {
                    if (lane == 0) { d[i + 1] = dip1 - pp; e[mm] = 0.0; }
                    zero_r = true;
                    break;
                }
                s = f / r; c = g / r;
                g = dip1 - pp;
                r = (di - g) * s + 2.0 * c * bb;
                pp = s * r;
                if (lane == 0) d[i + 1] = g + pp;
                g = c * r - bb;
                for (int k = lane; k < n; k += 32) {
                    const double f2 = z[k + n * (i + 1)], zi = z[k + n * i];
                    z[k + n * (i + 1)] = s * zi + c * f2;
                    z[k + n * i] = c * zi - s * f2;
                }
            }
            __syncwarp();
            if (!zero_r && lane == 0) { d[l] -= pp; e[l] = g; e[mm] = 0.0; }
            __syncwarp();
        }
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        int r = 0;
        const double di = d[i];
        for (int j = 0; j < n; ++j) r += (d[j] < di || (d[j] == di && j < i)) ? 1 : 0;
        rank[i] = r;
    }
    __syncwarp();
    double* Dout = p.D[0] + (long long)b * p.ldq;
    double* Q = p.Q[0] + (long long)b * p.qstride;
    for (int i = lane; i < n; i += 32) Dout[lo + rank[i]] = d[i];
    for (int i = 0; i < n; ++i) {
        double* col = Q + (lo) + (long long)p.ldq * (lo + rank[i]);
        for (int k = lane; k < n; k += 32) col[k] = z[k + n * i];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// merge geometry of (member b, merge index i) at level p.level
// ---------------------------------------------------------------------------------------------------------------------
struct BdcSeg { int lo, mid, hi, top, m; };
__device__ __forceinline__ bool bdc_segment(const BdcParams& p, int b, int i, BdcSeg& s) {
    const int lv = p.levels[b];
    if (p.level >= lv) return false;
    const int nl = p.nleaf[b];
    const int span = 1 << (p.level + 1);
    if (i * span >= nl) return false;
    s.m = p.mv[b];
    s.lo = bdc_bnd(i * span, s.m, nl);
    s.mid = bdc_bnd(i * span + span / 2, s.m, nl);
    s.hi = bdc_bnd((i + 1) * span, s.m, nl);
    s.top = (p.level == lv - 1) ? 1 : 0;
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// setup: z, sorted poles, deflation.  One CTA per merge.  Dynamic smem: Nmax * (8 + 8 + 4 + 4 + 4 + 4) bytes
// ---------------------------------------------------------------------------------------------------------------------
#define BDC_SETUP_THREADS 256
__global__ void __launch_bounds__(BDC_SETUP_THREADS) bdc_setup_kernel(BdcParams p, int nmax) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Ds = reinterpret_cast<double*>(smem_raw);
    double* zs = Ds + nmax;
    int* src = reinterpret_cast<int*>(zs + nmax);
    int* typ = src + nmax;       // 1: child 1, 2: mixed, 3: child 2, 4: deflated
    int* ndl = typ + nmax;
    int* dfl = ndl + nmax;
    __shared__ double red[32];
    const int b = blockIdx.y;
    BdcSeg sg;
    if (!bdc_segment(p, b, blockIdx.x, sg)) return;
    const int lo = sg.lo, mid = sg.mid, hi = sg.hi, Nm = hi - lo, n1 = mid - lo;
    const int tid = threadIdx.x;
    const int cur = p.level & 1;
    double* Qin = p.Q[cur] + (long long)b * p.qstride;
    const double* Din = p.D[cur] + (long long)b * p.ldq;
    const long long vb = (long long)b * p.ldq;
    const double beta = p.off[vb + mid - 1];
    const double rho = 2.0 * fabs(beta);
    const double sgn = (beta >= 0.0) ? 1.0 : -1.0;
    const double rs2 = 0.70710678118654752440;
    // merge-rank sort of the two sorted children
    for (int t = tid; t < Nm; t += BDC_SETUP_THREADS) {
        const double dv = Din[lo + t];
        int rnk;
        if (t < n1) {       // # of child-2 poles strictly below
            int a = 0, c = Nm - n1;
            while (a < c) { int h = (a + c) >> 1; if (Din[mid + h] < dv) a = h + 1; else c = h; }
            rnk = t + a;
        } else {            // # of child-1 poles <= dv
            int a = 0, c = n1;
            while (a < c) { int h = (a + c) >> 1; if (Din[lo + h] <= dv) a = h + 1; else c = h; }
            rnk = (t - n1) + a;
        }
        const double zv = (t < n1) ? Qin[(mid - 1) + (long long)p.ldq * (lo + t)] : sgn * Qin[mid + (long long)p.ldq * (lo + t)];
        Ds[rnk] = dv; zs[rnk] = zv * rs2; src[rnk] = lo + t; typ[rnk] = (t < n1) ? 1 : 3;
    }
    __syncthreads();
    double mxd = 0.0, mxz = 0.0;
    for (int t = tid; t < Nm; t += BDC_SETUP_THREADS) { mxd = fmax(mxd, fabs(Ds[t])); mxz = fmax(mxz, fabs(zs[t])); }
    mxd = block_max(mxd, red);
    __syncthreads();
    mxz = block_max(mxz, red);
    __syncthreads();
    const double tol = 8.0 * LLCK_EPS * fmax(mxd, mxz);
    int k = 0, ndf = 0;
    if (rho * mxz <= tol) {
        for (int t = tid; t < Nm; t += BDC_SETUP_THREADS) { typ[t] = 4; dfl[t] = t; }
        ndf = Nm;
        __syncthreads();
    } else {
        for (int t = tid; t < Nm; t += BDC_SETUP_THREADS) if (rho * fabs(zs[t]) <= tol) typ[t] = 4;
        __syncthreads();
        // close poles: sequential scan, executed redundantly by every thread (identical arithmetic); shared state is only
        // modified when a rotation happens, between barriers, and the rotation is applied to the Q columns by all threads
        int pj = -1;
        for (int nj = 0; nj < Nm; ++nj) {
            if (typ[nj] == 4) continue;
            if (pj < 0) { pj = nj; continue; }
            double s = zs[pj], c = zs[nj];
            const double tau = hypot(c, s);
            const double t = Ds[nj] - Ds[pj];
            c /= tau; s = -s / tau;
            if (fabs(t * c * s) <= tol) {
                const double dpj = Ds[pj], dnj = Ds[nj];
                const int tp = typ[pj], tn = typ[nj];
                double* cp = Qin + lo + (long long)p.ldq * src[pj];
                double* cn = Qin + lo + (long long)p.ldq * src[nj];
                for (int r = tid; r < Nm; r += BDC_SETUP_THREADS) {
                    const double x = cp[r], y = cn[r];
                    cp[r] = c * x + s * y;
                    cn[r] = c * y - s * x;
                }
                __syncthreads();
                if (tid == 0) {
                    zs[nj] = tau; zs[pj] = 0.0;
                    Ds[pj] = dpj * c * c + dnj * s * s;
                    Ds[nj] = dpj * s * s + dnj * c * c;
                    typ[pj] = 4;
                    if (tp != tn) typ[nj] = 2;
                }
                __syncthreads();
            }
            pj = nj;
        }
        // lists (thread 0; Nm <= 4096 steps)
        if (tid == 0) {
            int a = 0, c = 0;
            for (int t = 0; t < Nm; ++t) { if (typ[t] == 4) dfl[c++] = t; else ndl[a++] = t; }
            red[0] = (double)a;
        }
        __syncthreads();
        k = (int)red[0];
        ndf = Nm - k;
    }
    // K order of the non-deflated columns: child-1 only, mixed, child-2 only
    int k1 = 0, k2 = 0;
    if (k > 0) {
        if (tid == 0) {
            int c1 = 0, c2 = 0;
            for (int t = 0; t < k; ++t) { const int ty = typ[ndl[t]]; c1 += (ty == 1); c2 += (ty == 2); }
            red[1] = (double)c1; red[2] = (double)c2;
        }
        __syncthreads();
        k1 = (int)red[1]; k2 = (int)red[2];
        if (tid == 0) {
            int a1 = 0, a2 = k1, a3 = k1 + k2;
            for (int t = 0; t < k; ++t) {
                const int ty = typ[ndl[t]];
                const int pos = (ty == 1) ? a1++ : ((ty == 2) ? a2++ : a3++);
                p.kpos[vb + lo + t] = pos;
                p.ndsrc[vb + lo + pos] = src[ndl[t]];
            }
        }
    }
    for (int t = tid; t < k; t += BDC_SETUP_THREADS) { p.dd[vb + lo + t] = Ds[ndl[t]]; p.zz[vb + lo + t] = zs[ndl[t]]; }
    for (int t = tid; t < ndf; t += BDC_SETUP_THREADS) { p.dfD[vb + lo + t] = Ds[dfl[t]]; p.dfsrc[vb + lo + t] = src[dfl[t]]; }
    if (tid == 0) {
        int* mt = p.meta + vb + 8 * blockIdx.x;            // 8 ints per merge; merges per member < ldq / 64
        mt[BM_K] = k; mt[BM_K1] = k1; mt[BM_K2] = k2; mt[BM_NDF] = ndf; mt[BM_J0] = 0;
        p.rho[vb + lo] = rho;
    }
}

__device__ __forceinline__ int* bdc_meta(const BdcParams& p, int b, int i) { return p.meta + (long long)b * p.ldq + 8 * i; }

// ---------------------------------------------------------------------------------------------------------------------
// secular equation: one warp per root
// ---------------------------------------------------------------------------------------------------------------------
#define BDC_SEC_RPC 32      // roots per CTA (8 warps x 4)
__global__ void __launch_bounds__(256) bdc_secular_kernel(BdcParams p) {
    const int b = blockIdx.z;
    BdcSeg sg;
    if (!bdc_segment(p, b, blockIdx.y, sg)) return;
    const int* mt = bdc_meta(p, b, blockIdx.y);
    const int k = mt[BM_K];
    if ((int)blockIdx.x * BDC_SEC_RPC >= k) return;
    const long long vb = (long long)b * p.ldq + sg.lo;
    const double* dd = p.dd + vb;
    const double* zz = p.zz + vb;
    const double rho = p.rho[vb];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int rr = 0; rr < BDC_SEC_RPC / 8; ++rr) {
        const int j = blockIdx.x * BDC_SEC_RPC + warp + 8 * rr;
        if (j >= k) break;
        const bool last = (j == k - 1);
        double gap;
        if (!last) gap = dd[j + 1] - dd[j];
        else {
            double sacc = 0.0;
            for (int i = lane; i < k; i += 32) sacc = fma(zz[i], zz[i], sacc);
            gap = rho * warp_sum(sacc);
        }
        // sign of the secular function at the interval midpoint decides the origin
        int o = j;
        double lo = 0.0, hi = last ? gap : 0.5 * gap, mcur = 0.5 * gap;
        if (!last) {
            const double dj = dd[j];
            double acc = 0.0;
            for (int i = lane; i < k; i += 32) acc += rho * zz[i] * zz[i] / ((dd[i] - dj) - mcur);
            const double gmid = 1.0 + warp_sum(acc);
            if (!(gmid >= 0.0)) { o = j + 1; lo = -0.5 * gap; hi = 0.0; mcur = -0.5 * gap; }
        }
        const double dorg = dd[o];
        const double dL = dd[j] - dorg;
        const double dR = last ? 0.0 : dd[j + 1] - dorg;
        for (int it = 0; it < 80; ++it) {
            double psi = 0.0, phi = 0.0, dpsi = 0.0, dphi = 0.0;
            for (int i = lane; i < k; i += 32) {
                const double r = 1.0 / ((dd[i] - dorg) - mcur);
                const double t = rho * zz[i] * zz[i] * r;
                if (i <= j) { psi += t; dpsi = fma(t, r, dpsi); } else { phi += t; dphi = fma(t, r, dphi); }
            }
            psi = warp_sum(psi); phi = warp_sum(phi); dpsi = warp_sum(dpsi); dphi = warp_sum(dphi);
            const double g = 1.0 + psi + phi;
            if (g > 0.0) hi = fmin(hi, mcur); else lo = fmax(lo, mcur);
            if (fabs(g) <= 8.0 * LLCK_EPS * (1.0 + fabs(psi) + fabs(phi))) break;
            if ((hi - lo) <= 2.0 * LLCK_EPS * fmax(fabs(lo), fabs(hi))) break;
            const double DL = dL - mcur;
            const double a = dpsi * DL * DL, c1 = psi - dpsi * DL;
            double cand0, cand1;
            if (last) {
                const double c = 1.0 + c1;
                cand0 = (c > 0.0) ? (DL + a / c) : INFINITY;
                cand1 = INFINITY;
            } else {
                const double DR = dR - mcur;
                const double bq = dphi * DR * DR, c2 = phi - dphi * DR;
                const double c = 1.0 + c1 + c2;
                const double Bq = c * (DL + DR) + a + bq;
                const double Cq = DL * DR * g;
                if (c == 0.0) { cand0 = Cq / Bq; cand1 = INFINITY; }
                else {
                    const double disc = fma(Bq, Bq, -4.0 * c * Cq);
                    const double sq = sqrt(fmax(disc, 0.0));
                    const double q = 0.5 * (Bq + ((Bq >= 0.0) ? sq : -sq));
                    cand0 = q / c;
                    cand1 = (q != 0.0) ? Cq / q : INFINITY;
                }
            }
            double nm = mcur + cand0;
            if (!(isfinite(nm) && nm > lo && nm < hi)) {
                nm = mcur + cand1;
                if (!(isfinite(nm) && nm > lo && nm < hi)) {
                    if (lo == 0.0) nm = 0.1 * hi;
                    else if (hi == 0.0) nm = 0.1 * lo;
                    else nm = 0.5 * (lo + hi);
                }
            }
            mcur = nm;
        }
        if (lane == 0) { p.org[vb + j] = o; p.mu[vb + j] = mcur; }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// order: sorted positions of the roots and of the deflated eigenvalues inside [lo, hi); next level's pole array
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bdc_order_kernel(BdcParams p, int nmax) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* lam = reinterpret_cast<double*>(smem_raw);     // nmax
    double* dfv = lam + nmax;                               // nmax
    int* dfi = reinterpret_cast<int*>(dfv + nmax);          // nmax
    const int b = blockIdx.y;
    BdcSeg sg;
    if (!bdc_segment(p, b, blockIdx.x, sg)) return;
    int* mt = bdc_meta(p, b, blockIdx.x);
    const int k = mt[BM_K], ndf = mt[BM_NDF];
    const int Nm = sg.hi - sg.lo, tid = threadIdx.x;
    const long long vb = (long long)b * p.ldq + sg.lo;
    for (int j = tid; j < k; j += 256) lam[j] = p.dd[vb + p.org[vb + j]] + p.mu[vb + j];
    for (int t = tid; t < ndf; t += 256) { dfv[t] = p.dfD[vb + t]; dfi[t] = p.dfsrc[vb + t]; }
    __syncthreads();
    // the deflated list is sorted except where a rotation moved a pole: rank by counting (stable; the list is short)
    double* Dout = p.D[(p.level & 1) ^ 1] + (long long)b * p.ldq + sg.lo;
    const int cut = sg.top ? (Nm - sg.m) : 0;      // top level: only positions >= cut (the m positive eigenvalues) are produced
    for (int t = tid; t < ndf; t += 256) {
        const double v = dfv[t];
        int r = 0;
        for (int u = 0; u < ndf; ++u) r += (dfv[u] < v || (dfv[u] == v && u < t)) ? 1 : 0;
        int a = 0, c = k;                              // # roots <= v
        while (a < c) { int h = (a + c) >> 1; if (lam[h] <= v) a = h + 1; else c = h; }
        const int pos = r + a;
        p.posdf[vb + t] = pos;
        Dout[pos] = v;
    }
    int j0 = 0;
    for (int j = tid; j < k; j += 256) {
        const double v = lam[j];
        int r = 0;
        for (int u = 0; u < ndf; ++u) r += (dfv[u] < v) ? 1 : 0;
        const int pos = j + r;
        p.posnd[vb + j] = pos;
        Dout[pos] = v;
        if (pos < cut) j0 = j + 1;
    }
    if (sg.top) {
        __shared__ int sj0;
        if (tid == 0) sj0 = 0;
        __syncthreads();
        atomicMax(&sj0, j0);
        __syncthreads();
        if (tid == 0) mt[BM_J0] = sj0;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// zhat (Loewner formula): one warp per pole
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bdc_zhat_kernel(BdcParams p) {
    const int b = blockIdx.z;
    BdcSeg sg;
    if (!bdc_segment(p, b, blockIdx.y, sg)) return;
    const int* mt = bdc_meta(p, b, blockIdx.y);
    const int k = mt[BM_K];
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= k) return;
    const long long vb = (long long)b * p.ldq + sg.lo;
    const double* dd = p.dd + vb;
    const double di = dd[i];
    double prod = 1.0;
    for (int j = lane; j < k; j += 32) {
        const double num = (dd[p.org[vb + j]] - di) + p.mu[vb + j];      // lambda_j - d_i
        prod *= (j == i) ? num : num / (dd[j] - di);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) prod *= __shfl_xor_sync(0xffffffffu, prod, o);
    if (lane == 0) {
        const double zh = sqrt(fabs(prod / p.rho[vb]));
        p.zhat[vb + i] = copysign(zh, p.zz[vb + i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// X[kpos_i, j - j0] = zhat_i / (d_i - lambda_j), columns normalised; ld = kx = round_up(k, 2) (pad row zero); one warp per column
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double* bdc_xptr(const BdcParams& p, int b, const BdcSeg& sg) {
    // per member ldq*ldq/2 doubles; merge [lo, hi) owns [lo * ldq/2, hi * ldq/2)  (k^2 <= (hi-lo) * ldq/2 below the top level;
    // at the top level only <= m <= ldq/2 columns are produced)
    return p.X + (long long)b * p.xstride + (long long)sg.lo * (p.ldq / 2);
}
__global__ void __launch_bounds__(256) bdc_xmat_kernel(BdcParams p) {
    const int b = blockIdx.z;
    BdcSeg sg;
    if (!bdc_segment(p, b, blockIdx.y, sg)) return;
    const int* mt = bdc_meta(p, b, blockIdx.y);
    const int k = mt[BM_K], j0 = mt[BM_J0];
    const int j = j0 + blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= k) return;
    const int kx = (k + 1) & ~1;
    const long long vb = (long long)b * p.ldq + sg.lo;
    const double* dd = p.dd + vb;
    const double dorg = dd[p.org[vb + j]], muj = p.mu[vb + j];
    double* col = bdc_xptr(p, b, sg) + (long long)kx * (j - j0);
    double nrm = 0.0;
    for (int i = lane; i < k; i += 32) {
        const double x = p.zhat[vb + i] / ((dd[i] - dorg) - muj);
        nrm = fma(x, x, nrm);
    }
    nrm = warp_sum(nrm);
    const double sc = 1.0 / sqrt(nrm);
    for (int i = lane; i < k; i += 32) {
        const double x = p.zhat[vb + i] / ((dd[i] - dorg) - muj);
        col[p.kpos[vb + i]] = x * sc;
    }
    if (lane == 0 && kx > k) col[k] = 0.0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Q_out[lo:hi, lo + posnd[j]] = sum_t Q_in[lo:hi, ndsrc[t]] * X[t, j - j0]   (real FP64 DMMA, 64 x 64 tiles, BK = 16)
// ---------------------------------------------------------------------------------------------------------------------
#define BG_LDA 68      // = 4 mod 16: conflict-free LDS.64 fragment reads
#define BG_LDB 20
#define BG_A (BG_LDA * 16)
#define BG_B (BG_LDB * 64)
__global__ void __launch_bounds__(256, 2) bdc_gemm_kernel(BdcParams p, int tiles_n) {
    __shared__ __align__(16) double As[2][BG_A];
    __shared__ __align__(16) double Bs[2][BG_B];
    const int b = blockIdx.z, mi = blockIdx.y / tiles_n;
    BdcSeg sg;
    if (!bdc_segment(p, b, mi, sg)) return;
    const int* mt = bdc_meta(p, b, mi);
    const int k = mt[BM_K], k1 = mt[BM_K1], k2 = mt[BM_K2], j0 = mt[BM_J0];
    const int Nm = sg.hi - sg.lo, ncol = k - j0;
    const int r0 = blockIdx.x * 64, c0 = (blockIdx.y % tiles_n) * 64;
    if (r0 >= Nm || c0 >= ncol) return;
    const int kx = (k + 1) & ~1;
    const int cur = p.level & 1;
    const long long blk = (long long)b * p.qstride + sg.lo + (long long)p.ldq * sg.lo;     // origin of the merged diagonal block
    const double* Qin = p.Q[cur] + blk;
    double* Qout = p.Q[cur ^ 1] + blk;
    const long long vb = (long long)b * p.ldq + sg.lo;
    const int* ndsrc = p.ndsrc + vb;
    const double* X = bdc_xptr(p, b, sg);
    // K range: rows of child 1 see only child-1 and mixed columns, rows of child 2 only mixed and child-2 columns
    const int n1 = sg.mid - sg.lo;
    int kbeg = 0, kend = k;
    if (r0 + 64 <= n1) kend = k1 + k2;
    else if (r0 >= n1) kbeg = k1 & ~1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wr = warp >> 1, wc = warp & 1;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    auto load = [&](int buf, int kk0) {
        // A: 16 gathered columns x 64 rows (32 chunks of 2 rows each): chunk = tid & 31, column = tid >> 5 (+8)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int kk = (tid >> 5) + 8 * r, ch = tid & 31;
            const int kt = kk0 + kk, row = r0 + 2 * ch;
            const bool ok = (kt < kend) && (row < Nm);       // Nm even, row even: both rows valid together
            const double* srcp = ok ? (Qin + row + (long long)p.ldq * (ndsrc[kt] - sg.lo)) : Qin;
            cp_async16(&As[buf][2 * ch + BG_LDA * kk], srcp, ok);
        }
        // B: X[kk0 + 0..15, c0 + 0..63]: 8 chunks of 2 rows per column: chunk = tid & 7, column = tid >> 3 (+32)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int ch = tid & 7, jj = (tid >> 3) + 32 * r;
            const int kt = kk0 + 2 * ch, cj = c0 + jj;
            const bool ok = (kt < kend) && (cj < ncol);      // kend <= k <= kx; pad row of X is zero and its A column is zero-filled
            const double* srcp = ok ? (X + kt + (long long)kx * cj) : X;
            cp_async16(&Bs[buf][2 * ch + BG_LDB * jj], srcp, ok);
        }
        cp_async_commit();
    };
    const int nk = (kend - kbeg + 15) / 16;
    if (nk > 0) load(0, kbeg);
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) { load(buf ^ 1, kbeg + (kt + 1) * 16); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const double* ap = As[buf] + (16 * wr + g) + BG_LDA * t;
        const double* bp = Bs[buf] + t + BG_LDB * (32 * wc + g);
#pragma unroll
        for (int kk = 0; kk < 16; kk += 4) {
            double a[2], bb[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = ap[8 * i + BG_LDA * kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = bp[kk + BG_LDB * 8 * j];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], bb[j]);
        }
        __syncthreads();
    }
    const int* posnd = p.posnd + vb;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c0 + 32 * wc + 8 * j + 2 * t;
        const long long o0 = (c < ncol) ? (long long)p.ldq * posnd[j0 + c] : 0;
        const long long o1 = (c + 1 < ncol) ? (long long)p.ldq * posnd[j0 + c + 1] : 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int row = r0 + 16 * wr + 8 * i + g;
            if (row < Nm) {
                if (c < ncol) Qout[row + o0] = acc[i][j][0];
                if (c + 1 < ncol) Qout[row + o1] = acc[i][j][1];
            }
        }
    }
}

// deflated columns: copy to their sorted positions (top level: only the positive-eigenvalue positions)
__global__ void __launch_bounds__(256) bdc_copydefl_kernel(BdcParams p) {
    const int b = blockIdx.z;
    BdcSeg sg;
    if (!bdc_segment(p, b, blockIdx.y, sg)) return;
    const int* mt = bdc_meta(p, b, blockIdx.y);
    const int ndf = mt[BM_NDF];
    const int Nm = sg.hi - sg.lo;
    const int cut = sg.top ? (Nm - sg.m) : 0;
    const int cur = p.level & 1;
    const long long vb = (long long)b * p.ldq + sg.lo;
    const long long blk = (long long)b * p.qstride + sg.lo + (long long)p.ldq * sg.lo;
    const double* Qin = p.Q[cur] + blk;
    double* Qout = p.Q[cur ^ 1] + blk;
    for (int t = blockIdx.x; t < ndf; t += gridDim.x) {
        const int pos = p.posdf[vb + t];
        if (pos < cut) continue;
        const double* s = Qin + (long long)p.ldq * (p.dfsrc[vb + t] - sg.lo);
        double* d = Qout + (long long)p.ldq * pos;
        for (int r = threadIdx.x; r < Nm; r += 256) d[r] = s[r];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// final: sigma_k = k-th largest eigenvalue, u/v de-interleaved and normalised separately, scaled for the back-multiplication:
//   Lpre[:,k] = u_k * g_k^{-1/2},  Rpre[:,k] = v_k * g_k^{-1/2}   (g = s or s + q^2/s, reference kbdm.py:179-186)
// scale_mode = 1: unscaled (Lpre = U * diag(s), Rpre = V) for the stage checker.  fallback[b] = 1: member left to the Jacobi path.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void bdc_sv_kernel(BdcParams p, double* sing_vals, long long sv_stride, int* fallback) {
    const int b = blockIdx.x, m = p.mv[b], N = 2 * m;
    const int fin = p.levels[b] & 1;
    const double* D = p.D[fin] + (long long)b * p.ldq;
    for (int k = threadIdx.x; k < m; k += blockDim.x) sing_vals[(long long)b * sv_stride + k] = D[N - 1 - k];
    if (threadIdx.x == 0) {
        const double smax = D[N - 1], smin = D[m];
        fallback[b] = (!(smin > BDC_FALLBACK_RATIO * smax) || !isfinite(smax)) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(128) bdc_gather_kernel(BdcParams p, const int* lv, const double* sing_vals, long long sv_stride, double q,
                                                         cplx* Lpre, cplx* Rpre, long long cstride, int ldc, int* status,
                                                         const int* fallback, int* imbalance, int scale_mode) {
    const int b = blockIdx.y, k = blockIdx.x;
    const int m = p.mv[b], l = scale_mode ? m : lv[b];
    if (k >= l || fallback[b]) return;
    const int N = 2 * m;
    const int fin = p.levels[b] & 1;
    const double* col = p.Q[fin] + (long long)b * p.qstride + (long long)p.ldq * (N - 1 - k);
    __shared__ double red[32];
    double nu = 0.0, nv = 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const double v = col[2 * i], u = col[2 * i + 1];
        nv = fma(v, v, nv); nu = fma(u, u, nu);
    }
    nu = block_sum(nu, red);
    __syncthreads();
    nv = block_sum(nv, red);
    // safety net: an eigenvector (v1,u1,v2,u2,...)/sqrt2 of the Golub-Kahan matrix has |u|^2 = |v|^2 = 1/2; a visible imbalance means the
    // +sigma/-sigma pair was not separated -> hand the member to the Jacobi path.  The flag goes to a separate array (merged into
    // fallback[] by the kernel that follows), so no block of this kernel reads a word a sibling block is writing.
    if (threadIdx.x == 0 && !scale_mode && !(fabs(nu - 0.5) < 1e-5 && fabs(nv - 0.5) < 1e-5)) atomicExch(&imbalance[b], 1);
    const double s = sing_vals[(long long)b * sv_stride + k];
    double fu = rsqrt(nu), fv = rsqrt(nv);
    if (scale_mode) fu *= s;
    else {
        const double gq = (q > 0.0) ? (s + q * q / s) : s;
        if (!(gq > 0.0) || !isfinite(gq)) {
            if (threadIdx.x == 0) atomicMax(&status[b], 2);
            fu = 0.0; fv = 0.0;
        } else {
            const double f = 1.0 / sqrt(gq);
            fu *= f; fv *= f;
        }
    }
    cplx* ldst = Lpre + (long long)b * cstride + (long long)ldc * k;
    cplx* rdst = Rpre + (long long)b * cstride + (long long)ldc * k;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        ldst[i] = mkc(col[2 * i + 1] * fu, 0.0);
        rdst[i] = mkc(col[2 * i] * fv, 0.0);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host driver.  Workspace carving is the caller's; launches ~7 kernels per level.
// ---------------------------------------------------------------------------------------------------------------------
static inline size_t bdc_vec_bytes(int batch, int ldq) {      // all per-member vectors (10 double + 7 int arrays of ldq, + nleaf/levels)
    return (size_t)batch * ldq * (10 * sizeof(double) + 7 * sizeof(int)) + 2 * sizeof(int) * (size_t)batch + 4096;
}
static inline void bdc_carve_vectors(BdcParams& p, unsigned char* base, int batch, int ldq) {
    const size_t n = (size_t)batch * ldq;
    double* dp = reinterpret_cast<double*>(base);
    p.off = dp; p.diag = dp + n; p.D[0] = dp + 2 * n; p.D[1] = dp + 3 * n; p.dd = dp + 4 * n; p.zz = dp + 5 * n;
    p.dfD = dp + 6 * n; p.mu = dp + 7 * n; p.zhat = dp + 8 * n; p.rho = dp + 9 * n;
    int* ip = reinterpret_cast<int*>(dp + 10 * n);
    p.ndsrc = ip; p.dfsrc = ip + n; p.org = ip + 2 * n; p.posnd = ip + 3 * n; p.posdf = ip + 4 * n; p.kpos = ip + 5 * n; p.meta = ip + 6 * n;
    p.nleaf = ip + 7 * n; p.levels = p.nleaf + batch;
}

static int bdc_driver(BdcParams p, int mmax, cudaStream_t st) {
    const int batch = p.batch;
    const int nlmax = bdc_nleaf(mmax);
    int lvmax = 0;
    while ((1 << lvmax) < nlmax) ++lvmax;
    cudaError_t e;
    if ((e = cudaMemsetAsync(p.Q[0], 0, sizeof(double) * (size_t)batch * p.qstride, st)) != cudaSuccess) return -(int)e;
    if ((e = cudaMemsetAsync(p.Q[1], 0, sizeof(double) * (size_t)batch * p.qstride, st)) != cudaSuccess) return -(int)e;
    bdc_init_kernel<<<batch, 256, 0, st>>>(p);
    LLCK_LAUNCHED();
    {
        dim3 grid(nlmax, batch);
        bdc_leaf_kernel<<<grid, 32, 0, st>>>(p);
        LLCK_LAUNCHED();
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return -(int)e;
    const int Nmax = 2 * mmax;
    for (int lv = 0; lv < lvmax; ++lv) {
        p.level = lv;
        const int merges = nlmax >> (lv + 1);
        // largest merged block at this level (upper bound over members)
        int nm = BDC_LEAF << (lv + 1);       // leaves are <= BDC_LEAF wide
        if (nm > Nmax) nm = Nmax;
        const size_t sm_setup = (size_t)nm * 32 + 64;
        const size_t sm_order = (size_t)nm * 20 + 64;
        if ((e = cudaFuncSetAttribute(bdc_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_setup)) != cudaSuccess) return -(int)e;
        if ((e = cudaFuncSetAttribute(bdc_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_order)) != cudaSuccess) return -(int)e;
        dim3 g1(merges, batch);
        bdc_setup_kernel<<<g1, BDC_SETUP_THREADS, sm_setup, st>>>(p, nm);
        dim3 g2((nm + BDC_SEC_RPC - 1) / BDC_SEC_RPC, merges, batch);
        bdc_secular_kernel<<<g2, 256, 0, st>>>(p);
        bdc_order_kernel<<<g1, 256, sm_order, st>>>(p, nm);
        dim3 g3((nm + 7) / 8, merges, batch);
        bdc_zhat_kernel<<<g3, 256, 0, st>>>(p);
        bdc_xmat_kernel<<<g3, 256, 0, st>>>(p);
        const int tn = (nm + 63) / 64;
        dim3 g4(tn, tn * merges, batch);
        bdc_gemm_kernel<<<g4, 256, 0, st>>>(p, tn);
        dim3 g5(64, merges, batch);
        bdc_copydefl_kernel<<<g5, 256, 0, st>>>(p);
        llck_launch_count += 7;
        if ((e = cudaGetLastError()) != cudaSuccess) return -(int)e;
    }
    return 0;
}
