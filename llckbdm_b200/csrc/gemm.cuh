// Batched complex-FP64 GEMM on DMMA tensor cores, one launch for a whole ensemble of members with
// per-member sizes.  C_b = opA(A_b) * B_b, column-major, 64x64 CTA tiles, K chunks of 16 staged with
// cp.async double buffering.
//
// A operand modes:
//   A_NORMAL : A_b is M x K column-major (lda)
//   A_CONJT  : A_b is stored K x M column-major (lda); the product uses conj(A_b)^T  (L^H * X)
//   A_HANKEL : A_b[i,k] = c_b[i + k + shift] -- the Hankel matrix U^shift is NEVER materialised: the
//              member's FID slice is bulk-copied (TMA, cp.async.bulk) into shared memory once per CTA
//              and every A fragment is read from it as a sliding window.
//              (reference: llckbdm/kbdm.py:95-130 builds the m x m matrices row by row on the host)
#pragma once
#include "common.cuh"

enum { A_NORMAL = 0, A_CONJT = 1, A_HANKEL = 2 };

struct GemmParams {
    const cplx* A; long long strideA; int lda;       // unused for A_HANKEL
    const cplx* B; long long strideB; int ldb;
    cplx* C;       long long strideC; int ldc;
    const int* Mv; const int* Nv; const int* Kv;     // per-member dims (device arrays, length batch; may be null)
    int Mc, Nc, Kc;                                  // constants added to the per-member dims (dims = v[b] + c)
    int accum;                                       // 0: C = A*B ; 1: C -= A*B
    int Kcap;                                        // if > 0: K = min(K, Kcap)
    int triB;                                        // B upper triangular: K is capped at the tile's last column + 1
    // Hankel source
    const cplx* sig; const long long* sig_off; int shift;
};

static inline GemmParams gemm_params_zero() {
    GemmParams p;
    p.A = nullptr; p.strideA = 0; p.lda = 0; p.B = nullptr; p.strideB = 0; p.ldb = 0; p.C = nullptr; p.strideC = 0; p.ldc = 0;
    p.Mv = p.Nv = p.Kv = nullptr; p.Mc = p.Nc = p.Kc = 0; p.accum = 0; p.Kcap = 0; p.triB = 0; p.sig = nullptr; p.sig_off = nullptr; p.shift = 0;
    return p;
}

#define G_BM 64
#define G_BN 64
#define G_BK 16
#define G_LDA_N 66   // normal A tile: As[i + 66*k]   (ld = 2 mod 8: conflict-free DMMA A-fragment reads)
#define G_LDA_T 20   // conj-trans A tile: As[k + 20*i] (ld = 4 mod 8)
#define G_LDB 20     // B tile: Bs[k + 20*j]
#define G_A_ELEMS 1280
#define G_B_ELEMS 1280
#define G_SIG_MAX 4352  // max Hankel slice (complex) kept in smem: supports m up to 2048+

// BCONJT: the B operand is stored N x K column-major and used as conj(B)^T (C -= Y * V^H)
// BREAL : the B operand has zero imaginary parts (complex x real product: half the DMMAs)
template <int AMODE, bool BCONJT, bool BREAL = false>
__global__ void __launch_bounds__(256) zgemm_batched_kernel(GemmParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* smem = reinterpret_cast<cplx*>(smem_raw);
    const int b = blockIdx.z;
    const int M = (p.Mv ? p.Mv[b] : 0) + p.Mc, N = (p.Nv ? p.Nv[b] : 0) + p.Nc;
    int K = (p.Kv ? p.Kv[b] : 0) + p.Kc;
    if (p.Kcap > 0 && K > p.Kcap) K = p.Kcap;
    if (p.triB && K > (int)blockIdx.y * G_BN + G_BN) K = (int)blockIdx.y * G_BN + G_BN;
    const int row0 = blockIdx.x * G_BM, col0 = blockIdx.y * G_BN;
    if (row0 >= M || col0 >= N || K <= 0) return;

    cplx* As[2];
    cplx* Bs[2];
    cplx* sigs = nullptr;
    if (AMODE == A_HANKEL) {
        Bs[0] = smem; Bs[1] = smem + G_B_ELEMS;
        sigs = smem + 2 * G_B_ELEMS;
        As[0] = As[1] = nullptr;
    } else {
        As[0] = smem; As[1] = smem + G_A_ELEMS;
        Bs[0] = smem + 2 * G_A_ELEMS; Bs[1] = Bs[0] + G_B_ELEMS;
    }
    const int tid = threadIdx.x, warp = tid >> 5;
    const int wr = warp >> 1, wc = warp & 1;   // warp tile: rows 16*wr.., cols 32*wc..

    const cplx* Ag = (AMODE == A_HANKEL) ? nullptr : p.A + (long long)b * p.strideA;
    const cplx* Bg = p.B + (long long)b * p.strideB;
    cplx* Cg = p.C + (long long)b * p.strideC;

    __shared__ uint64_t bar;
    int nsig = 0;
    if (AMODE == A_HANKEL) {
        // slice needed by this CTA: c[shift + row0 .. shift + row0 + 64 + K - 2]
        nsig = min(G_BM, M - row0) + K - 1;
        const cplx* src = p.sig + p.sig_off[b] + p.shift + row0;
        if (tid == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0) tma_load_1d(sigs, src, (uint32_t)nsig * 16u, &bar);
        // zero the tail so out-of-range rows/k read finite data (their products are discarded or hit zero B rows)
        for (int i = nsig + tid; i < G_BM + ((K + G_BK - 1) / G_BK) * G_BK + 8; i += 256)
            if (i < G_SIG_MAX) sigs[i] = mkc(0.0, 0.0);
    }

    auto load_tiles = [&](int buf, int k0) {
        if (AMODE == A_NORMAL) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int idx = tid + 256 * r;
                int i = idx & 63, k = idx >> 6;
                bool ok = (row0 + i < M) && (k0 + k < K);
                const cplx* src = ok ? (Ag + (row0 + i) + (long long)p.lda * (k0 + k)) : Ag;
                cp_async16(&As[buf][i + G_LDA_N * k], src, ok);
            }
        } else if (AMODE == A_CONJT) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int idx = tid + 256 * r;
                int k = idx & 15, i = idx >> 4;
                bool ok = (row0 + i < M) && (k0 + k < K);
                const cplx* src = ok ? (Ag + (k0 + k) + (long long)p.lda * (row0 + i)) : Ag;
                cp_async16(&As[buf][k + G_LDA_T * i], src, ok);
            }
        }
        if (BCONJT) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int idx = tid + 256 * r;
                int j = idx & 63, k = idx >> 6;
                bool ok = (col0 + j < N) && (k0 + k < K);
                const cplx* src = ok ? (Bg + (col0 + j) + (long long)p.ldb * (k0 + k)) : Bg;
                cp_async16(&Bs[buf][k + G_LDB * j], src, ok);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int idx = tid + 256 * r;
                int k = idx & 15, j = idx >> 4;
                bool ok = (col0 + j < N) && (k0 + k < K);
                const cplx* src = ok ? (Bg + (k0 + k) + (long long)p.ldb * (col0 + j)) : Bg;
                cp_async16(&Bs[buf][k + G_LDB * j], src, ok);
            }
        }
        cp_async_commit();
    };

    double acc[2][4][4];
    zero_acc<2, 4>(acc);

    const int nk = (K + G_BK - 1) / G_BK;
    load_tiles(0, 0);
    if (AMODE == A_HANKEL) {
        mbar_wait(&bar, 0);
        __syncthreads();
    }
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) {
            load_tiles(buf ^ 1, (kt + 1) * G_BK);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (AMODE == A_NORMAL) {
            warp_zmma<2, 4, false, BCONJT, BREAL>(acc, As[buf] + 16 * wr, 1, G_LDA_N, Bs[buf] + G_LDB * (32 * wc), 1, G_LDB, G_BK);
        } else if (AMODE == A_CONJT) {
            warp_zmma<2, 4, true, BCONJT>(acc, As[buf] + G_LDA_T * (16 * wr), G_LDA_T, 1, Bs[buf] + G_LDB * (32 * wc), 1, G_LDB, G_BK);
        } else {
            warp_zmma<2, 4, false, BCONJT>(acc, sigs + 16 * wr + kt * G_BK, 1, 1, Bs[buf] + G_LDB * (32 * wc), 1, G_LDB, G_BK);
        }
        __syncthreads();
    }

    // store
    const int lane = tid & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r = row0 + 16 * wr + 8 * i + g;
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = col0 + 32 * wc + 8 * j + 2 * t;
            if (p.accum) {
                if (c < N) { cplx* e = Cg + r + (long long)p.ldc * c; cplx o = *e; *e = mkc(o.x - acc[i][j][0], o.y - acc[i][j][2]); }
                if (c + 1 < N) { cplx* e = Cg + r + (long long)p.ldc * (c + 1); cplx o = *e; *e = mkc(o.x - acc[i][j][1], o.y - acc[i][j][3]); }
            } else {
                if (c < N) Cg[r + (long long)p.ldc * c] = mkc(acc[i][j][0], acc[i][j][2]);
                if (c + 1 < N) Cg[r + (long long)p.ldc * (c + 1)] = mkc(acc[i][j][1], acc[i][j][3]);
            }
        }
    }
}

// Skinny variant for the block-reflector products W = (V T)^H X with M <= 32 rows (A stored K x M, used as conj(A)^T):
// 32 x 128 tiles (the 64 x 64 tile would issue half of its DMMAs on padding rows).
#define GW_BN 128
#define GW_A_ELEMS (G_LDA_T * 32)
#define GW_B_ELEMS (G_LDB * GW_BN)
__global__ void __launch_bounds__(256) zgemm_w32_kernel(GemmParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* smem = reinterpret_cast<cplx*>(smem_raw);
    const int b = blockIdx.z;
    const int M = (p.Mv ? p.Mv[b] : 0) + p.Mc, N = (p.Nv ? p.Nv[b] : 0) + p.Nc;
    int K = (p.Kv ? p.Kv[b] : 0) + p.Kc;
    if (p.Kcap > 0 && K > p.Kcap) K = p.Kcap;
    const int col0 = blockIdx.y * GW_BN;
    if (M <= 0 || col0 >= N || K <= 0) return;
    cplx* As[2] = {smem, smem + GW_A_ELEMS};
    cplx* Bs[2] = {smem + 2 * GW_A_ELEMS, smem + 2 * GW_A_ELEMS + GW_B_ELEMS};
    const int tid = threadIdx.x, warp = tid >> 5;
    const int wr = warp >> 2, wc = warp & 3;        // warp tile: rows 16*wr.., cols 32*wc..
    const cplx* Ag = p.A + (long long)b * p.strideA;
    const cplx* Bg = p.B + (long long)b * p.strideB;
    cplx* Cg = p.C + (long long)b * p.strideC;
    auto load_tiles = [&](int buf, int k0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            int idx = tid + 256 * r;
            int k = idx & 15, i = idx >> 4;
            bool ok = (i < M) && (k0 + k < K);
            const cplx* src = ok ? (Ag + (k0 + k) + (long long)p.lda * i) : Ag;
            cp_async16(&As[buf][k + G_LDA_T * i], src, ok);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int idx = tid + 256 * r;
            int k = idx & 15, j = idx >> 4;
            bool ok = (col0 + j < N) && (k0 + k < K);
            const cplx* src = ok ? (Bg + (k0 + k) + (long long)p.ldb * (col0 + j)) : Bg;
            cp_async16(&Bs[buf][k + G_LDB * j], src, ok);
        }
        cp_async_commit();
    };
    double acc[2][4][4];
    zero_acc<2, 4>(acc);
    const int nk = (K + G_BK - 1) / G_BK;
    load_tiles(0, 0);
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) { load_tiles(buf ^ 1, (kt + 1) * G_BK); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        warp_zmma<2, 4, true, false>(acc, As[buf] + G_LDA_T * (16 * wr), G_LDA_T, 1, Bs[buf] + G_LDB * (32 * wc), 1, G_LDB, G_BK);
        __syncthreads();
    }
    const int lane = tid & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r = 16 * wr + 8 * i + g;
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = col0 + 32 * wc + 8 * j + 2 * t;
            if (c < N) Cg[r + (long long)p.ldc * c] = mkc(acc[i][j][0], acc[i][j][2]);
            if (c + 1 < N) Cg[r + (long long)p.ldc * (c + 1)] = mkc(acc[i][j][1], acc[i][j][3]);
        }
    }
}

// Rank-k update  C -= A * op(B)  with small K (<= KMAX = 32 or 64): the trailing / accumulation updates of the blocked
// bidiagonalisation and Hessenberg reductions.  One 64x64 tile per CTA is latency bound here (two k-tiles of work between the
// operand loads and the read-modify-write of C), so a CTA owns a 64-column block of C and streams through its row tiles:
// op(B) stays in shared memory, the next A tile arrives by cp.async and the C fragments of the current tile are fetched into
// registers before its DMMAs are issued, so every global latency is covered by tensor work.
#define GR_LDA 66          // A tile: As[i + 66*k]      (= 2 mod 8)
// GR_BN columns per CTA, 16 x 16 warp tiles: BN = 64 -> 16 warps, one CTA per SM; BN = 32 -> 8 warps, two CTAs per SM whose
// epilogues (C read-modify-write) overlap each other's DMMAs.
template <int KMAX, bool BCONJT, int BN>
__global__ void __launch_bounds__(BN * 8, 64 / BN) zgemm_rankk_kernel(GemmParams p) {
    constexpr int LDB = KMAX + 4;                        // B tile: Bs[k + LDB*j]  (= 4 mod 8)
    constexpr int NT_ = BN * 8;                          // threads
    constexpr int WC = BN / 16;                          // warps across the columns
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* Bs = reinterpret_cast<cplx*>(smem_raw);
    cplx* As[2] = {Bs + LDB * BN, Bs + LDB * BN + GR_LDA * KMAX};
    const int b = blockIdx.y;
    const int M = (p.Mv ? p.Mv[b] : 0) + p.Mc, N = (p.Nv ? p.Nv[b] : 0) + p.Nc;
    int K = (p.Kv ? p.Kv[b] : 0) + p.Kc;
    if (p.Kcap > 0 && K > p.Kcap) K = p.Kcap;
    const int col0 = blockIdx.x * BN;
    if (M <= 0 || col0 >= N || K <= 0) return;
    const int K4 = (K + 3) & ~3;                         // zero-filled up to a multiple of 4
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wr = warp / WC, wc = warp % WC;            // warp tile: 16 rows x 16 columns
    const cplx* Ag = p.A + (long long)b * p.strideA;
    const cplx* Bg = p.B + (long long)b * p.strideB;
    cplx* Cg = p.C + (long long)b * p.strideC;
    // op(B) block, once
    for (int idx = tid; idx < BN * KMAX; idx += NT_) {
        int k, j;
        if (BCONJT) { j = idx % BN; k = idx / BN; } else { k = idx % KMAX; j = idx / KMAX; }
        const bool ok = (col0 + j < N) && (k < K);
        const cplx* src = ok ? (BCONJT ? (Bg + (col0 + j) + (long long)p.ldb * k) : (Bg + k + (long long)p.ldb * (col0 + j))) : Bg;
        cp_async16(&Bs[k + LDB * j], src, ok);
    }
    auto load_a = [&](int buf, int row0) {
        for (int idx = tid; idx < 64 * KMAX; idx += NT_) {
            const int i = idx & 63, k = idx >> 6;
            const bool ok = (row0 + i < M) && (k < K);
            const cplx* src = ok ? (Ag + (row0 + i) + (long long)p.lda * k) : Ag;
            cp_async16(&As[buf][i + GR_LDA * k], src, ok);
        }
        cp_async_commit();
    };
    const int ntiles = (M + 63) / 64;
    load_a(0, 0);                                        // commits the B block together with the first A tile
    for (int rt = 0; rt < ntiles; ++rt) {
        const int buf = rt & 1, row0 = rt * 64;
        if (rt + 1 < ntiles) load_a(buf ^ 1, row0 + 64);
        // C fragments of this tile: issued now, consumed after the DMMAs
        cplx cold[2][2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = row0 + 16 * wr + 8 * i + g;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = col0 + 16 * wc + 8 * j + 2 * t;
                cold[i][j][0] = (r < M && c < N) ? Cg[r + (long long)p.ldc * c] : mkc(0.0, 0.0);
                cold[i][j][1] = (r < M && c + 1 < N) ? Cg[r + (long long)p.ldc * (c + 1)] : mkc(0.0, 0.0);
            }
        }
        if (rt + 1 < ntiles) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        double acc[2][2][4];
        zero_acc<2, 2>(acc);
        warp_zmma<2, 2, false, BCONJT>(acc, As[buf] + 16 * wr, 1, GR_LDA, Bs + LDB * (16 * wc), 1, LDB, K4);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = row0 + 16 * wr + 8 * i + g;
            if (r >= M) continue;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = col0 + 16 * wc + 8 * j + 2 * t;
                if (c < N) Cg[r + (long long)p.ldc * c] = mkc(cold[i][j][0].x - acc[i][j][0], cold[i][j][0].y - acc[i][j][2]);
                if (c + 1 < N) Cg[r + (long long)p.ldc * (c + 1)] = mkc(cold[i][j][1].x - acc[i][j][1], cold[i][j][1].y - acc[i][j][3]);
            }
        }
        __syncthreads();                                 // As[buf] is refilled two iterations ahead
    }
}

template <int KMAX, bool BCONJT>
static inline cudaError_t zgemm_rankk_launch(const GemmParams& p, int Nmax, int batch, cudaStream_t stream) {
    // K <= 32: 32-column blocks, two CTAs per SM (measured 90.5 vs 89.3 solves/s); K = 64 does not fit twice in shared memory
    LLCK_LAUNCHED();
    if (KMAX <= 32) {
        const size_t smem = (size_t)((KMAX + 4) * 32 + 2 * GR_LDA * KMAX) * sizeof(cplx);
        cudaError_t e = cudaFuncSetAttribute(zgemm_rankk_kernel<KMAX, BCONJT, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        dim3 grid((Nmax + 31) / 32, batch);
        zgemm_rankk_kernel<KMAX, BCONJT, 32><<<grid, 256, smem, stream>>>(p);
        return cudaGetLastError();
    }
    const size_t smem = (size_t)((KMAX + 4) * 64 + 2 * GR_LDA * KMAX) * sizeof(cplx);
    cudaError_t e = cudaFuncSetAttribute(zgemm_rankk_kernel<KMAX, BCONJT, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((Nmax + 63) / 64, batch);
    zgemm_rankk_kernel<KMAX, BCONJT, 64><<<grid, 512, smem, stream>>>(p);
    return cudaGetLastError();
}

static inline size_t zgemm_smem_bytes(int amode, int Kmax) {
    if (amode == A_HANKEL) return (size_t)(2 * G_B_ELEMS + G_SIG_MAX) * sizeof(cplx);
    return (size_t)(2 * G_A_ELEMS + 2 * G_B_ELEMS) * sizeof(cplx);
}

// Launch: grid = (ceil(Mmax/64), ceil(Nmax/64), batch)
template <int AMODE, bool BCONJT, bool BREAL = false>
static inline cudaError_t zgemm_launch(const GemmParams& p, dim3 grid, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(zgemm_batched_kernel<AMODE, BCONJT, BREAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    zgemm_batched_kernel<AMODE, BCONJT, BREAL><<<grid, 256, smem, stream>>>(p);
    LLCK_LAUNCHED();
    return cudaGetLastError();
}

static inline cudaError_t zgemm_batched(int amode, const GemmParams& p, int Mmax, int Nmax, int Kmax, int batch,
                                        cudaStream_t stream, bool bconjt = false, bool breal = false) {
    if (batch <= 0 || Mmax <= 0 || Nmax <= 0 || Kmax <= 0) return cudaSuccess;
    dim3 grid((Mmax + G_BM - 1) / G_BM, (Nmax + G_BN - 1) / G_BN, batch);
    size_t smem = zgemm_smem_bytes(amode, Kmax);
    if (amode == A_NORMAL && !breal && p.accum && !p.triB && Kmax <= 64 && batch <= 65535) {
        if (Kmax <= 32) return bconjt ? zgemm_rankk_launch<32, true>(p, Nmax, batch, stream) : zgemm_rankk_launch<32, false>(p, Nmax, batch, stream);
        return bconjt ? zgemm_rankk_launch<64, true>(p, Nmax, batch, stream) : zgemm_rankk_launch<64, false>(p, Nmax, batch, stream);
    }
    if (amode == A_NORMAL && breal) return zgemm_launch<A_NORMAL, false, true>(p, grid, smem, stream);
    if (amode == A_NORMAL) return bconjt ? zgemm_launch<A_NORMAL, true>(p, grid, smem, stream) : zgemm_launch<A_NORMAL, false>(p, grid, smem, stream);
    if (amode == A_CONJT && Mmax <= 32 && !p.accum && !p.triB) {
        const size_t sm32 = (size_t)(2 * GW_A_ELEMS + 2 * GW_B_ELEMS) * sizeof(cplx);
        cudaError_t e = cudaFuncSetAttribute(zgemm_w32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm32);
        if (e != cudaSuccess) return e;
        dim3 g32(1, (Nmax + GW_BN - 1) / GW_BN, batch);
        zgemm_w32_kernel<<<g32, 256, sm32, stream>>>(p);
        LLCK_LAUNCHED();
        return cudaGetLastError();
    }
    if (amode == A_CONJT) return zgemm_launch<A_CONJT, false>(p, grid, smem, stream);
    if (64 + Kmax + 32 > G_SIG_MAX) return cudaErrorInvalidValue;
    return zgemm_launch<A_HANKEL, false>(p, grid, smem, stream);
}
