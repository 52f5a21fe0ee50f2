// Common device helpers for the llckbdm_b200 kernels (sm_100a): complex-FP64 arithmetic, warp/block
// reductions, cp.async staging and the DMMA (mma.sync m8n8k4 f64) warp-level complex tile product.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

typedef double2 cplx;

// kernels enqueued by the current llck_kbdm_batched call on this host thread (reported in info[13]); bumped at every launch site
static thread_local int llck_launch_count = 0;
#define LLCK_LAUNCHED() (++llck_launch_count)

#define LLCK_EPS 2.220446049250313e-16
#define LLCK_SAFMIN 2.2250738585072014e-308

// ---------------------------------------------------------------------------------------------
// complex helpers
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ cplx mkc(double r, double i) { return make_double2(r, i); }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return mkc(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return mkc(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return mkc(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cneg(cplx a) { return mkc(-a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cscale(cplx a, double s) { return mkc(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return mkc(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// conj(a) * b
__host__ __device__ __forceinline__ cplx cmulc(cplx a, cplx b) {
    return mkc(fma(a.x, b.x, a.y * b.y), fma(a.x, b.y, -a.y * b.x));
}
// acc + a*b
__host__ __device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx acc) {
    return mkc(fma(a.x, b.x, fma(-a.y, b.y, acc.x)), fma(a.x, b.y, fma(a.y, b.x, acc.y)));
}
// acc + conj(a)*b
__host__ __device__ __forceinline__ cplx cfmac(cplx a, cplx b, cplx acc) {
    return mkc(fma(a.x, b.x, fma(a.y, b.y, acc.x)), fma(a.x, b.y, fma(-a.y, b.x, acc.y)));
}
__host__ __device__ __forceinline__ double cabs2(cplx a) { return fma(a.x, a.x, a.y * a.y); }
__host__ __device__ __forceinline__ double cabs1(cplx a) { return fabs(a.x) + fabs(a.y); }
__host__ __device__ __forceinline__ double cabs_(cplx a) { return hypot(a.x, a.y); }
// robust complex division a / b (Smith)
__host__ __device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    if (fabs(b.x) >= fabs(b.y)) {
        double r = b.y / b.x, d = b.x + b.y * r;
        return mkc((a.x + a.y * r) / d, (a.y - a.x * r) / d);
    } else {
        double r = b.x / b.y, d = b.x * r + b.y;
        return mkc((a.x * r + a.y) / d, (a.y * r - a.x) / d);
    }
}
// principal square root
__host__ __device__ __forceinline__ cplx csqrt_(cplx z) {
    double r = hypot(z.x, z.y);
    if (r == 0.0) return mkc(0.0, 0.0);
    double s = sqrt(0.5 * (r + fabs(z.x)));
    double t = 0.5 * z.y / s;
    if (z.x >= 0) return mkc(s, t);
    return mkc(fabs(t), copysign(s, z.y));
}

// Givens rotation: c real, s complex with [c s; -conj(s) c] * [a; b] = [r; 0]   (LAPACK zlartg convention)
__host__ __device__ __forceinline__ void givens(cplx a, cplx b, double& c, cplx& s) {
    const double nb = cabs2(b);
    if (nb == 0.0) { c = 1.0; s = mkc(0.0, 0.0); return; }
    const double na2 = cabs2(a);
    if (na2 == 0.0) {
        const double ab = sqrt(nb);
        c = 0.0; s = mkc(b.x / ab, -b.y / ab);   // conj(b)/|b|
        return;
    }
    // c = |a|/nrm, s = a conj(b) / (|a| nrm): two reciprocal square roots, no division
#ifdef __CUDA_ARCH__
    const double ra = rsqrt(na2), rn = rsqrt(na2 + nb);
#else
    const double ra = 1.0 / sqrt(na2), rn = 1.0 / sqrt(na2 + nb);
#endif
    c = na2 * ra * rn;
    const double f = ra * rn;
    const cplx ab = cmul(a, cconj(b));     // a * conj(b)
    s = mkc(ab.x * f, ab.y * f);
}

// ---------------------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ cplx warp_sum(cplx v) {
    v.x = warp_sum(v.x); v.y = warp_sum(v.y);
    return v;
}
// block-wide sum; scratch must hold >= 32 doubles; result broadcast to all threads. Ends with __syncthreads().
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = (lane < nw) ? scratch[lane] : 0.0;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ double block_max(double v, double* scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = (lane < nw) ? scratch[lane] : 0.0;
    r = warp_max(r);
    return r;
}

// ---------------------------------------------------------------------------------------------
// cp.async (LDGSTS) 16-byte staging with zero fill
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------
// DMMA: D(8x8) += A(8x4,row) * B(4x8,col), FP64.  Fragment layout (PTX ISA m8n8k4 .f64):
//   g = lane>>2, t = lane&3:  a = A[g][t],  b = B[t][g],  d0,d1 = D[g][2t], D[g][2t+1]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}

// Warp-level complex tile product: acc(8*MT x 8*NT) += opA(A)(8*MT x K) * opB(B)(K x 8*NT), K multiple of 4.
// A element (i,k) at A[i*a_si + k*a_sk]; B element (k,j) at B[k*b_sk + j*b_sj]  (shared memory, complex128).
// acc[i][j][0..1] = Re of (row g, cols 2t, 2t+1) of tile (i,j); acc[i][j][2..3] = Im.
template <int MT, int NT, bool CONJA, bool CONJB>
__device__ __forceinline__ void zmma_load_frags(cplx (&a)[MT], cplx (&b)[NT], const cplx* __restrict__ ap, int a_si, int a_sk,
                                                const cplx* __restrict__ bp, int b_sk, int b_sj, int k) {
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        a[i] = ap[(8 * i) * a_si + k * a_sk];
        if (CONJA) a[i].y = -a[i].y;
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        b[j] = bp[k * b_sk + (8 * j) * b_sj];
        if (CONJB) b[j].y = -b[j].y;
    }
}

// 4 real DMMAs per complex tile product, issued as two passes over all tiles so that the two DMMAs that hit the same
// accumulator pair are 2*MT*NT instructions apart (the dependent-issue latency of DMMA is long).
// real B operand (imaginary parts known to be zero): two DMMAs per tile product instead of four
template <int MT, int NT>
__device__ __forceinline__ void zmma_compute_breal(double (&acc)[MT][NT][4], const cplx (&a)[MT], const cplx (&b)[NT]) {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            dmma(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
            dmma(acc[i][j][2], acc[i][j][3], a[i].y, b[j].x);
        }
}

template <int MT, int NT>
__device__ __forceinline__ void zmma_compute(double (&acc)[MT][NT][4], const cplx (&a)[MT], const cplx (&b)[NT]) {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            dmma(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
            dmma(acc[i][j][2], acc[i][j][3], a[i].x, b[j].y);
        }
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const double nai = -a[i].y;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            dmma(acc[i][j][0], acc[i][j][1], nai, b[j].y);
            dmma(acc[i][j][2], acc[i][j][3], a[i].y, b[j].x);
        }
    }
}

template <int MT, int NT, bool CONJA, bool CONJB, bool BREAL = false>
__device__ __forceinline__ void warp_zmma(double (&acc)[MT][NT][4], const cplx* __restrict__ A, int a_si, int a_sk,
                                          const cplx* __restrict__ B, int b_sk, int b_sj, int kcount) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const cplx* ap = A + g * a_si + t * a_sk;
    const cplx* bp = B + t * b_sk + g * b_sj;
    if (kcount <= 0) return;
    // fragments are double-buffered in registers: the LDS.128 of step k+4 are in flight while step k's DMMAs issue
    cplx a0[MT], b0[NT], a1[MT], b1[NT];
    zmma_load_frags<MT, NT, CONJA, CONJB>(a0, b0, ap, a_si, a_sk, bp, b_sk, b_sj, 0);
    int k = 0;
    while (true) {
        if (k + 4 < kcount) zmma_load_frags<MT, NT, CONJA, CONJB>(a1, b1, ap, a_si, a_sk, bp, b_sk, b_sj, k + 4);
        if (BREAL) zmma_compute_breal<MT, NT>(acc, a0, b0); else zmma_compute<MT, NT>(acc, a0, b0);
        k += 4;
        if (k >= kcount) break;
        if (k + 4 < kcount) zmma_load_frags<MT, NT, CONJA, CONJB>(a0, b0, ap, a_si, a_sk, bp, b_sk, b_sj, k + 4);
        if (BREAL) zmma_compute_breal<MT, NT>(acc, a1, b1); else zmma_compute<MT, NT>(acc, a1, b1);
        k += 4;
        if (k >= kcount) break;
    }
}

template <int MT, int NT>
__device__ __forceinline__ void zero_acc(double (&acc)[MT][NT][4]) {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// 1-D bulk TMA (cp.async.bulk, SASS UBLKCP) global -> shared with an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(b), "r"(count));
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
                 "l"(gsrc), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t ok = 0;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(b), "r"(parity)
            : "memory");
    } while (!ok);
}
