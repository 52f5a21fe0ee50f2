// C-ABI of libllck.so: orchestration of the batched KBDM solve (see include/llck.h).
// Pipeline per ensemble member (reference llckbdm/kbdm.py:19-92, 133-240):
//   X <- Hankel(U^{p-1}) = Q B P^H (blocked bidiagonalisation);  B = L_b S R_b^T by divide and conquer   (kbdm.py:166)
//   Rs = R_l g^{-1/2},  Lt = L_l g^{-1/2}   (g = s or s + q^2/s)                (kbdm.py:171-186)
//   T1 = U^p Rs (implicit Hankel GEMM);  Ured = Lt^H T1                         (kbdm.py:189)
//   Ured = Q H Q^H -> Z T Z^H (Hessenberg + multishift QR);  Xev = trevc(T)     (kbdm.py:192)
//   P = Z Xev;  B = Rs P                                                        (kbdm.py:198)
//   W = U0 B (implicit Hankel GEMM);  N_k = sum_i B_ik W_ik;  D_k = W_0k^2/N_k  (kbdm.py:215-240, 71-75)
//   (A, T2, F, PH) from D_k and mu_k = T_kk                                      (kbdm.py:78-92)
#include "../../include/llck.h"
#include "common.cuh"
#include "gemm.cuh"
#include "eig.cuh"
#include "bidiag.cuh"
#include "svd_real.cuh"
#include "bdc.cuh"
#include "rmse.cuh"
#include "silhouette.cuh"
#include "features.cuh"
#include "hdbscan_mst.cuh"
#include "hdbscan_tree.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <mutex>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return -(int)e_; } while (0)

// ---- epilogue: one warp per eigen-pair ------------------------------------------------------------
__global__ void __launch_bounds__(256) epilogue_kernel(const cplx* Bm, const cplx* Wm, const cplx* Tm, long long stride, int ld,
                                                       const int* mv, const int* lv, double dwell,
                                                       double* line_lists, long long ll_stride, cplx* mu_out, cplx* d_out,
                                                       long long mu_stride, int* n_valid, int* status) {
    const int b = blockIdx.y;
    const int m = mv[b], l = lv[b];
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= l) return;
    const cplx* bc = Bm + (long long)b * stride + (long long)ld * k;
    const cplx* wc = Wm + (long long)b * stride + (long long)ld * k;
    cplx nk = mkc(0.0, 0.0);
    for (int i = lane; i < m; i += 32) nk = cfma(bc[i], wc[i], nk);   // bilinear, NO conjugate (kbdm.py:232)
    nk = warp_sum(nk);
    if (lane == 0) {
        const cplx w0 = wc[0];
        const cplx D = cdiv(cmul(w0, w0), nk);
        const cplx mu = Tm[(long long)b * stride + k + (long long)ld * k];
        const double A = hypot(D.x, D.y);
        const double PH = atan2(D.y, D.x);
        const double arg = atan2(mu.y, mu.x);
        const double lnabs = log(hypot(mu.x, mu.y));
        const double F = arg / (dwell * 6.283185307179586476925286766559);
        const double T2 = -dwell / lnabs;
        double* row = line_lists + (long long)b * ll_stride + 4 * k;
        row[0] = A; row[1] = T2; row[2] = F; row[3] = PH;
        if (mu_out) mu_out[(long long)b * mu_stride + k] = mu;
        if (d_out) d_out[(long long)b * mu_stride + k] = D;
        if (A > 1e-6 && T2 > 0.0) atomicAdd(&n_valid[b], 1);
        if (!isfinite(A) || !isfinite(mu.x) || !isfinite(mu.y)) atomicMax(&status[b], LLCK_STATUS_NONFINITE);
    }
}

__global__ void mark_unconverged_kernel(const int* done, int* status, int batch) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch && !done[b]) atomicMax(&status[b], LLCK_STATUS_SVD_NOCONV);
}

// members flagged by the divide-and-conquer SVD are the only ones the Jacobi path still has to solve
__global__ void fallback_to_done_kernel(int* fallback, const int* imbalance, int* done, int batch) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const int fb = (fallback[b] | imbalance[b]) ? 1 : 0;      // rank-deficiency test of bdc_sv_kernel | u/v imbalance seen by bdc_gather_kernel
    fallback[b] = fb;
    done[b] = fb ? 0 : 1;
}

// loop condition of the Jacobi sweeps (CUDA graph WHILE node): another sweep iff some member is not converged and sweeps are left.
// First node of the graph (decrement = 0) and last node of every sweep (decrement = 1).
__global__ void jacobi_cond_kernel(cudaGraphConditionalHandle handle, const int* done, int batch, int* sweeps_left, int decrement) {
    __shared__ int any;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    int mine = 0;
    for (int b = threadIdx.x; b < batch; b += blockDim.x) mine |= !done[b];
    if (mine) any = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        int left = *sweeps_left - decrement;
        *sweeps_left = left;
        cudaGraphSetConditional(handle, (any && left > 0) ? 1u : 0u);
    }
}

// per-call device state: convergence flags, outputs that are accumulated with atomics
__global__ void meta_zero_kernel(int* done, unsigned long long* sweep_off, int* status, int* n_valid, int* hqr_sweeps, int* imbalance, int batch) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    done[b] = 0; sweep_off[b] = 0ull; status[b] = 0; n_valid[b] = 0; hqr_sweeps[b] = 0; imbalance[b] = 0;
}

// ---- workspace layout -----------------------------------------------------------------------------
struct WsLayout {
    size_t mv, lv, nbv, done, scalars, imbalance, hqr_sweeps, perm, sig_off, sweep_off, vp, yp, vtp, wp, tws, pan6, pan7, dws, ews, jws, gws, offws, skip, bdcvec, fallback, ypart, mats, total;
    int nmats;
};
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
static WsLayout ws_layout(int batch, int ld, int flags) {
    WsLayout L;
    size_t o = 0;
    L.mv = o; o = al256(o + sizeof(int) * batch);
    L.lv = o; o = al256(o + sizeof(int) * batch);
    L.nbv = o; o = al256(o + sizeof(int) * batch);
    L.done = o; o = al256(o + sizeof(int) * batch);
    L.scalars = o; o = al256(o + 256);
    L.imbalance = o; o = al256(o + sizeof(int) * batch);
    L.hqr_sweeps = o; o = al256(o + sizeof(int) * batch);
    L.perm = o; o = al256(o + sizeof(int) * (size_t)batch * ld);
    L.sig_off = o; o = al256(o + sizeof(long long) * batch);
    L.sweep_off = o; o = al256(o + sizeof(unsigned long long) * batch);
    const size_t pan = sizeof(cplx) * (size_t)batch * ld * HB_NB;
    L.vp = o; o = al256(o + pan);
    L.yp = o; o = al256(o + pan);
    L.vtp = o; o = al256(o + pan);
    L.wp = o; o = al256(o + pan);
    L.tws = o; o = al256(o + pan);
    L.pan6 = o; o = al256(o + pan);
    L.pan7 = o; o = al256(o + pan);
    L.dws = o; o = al256(o + sizeof(double) * (size_t)batch * ld);
    L.ews = o; o = al256(o + sizeof(double) * (size_t)batch * ld);
    const int pairs_max = ((ld / J_B) + 1) / 2 + 1;
    L.jws = o; o = al256(o + sizeof(cplx) * (size_t)batch * pairs_max * 4096);
    L.skip = o; o = al256(o + sizeof(int) * (size_t)batch * pairs_max);
    L.gws = o; o = al256(o + sizeof(cplx) * (size_t)batch * pairs_max * 2080);
    L.offws = o; o = al256(o + sizeof(double) * (size_t)batch * pairs_max);
    L.bdcvec = o; o = al256(o + bdc_vec_bytes(batch, 2 * ld));
    L.fallback = o; o = al256(o + sizeof(int) * batch);
    // partial gemv results of the cluster-cooperative panel kernels (small batches only: [member][parity][rank <= 16][ld])
    L.ypart = o; o = al256(o + sizeof(cplx) * (size_t)(batch <= 74 ? batch : 0) * 2 * 16 * ld);
    L.mats = o;
    // 6 pipeline matrices (14 in debug mode) + 5 for the divide-and-conquer SVD of the bidiagonal (two 2ld x 2ld real
    // eigenvector buffers and the secular eigenvector matrices)
    L.nmats = ((flags & LLCK_FLAG_DEBUG_KEEP) ? 14 : 6) + 5;
    o += (size_t)L.nmats * batch * ld * ld * sizeof(cplx);
    L.total = o;
    return L;
}


// Portable cluster size only: 16-CTA clusters were measured bimodal across processes (hqr of one m=1024 member 187 ms or ~500 ms
// depending on where the hardware places the cluster), 8-CTA clusters repeatable (220 ms).
#define LLCK_MAX_CLUSTER 8

// ---- thread-block cluster launches for small batches ----------------------------------------------------------------
// largest power-of-two cluster size (<= 16) such that all `batch` clusters of `kernel` are co-resident in one wave
template <typename K>
static int pick_cluster_size(K kernel, int batch, int threads, size_t smem, int max_size, int forced) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
    int csize = 1;
    while (csize < max_size && 2 * csize * batch <= sms) csize *= 2;
    if (forced > 0) csize = forced < max_size ? forced : max_size;      // llck_options.cluster_size
    if (csize > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { (void)cudaGetLastError(); csize = 8; }
    while (csize > 1) {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.gridDim = dim3(batch * csize);
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int nclusters = 0;
        cudaError_t eo = cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg);
        // all clusters co-resident in ONE wave, with 1/8 head-room: at exactly full occupancy (74 pairs on 148 SMs) a second wave was observed
        if (eo == cudaSuccess && nclusters >= batch + (batch + 7) / 8) break;
        (void)cudaGetLastError();
        csize /= 2;
    }
    return csize;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_clustered(void (*kernel)(KArgs...), int batch, int csize, int threads, size_t smem, cudaStream_t st, Args... args) {
    LLCK_LAUNCHED();
    if (csize <= 1) {
        kernel<<<batch, threads, smem, st>>>(args...);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.gridDim = dim3(batch * csize); cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}


// The level-2 panel kernels are HBM-bound with serial vector phases in between the streaming passes: when the batch is larger than
// one wave (one CTA per SM), two members per SM keep the memory system busy through those phases (needs 2 x smem <= the SM's 227 KB).
static bool panels_two_per_sm(int batch, size_t smem) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
    return batch > sms && 2 * (smem + 1024) <= 227 * 1024;
}

// ---- blocked bidiagonalisation driver (shared by llck_kbdm_batched and the stage test entry) ---------------------
// A (in: matrices, out: reflectors + d/e), Q and P (out, ld x m each), panel buffers Vp/Yp/Xp/Up (ld x 32 each), Wp (32 x ld),
// TQws/TPws ((ld/32) x 32 x 32 each), dws/ews (ld doubles each) -- all per member with the given strides.
static int bidiag_driver(cplx* A, cplx* Q, cplx* P, long long stride, int ld, const int* d_mv, int mmax, int batch,
                         cplx* VX, cplx* YU, cplx* Wp, long long pstride, cplx* TQws, cplx* TPws,
                         double* dws, double* ews, cudaStream_t st, cplx* ypart = nullptr, int forced_cluster = 0) {
    // VX = per member [V | X] (ld x 64), YU = per member [Y | U] (ld x 64): the panel's trailing update
    //   A[e:, e:] -= V Y^H + X U^H  is ONE rank-64 GEMM  A[e:, e:] -= [V X] [Y U]^H
    const long long pstride2 = 2 * pstride;
    cplx* Vp = VX; cplx* Xp = VX + pstride; cplx* Yp = YU; cplx* Up = YU + pstride;
    size_t sm = (size_t)(2 * ld + 2 * BD_NB * BD_NB + 8 * BD_NB + 8) * 16 + 512;
    CK(cudaFuncSetAttribute(bidiag_panel_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    CK(cudaFuncSetAttribute(bidiag_panel_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const int bp_csize = (ypart != nullptr && batch <= 74) ? pick_cluster_size(bidiag_panel_kernel<1>, batch, E_THREADS, sm, LLCK_MAX_CLUSTER, forced_cluster) : 1;
    const bool two_per_sm = panels_two_per_sm(batch, sm);
    int k0_last = 0;
    for (int k0 = 0; k0 < mmax; k0 += BD_NB) {
        k0_last = k0;
        CK(launch_clustered(two_per_sm ? bidiag_panel_kernel<2> : bidiag_panel_kernel<1>, batch, bp_csize, E_THREADS, sm, st, A, stride, ld, d_mv, k0,
                            Vp, Yp, Xp, Up, pstride2, TQws, TPws, pstride, dws, ews, ypart, bp_csize));
        const int e = k0 + BD_NB;
        if (e >= mmax) continue;
        GemmParams g = gemm_params_zero();
        g.A = VX + e; g.strideA = pstride2; g.lda = ld;
        g.B = YU + e; g.strideB = pstride2; g.ldb = ld;
        g.C = A + e + (long long)ld * e; g.strideC = stride; g.ldc = ld;
        g.Mv = d_mv; g.Mc = -e; g.Nv = d_mv; g.Nc = -e; g.Kc = 2 * BD_NB; g.accum = 1;
        CK(zgemm_batched(A_NORMAL, g, mmax - e, mmax - e, 2 * BD_NB, batch, st, true));
    }
    dim3 gi(128, batch);
    set_identity_kernel<<<gi, 256, 0, st>>>(Q, stride, ld, d_mv);
    set_identity_kernel<<<gi, 256, 0, st>>>(P, stride, ld, d_mv);
    llck_launch_count += 2;
    CK(cudaGetLastError());
    for (int right = 0; right < 2; ++right) {
        cplx* Acc = right ? P : Q;
        const cplx* Tws = right ? TPws : TQws;
        for (int k0 = k0_last; k0 >= 0; k0 -= BD_NB) {
            const int o = k0 + right;
            if (o >= mmax) continue;
            bidiag_qpanel_kernel<<<batch, E_THREADS, 0, st>>>(A, stride, ld, d_mv, k0, right, Vp, Yp /*VTh*/, pstride2, Tws, pstride);
            LLCK_LAUNCHED();
            CK(cudaGetLastError());
            GemmParams g = gemm_params_zero();          // W = (V T^H)[o:, :]^H * Acc[o:, o:]
            g.A = Yp + o; g.strideA = pstride2; g.lda = ld;
            g.B = Acc + o + (long long)ld * o; g.strideB = stride; g.ldb = ld;
            g.C = Wp; g.strideC = pstride; g.ldc = BD_NB;
            g.Mc = BD_NB; g.Nv = d_mv; g.Nc = -o; g.Kv = d_mv; g.Kc = -o;
            CK(zgemm_batched(A_CONJT, g, BD_NB, mmax - o, mmax - o, batch, st));
            g = gemm_params_zero();                     // Acc[o:, o:] -= V[o:, :] * W
            g.A = Vp + o; g.strideA = pstride2; g.lda = ld;
            g.B = Wp; g.strideB = pstride; g.ldb = BD_NB;
            g.C = Acc + o + (long long)ld * o; g.strideC = stride; g.ldc = ld;
            g.Mv = d_mv; g.Mc = -o; g.Nv = d_mv; g.Nc = -o; g.Kc = BD_NB; g.accum = 1;
            CK(zgemm_batched(A_NORMAL, g, mmax - o, mmax - o, BD_NB, batch, st));
        }
    }
    return 0;
}

// Executable graphs cannot be destroyed while in flight without blocking the caller (measured: cudaGraphExecDestroy right after
// cudaGraphLaunch waits for the graph), so a launched graph is parked here with an event recorded behind it and released by a
// later call -- or by llck_release_resources() -- once that event has completed.  This is the library's only process-wide state:
// host-side handles awaiting release, never read by the computation.
struct PendingGraph { cudaGraphExec_t exec; cudaGraph_t graph; cudaEvent_t done; };
static std::mutex g_pending_mu;
static std::vector<PendingGraph> g_pending;

static void reap_pending_graphs(bool wait) {
    std::lock_guard<std::mutex> lock(g_pending_mu);
    size_t keep = 0;
    for (size_t i = 0; i < g_pending.size(); ++i) {
        PendingGraph& pg = g_pending[i];
        cudaError_t q = wait ? cudaEventSynchronize(pg.done) : cudaEventQuery(pg.done);
        if (q == cudaErrorNotReady) { g_pending[keep++] = pg; continue; }
        if (q != cudaSuccess) (void)cudaGetLastError();
        cudaGraphExecDestroy(pg.exec);
        cudaGraphDestroy(pg.graph);
        cudaEventDestroy(pg.done);
    }
    g_pending.resize(keep);
}

// One Jacobi sweep on `st` (all rounds + the convergence bookkeeping); CTAs of converged members exit at once.
static cudaError_t enqueue_jacobi_sweep(RJacobiParams rp, int nbmax, int batch, unsigned long long* d_swoff, int* d_done, double conv2, cudaStream_t st) {
    for (int r = 0; r < nbmax - 1; ++r) {
        rp.round = r;
        dim3 gA(nbmax / 2, batch), gA2(nbmax / 2, batch, 2);
        rjacobi_gram_kernel<<<gA, 256, RJ_GRAM_SMEM, st>>>(rp);
        rjacobi_eig_kernel<<<gA, RJE_THREADS, RJE_SMEM, st>>>(rp);
        rjacobi_update_kernel<<<gA2, 256, RJ_UPD_SMEM, st>>>(rp);
    }
    jacobi_sweep_end_kernel<<<(batch + 127) / 128, 128, 0, st>>>(d_swoff, d_done, batch, conv2);
    return cudaGetLastError();
}

// The Jacobi iteration as ONE graph launch: a device-side WHILE node (CUDA graph conditional node) repeats the sweep while some
// member is unconverged and sweeps are left -- no host read-back, and nothing at all runs when the divide-and-conquer SVD solved
// every member.  Returns cudaErrorNotSupported-like errors to the caller, which then enqueues the sweeps unconditionally.
static cudaError_t launch_jacobi_while_graph(const RJacobiParams& rp, int nbmax, int batch, int max_sweeps, unsigned long long* d_swoff,
                                             int* d_done, double conv2, int* d_sweeps_left, cudaStream_t st, int* host_us) {
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return (int)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count(); };
    const auto t0 = now();
    auto t1 = t0, t2 = t0, t3 = t0;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaStream_t cap = nullptr;
    cudaError_t e = cudaSuccess;
    bool capturing = false;
    do {
        if ((e = cudaMemcpyAsync(d_sweeps_left, &max_sweeps, sizeof(int), cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        if ((e = cudaGraphCreate(&graph, 0)) != cudaSuccess) break;
        cudaGraphConditionalHandle handle;
        if ((e = cudaGraphConditionalHandleCreate(&handle, graph, 0, 0)) != cudaSuccess) break;
        cudaGraphNode_t first = nullptr, loop = nullptr;
        {
            int dec = 0;
            const int* done_c = d_done;
            void* args[] = {&handle, &done_c, &batch, &d_sweeps_left, &dec};
            cudaKernelNodeParams kp = {};
            kp.func = (void*)jacobi_cond_kernel; kp.gridDim = dim3(1); kp.blockDim = dim3(256); kp.sharedMemBytes = 0; kp.kernelParams = args;
            if ((e = cudaGraphAddKernelNode(&first, graph, nullptr, 0, &kp)) != cudaSuccess) break;
        }
        cudaGraphNodeParams cp = {};
        cp.type = cudaGraphNodeTypeConditional;
        cp.conditional.handle = handle;
        cp.conditional.type = cudaGraphCondTypeWhile;
        cp.conditional.size = 1;
        if ((e = cudaGraphAddNode(&loop, graph, &first, 1, &cp)) != cudaSuccess) break;
        cudaGraph_t body = cp.conditional.phGraph_out[0];
        if ((e = cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((e = cudaStreamBeginCaptureToGraph(cap, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed)) != cudaSuccess) break;
        capturing = true;
        e = enqueue_jacobi_sweep(rp, nbmax, batch, d_swoff, d_done, conv2, cap);
        if (e == cudaSuccess) {
            jacobi_cond_kernel<<<1, 256, 0, cap>>>(handle, d_done, batch, d_sweeps_left, 1);
            e = cudaGetLastError();
        }
        cudaGraph_t ended = nullptr;
        cudaError_t e2 = cudaStreamEndCapture(cap, &ended);
        capturing = false;
        if (e == cudaSuccess) e = e2;
        if (e != cudaSuccess) break;
        t1 = now();
        if ((e = cudaGraphInstantiate(&exec, graph, 0)) != cudaSuccess) break;
        t2 = now();
        e = cudaGraphLaunch(exec, st);
        t3 = now();
    } while (0);
    if (capturing) { cudaGraph_t ended = nullptr; cudaStreamEndCapture(cap, &ended); }
    if (cap) cudaStreamDestroy(cap);             // nothing ever ran on the capture stream
    bool parked = false;
    if (e == cudaSuccess && exec) {
        cudaEvent_t done = nullptr;
        if (cudaEventCreateWithFlags(&done, cudaEventDisableTiming) == cudaSuccess) {
            if (cudaEventRecord(done, st) == cudaSuccess) {
                std::lock_guard<std::mutex> lock(g_pending_mu);
                g_pending.push_back({exec, graph, done});
                parked = true;
            } else {
                cudaEventDestroy(done);
            }
        }
    }
    if (!parked) {                               // error path (or no event): blocking release
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
    }
    if (e != cudaSuccess) (void)cudaGetLastError();
    if (host_us) { const auto t4 = now(); host_us[0] = us(t0, t1); host_us[1] = us(t1, t2); host_us[2] = us(t2, t3); host_us[3] = us(t3, t4); }
    return e;
}

extern "C" {

int llck_version(void) { return LLCK_VERSION; }

int llck_release_resources(void) {
    reap_pending_graphs(true);
    return 0;
}

int llck_leading_dim(int m_max) { return ((m_max + 63) / 64) * 64; }

size_t llck_workspace_bytes(int batch, int ld, int flags) {
    if (batch <= 0 || ld <= 0) return 0;
    return ws_layout(batch, ld, flags).total;
}

size_t llck_debug_offset(int batch, int ld, int which) {
    WsLayout L = ws_layout(batch, ld, LLCK_FLAG_DEBUG_KEEP);
    return L.mats + (size_t)which * batch * ld * ld * sizeof(cplx);
}

int llck_zgemm(int32_t amode, const void* A, int32_t lda, const void* B, int32_t ldb, void* C, int32_t ldc,
               int32_t M, int32_t N, int32_t K, const void* sig, int32_t shift, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int* dims = nullptr;
    long long* off = nullptr;
    CK(cudaMalloc(&dims, 3 * sizeof(int)));
    CK(cudaMalloc(&off, sizeof(long long)));
    int h[3] = {M, N, K};
    long long z = 0;
    CK(cudaMemcpyAsync(dims, h, sizeof(h), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(off, &z, sizeof(z), cudaMemcpyHostToDevice, st));
    GemmParams p = gemm_params_zero();
    p.A = (const cplx*)A; p.strideA = 0; p.lda = lda;
    p.B = (const cplx*)B; p.strideB = 0; p.ldb = ldb;
    p.C = (cplx*)C; p.strideC = 0; p.ldc = ldc;
    p.Mv = dims; p.Nv = dims + 1; p.Kv = dims + 2;
    p.sig = (const cplx*)sig; p.sig_off = off; p.shift = shift;
    cudaError_t e = zgemm_batched(amode, p, M, N, K, 1, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(dims); cudaFree(off);
    if (e != cudaSuccess) return -(int)e;
    if (e2 != cudaSuccess) return -(int)e2;
    return 0;
}


int llck_bidiag_test(void* A, int32_t m, int32_t ld, double* d_out, double* e_out, void* Q, void* P, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const size_t pan = sizeof(cplx) * (size_t)ld * BD_NB;
    unsigned char* w = nullptr;
    CK(cudaMalloc(&w, 7 * pan + 256));
    CK(cudaMemsetAsync(w, 0, 7 * pan + 256, st));
    cplx* VX = (cplx*)w; cplx* YU = (cplx*)(w + 2 * pan);
    cplx* Wp = (cplx*)(w + 4 * pan); cplx* TQ = (cplx*)(w + 5 * pan); cplx* TP = (cplx*)(w + 6 * pan);
    int* d_m = (int*)(w + 7 * pan);
    CK(cudaMemcpyAsync(d_m, &m, sizeof(int), cudaMemcpyHostToDevice, st));
    double* dd = nullptr;
    CK(cudaMalloc(&dd, sizeof(double) * 2 * ld));
    int rc = bidiag_driver((cplx*)A, (cplx*)Q, (cplx*)P, (long long)ld * ld, ld, d_m, m, 1, VX, YU, Wp, (long long)ld * BD_NB, TQ, TP, dd, dd + ld, st);
    if (rc == 0) {
        cudaMemcpyAsync(d_out, dd, sizeof(double) * m, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(e_out, dd + ld, sizeof(double) * (m > 1 ? m - 1 : 0), cudaMemcpyDeviceToDevice, st);
    }
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(w); cudaFree(dd);
    if (rc) return rc;
    return e2 == cudaSuccess ? 0 : -(int)e2;
}

int llck_rmse_batched(const void* data, int32_t N, double dwell, const double* line_lists, int64_t ll_stride,
                      const int32_t* n_rows, int32_t batch, int32_t filter, double amplitude_tol, double* rmse_out, void* stream) {
    if (!data || !line_lists || !n_rows || !rmse_out || N < 1 || batch < 1 || !(dwell > 0.0) || ll_stride < 4) return LLCK_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= RMSE_SMEM_POINTS) {                          // whole model FID staged in shared memory
        const size_t sm = sizeof(cplx) * (size_t)N;
        CK(cudaFuncSetAttribute(rmse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        rmse_kernel<<<batch, RMSE_THREADS, sm, st>>>((const cplx*)data, N, dwell, line_lists, ll_stride, n_rows, filter, amplitude_tol, rmse_out);
    } else {                                              // long FIDs: tiled over n
        const size_t sm = sizeof(cplx) * 2 * RMSE_TILE;
        CK(cudaFuncSetAttribute(rmse_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        rmse_tiled_kernel<<<batch, RMSE_THREADS, sm, st>>>((const cplx*)data, N, dwell, line_lists, ll_stride, n_rows, filter, amplitude_tol, rmse_out);
    }
    CK(cudaGetLastError());
    return 0;
}

int llck_silhouette_batched(const double* X, int32_t n, const int32_t* order, const int32_t* seg, const int32_t* nseg,
                            const int32_t* cluster_of, int32_t nclusterings, double* out, void* stream) {
    if (!X || !order || !seg || !nseg || !cluster_of || !out || n < 1 || nclusterings < 1 || nclusterings > 65535) return LLCK_E_BADARG;
    dim3 grid((n + SIL_THREADS - 1) / SIL_THREADS, nclusterings);
    silhouette_kernel<<<grid, SIL_THREADS, 0, (cudaStream_t)stream>>>(X, n, order, seg, nseg, cluster_of, out);
    CK(cudaGetLastError());
    return 0;
}

int llck_pool_features(const double* line_lists, int64_t ll_stride, const int32_t* n_rows, const int64_t* offset, int32_t batch,
                       double dwell, double amplitude_tol, double* samples, double* features, void* stream) {
    if (!line_lists || !n_rows || !offset || !samples || !features || batch < 1 || ll_stride < 4 || !(dwell > 0.0)) return LLCK_E_BADARG;
    pool_features_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(line_lists, ll_stride, n_rows, (const long long*)offset, dwell, amplitude_tol,
                                                                   samples, features);
    CK(cudaGetLastError());
    return 0;
}

int llck_hdbscan_core_distances(const double* X, int32_t n, int32_t kmax, double* core, void* stream) {
    if (!X || !core || n < 1 || kmax < 1 || kmax > HDB_KMAX || kmax > n) return LLCK_E_BADARG;
    hdb_core_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(X, n, kmax, core);
    CK(cudaGetLastError());
    return 0;
}

int llck_hdbscan_mst(const double* X, int32_t n, const double* core, const int32_t* core_row, int32_t nfits,
                     double* min_reach, int32_t* cur_src, int64_t* mst_src, int64_t* mst_dst, double* mst_w, int32_t flags, void* stream) {
    if (!X || !core || !core_row || !min_reach || !cur_src || !mst_src || !mst_dst || !mst_w) return LLCK_E_BADARG;
    if (n < 2 || n > HDB_PRIM_THREADS * 128 || nfits < 1) return LLCK_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    // one thread-block cluster per fit with the points resident in shared memory / registers, when they fit
    const int pt = (n + HDB_CS * HDB_CT - 1) / (HDB_CS * HDB_CT);
    if (pt <= HDB_PT_MAX && !(flags & LLCK_MST_SINGLE_CTA)) {
        const size_t smem = (size_t)pt * HDB_CT * sizeof(double4) + 2 * HDB_CS * sizeof(HdbCand);
        auto kern = (flags & LLCK_MST_DIM3) ? hdb_prim_cluster_kernel<true> : hdb_prim_cluster_kernel<false>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute attr[1];
            cfg.blockDim = dim3(HDB_CT); cfg.dynamicSmemBytes = smem; cfg.gridDim = dim3(nfits * HDB_CS); cfg.stream = st;
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = HDB_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) == cudaSuccess && nclusters >= 1) {
                CK(cudaLaunchKernelEx(&cfg, kern, X, (int)n, core, core_row, pt, (long long*)mst_src, (long long*)mst_dst, mst_w));
                return 0;
            }
        }
        (void)cudaGetLastError();
    }
    hdb_prim_kernel<<<nfits, HDB_PRIM_THREADS, 0, st>>>(X, n, core, core_row, min_reach, cur_src,
                                                        (long long*)mst_src, (long long*)mst_dst, mst_w);
    CK(cudaGetLastError());
    return 0;
}

int llck_hdbscan_labels(const int64_t* mst_src, const int64_t* mst_dst, const double* mst_w, const int64_t* order, int32_t n, int32_t nfits,
                        int32_t min_cluster_size, int32_t nthreads, int32_t* labels) {
    if (!mst_src || !mst_dst || !mst_w || !labels || n < 1 || nfits < 1 || min_cluster_size < 2) return LLCK_E_BADARG;
    const int64_t ne = (int64_t)n - 1;
    for (int64_t i = 0; i < (int64_t)nfits * ne; ++i) {
        if (mst_src[i] < 0 || mst_src[i] >= n || mst_dst[i] < 0 || mst_dst[i] >= n) return LLCK_E_BADARG;
        if (order && (order[i] < 0 || order[i] >= ne)) return LLCK_E_BADARG;
    }
    try {
        llck_hdb::labels_all_fits(mst_src, mst_dst, mst_w, order, n, nfits, min_cluster_size, nthreads, labels);
    } catch (...) {
        return LLCK_E_BADARG;
    }
    return 0;
}

int llck_multi_fid_batched(const double* params, int64_t pstride, const int32_t* n_rows, int32_t batch, int32_t N, double dwell,
                           void* out, void* stream) {
    if (!params || !n_rows || !out || batch < 1 || batch > 65535 || N < 1 || pstride < 4) return LLCK_E_BADARG;
    dim3 grid((N + 255) / 256, batch);
    multi_fid_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(params, pstride, n_rows, N, dwell, (cplx*)out);
    CK(cudaGetLastError());
    return 0;
}

int llck_bdc_test(const double* d, const double* e, const int32_t* m, int32_t batch, int32_t ld,
                  double* sing_vals, void* Us, void* V, int32_t* fallback, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!d || !e || !m || batch <= 0 || ld <= 0 || (ld & 63)) return LLCK_E_BADARG;
    int mmax = 0;
    for (int b = 0; b < batch; ++b) { if (m[b] < 1 || m[b] > ld) return LLCK_E_BADARG; if (m[b] > mmax) mmax = m[b]; }
    const size_t qbytes = sizeof(double) * (size_t)batch * 4 * ld * ld;
    const size_t vbytes = bdc_vec_bytes(batch, 2 * ld);
    unsigned char* w = nullptr;
    CK(cudaMalloc(&w, 2 * qbytes + qbytes / 2 + vbytes + 3 * sizeof(int) * (size_t)batch + 1024));
    BdcParams bp;
    bp.dws = d; bp.ews = e; bp.ld = ld; bp.batch = batch; bp.ldq = 2 * ld;
    bp.qstride = 4LL * ld * ld; bp.xstride = 2LL * ld * ld;
    bp.Q[0] = (double*)w; bp.Q[1] = (double*)(w + qbytes); bp.X = (double*)(w + 2 * qbytes);
    bdc_carve_vectors(bp, w + 2 * qbytes + qbytes / 2, batch, 2 * ld);
    int* d_m = (int*)(w + 2 * qbytes + qbytes / 2 + vbytes);
    int* d_status = d_m + batch;
    int* d_imb = d_status + batch;
    bp.mv = d_m; bp.level = 0;
    cudaError_t e1 = cudaMemcpyAsync(d_m, m, sizeof(int) * batch, cudaMemcpyHostToDevice, st);
    if (e1 == cudaSuccess) e1 = cudaMemsetAsync(d_status, 0, 2 * sizeof(int) * batch, st);
    int rc = (e1 == cudaSuccess) ? bdc_driver(bp, mmax, st) : -(int)e1;
    if (rc == 0) {
        bdc_sv_kernel<<<batch, 256, 0, st>>>(bp, sing_vals, ld, fallback);
        dim3 grid(mmax, batch);
        bdc_gather_kernel<<<grid, 128, 0, st>>>(bp, d_m, sing_vals, ld, 0.0, (cplx*)Us, (cplx*)V, (long long)ld * ld, ld, d_status, fallback, d_imb, 1);
        cudaError_t e2 = cudaGetLastError();
        if (e2 != cudaSuccess) rc = -(int)e2;
    }
    cudaError_t e3 = cudaStreamSynchronize(st);
    cudaFree(w);
    if (rc) return rc;
    return e3 == cudaSuccess ? 0 : -(int)e3;
}

int llck_kbdm_batched(const void* signals, const int64_t* sig_offset, const int64_t* sig_len, const int32_t* m, const int32_t* l,
                      int32_t p, double q, double dwell, int32_t batch,
                      double* line_lists, int64_t ll_stride,
                      void* mu_out, void* d_out, int64_t mu_stride,
                      double* sing_vals, int64_t sv_stride,
                      int32_t* n_valid, int32_t* status,
                      void* workspace, size_t workspace_bytes, int32_t flags, const llck_options* opts,
                      void* stream, int32_t* info) {
    if (!signals || !sig_offset || !sig_len || !m || !l || !line_lists || !sing_vals || !n_valid || !status || !workspace) return LLCK_E_BADARG;
    if (batch <= 0 || p < 1 || q < 0.0 || !(dwell > 0.0)) return LLCK_E_BADARG;
    llck_options o;
    memset(&o, 0, sizeof(o));
    if (opts) {
        if (opts->struct_size < (int32_t)sizeof(int32_t) || opts->struct_size > (int32_t)sizeof(llck_options)) return LLCK_E_BADARG;
        memcpy(&o, opts, (size_t)opts->struct_size);
    }
    if (o.svd_mode != LLCK_SVD_DC && o.svd_mode != LLCK_SVD_JACOBI) return LLCK_E_BADARG;
    if (o.cluster_size != 0 && o.cluster_size != 1 && o.cluster_size != 2 && o.cluster_size != 4 && o.cluster_size != 8) return LLCK_E_BADARG;
    if (o.aed_window != 0 && (o.aed_window < 8 || o.aed_window > 48)) return LLCK_E_BADARG;
    if (o.aed_nibble < 0 || o.aed_nibble > 1000 || o.jacobi_max_sweeps < 0 || o.jacobi_max_sweeps > 60) return LLCK_E_BADARG;
    int mmax = 0, lmax = 0;
    for (int b = 0; b < batch; ++b) {
        if (m[b] < 1 || l[b] < 1 || l[b] > m[b]) return LLCK_E_BADARG;
        if (sig_offset[b] < 0 || sig_len[b] < 2 * (int64_t)m[b] + p - 1) return LLCK_E_SHORT_SIGNAL;    // kbdm.py:59-62: 2m + p - 1 <= N
        if (m[b] > mmax) mmax = m[b];
        if (l[b] > lmax) lmax = l[b];
        if (ll_stride < 4 * (int64_t)l[b] || sv_stride < m[b]) return LLCK_E_BADARG;
        if ((mu_out || d_out) && mu_stride < l[b]) return LLCK_E_BADARG;
    }
    const int ld = llck_leading_dim(mmax);
    if (ld > LLCK_M_MAX) return LLCK_E_TOO_LARGE;
    const WsLayout L = ws_layout(batch, ld, flags);
    if (workspace_bytes < L.total) return LLCK_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const bool timing = (flags & LLCK_FLAG_TIMING) != 0 && info != nullptr;
    // CUDA events of the timing mode and the staged host metadata are released on every exit path
    struct Guard {
        cudaEvent_t tev[10]; int ntev_created = 0; void* host = nullptr;
        ~Guard() { for (int i = 0; i < ntev_created; ++i) cudaEventDestroy(tev[i]); free(host); }
    } G;
    int ntev = 0;
    if (timing) for (int i = 0; i < 10; ++i) { CK(cudaEventCreate(&G.tev[i])); G.ntev_created = i + 1; }
#define TICK() do { if (timing) { CK(cudaEventRecord(G.tev[ntev++], st)); } } while (0)
    unsigned char* ws = (unsigned char*)workspace;
    int* d_mv = (int*)(ws + L.mv);
    int* d_lv = (int*)(ws + L.lv);
    int* d_nbv = (int*)(ws + L.nbv);
    int* d_done = (int*)(ws + L.done);
    int* d_hqrs = (int*)(ws + L.hqr_sweeps);
    int* d_perm = (int*)(ws + L.perm);
    long long* d_soff = (long long*)(ws + L.sig_off);
    unsigned long long* d_swoff = (unsigned long long*)(ws + L.sweep_off);
    cplx* mats = (cplx*)(ws + L.mats);
    const long long stride = (long long)ld * ld;
    const bool dbg = (flags & LLCK_FLAG_DEBUG_KEEP) != 0;
    auto mat = [&](int i) { return mats + (size_t)i * batch * stride; };
    // buffer assignment (production aliases dead buffers; debug keeps all 14)
    cplx *bX, *bRs, *bLt, *bT1, *bH, *bZ, *bXev, *bP, *bB, *bW;
    if (dbg) { bX = mat(0); bRs = mat(2); bLt = mat(3); bT1 = mat(4); bH = mat(8); bZ = mat(9); bXev = mat(10); bP = mat(11); bB = mat(12); bW = mat(13); }
    else     { bX = mat(0); bRs = mat(2); bLt = mat(3); bT1 = mat(0); bH = mat(1); bZ = mat(3); bXev = mat(0); bP = mat(4); bB = mat(5); bW = mat(0); }
    llck_launch_count = 0;
    reap_pending_graphs(false);         // release graphs of earlier calls whose launch has completed (non-blocking)
    int jacobi_graph = 0;
    int graph_us[4] = {0, 0, 0, 0};     // host microseconds: build + capture, instantiate, launch, release

    // ---- metadata: per-member sizes and FID offsets go to the device through ONE staged copy; the call never waits for the stream ----
    int nbmax = 2;
    {
        const size_t meta_bytes = L.sig_off + sizeof(long long) * (size_t)batch - L.mv;     // mv | lv | nbv | done | ... | sig_off are contiguous
        (void)meta_bytes;
        int* h = (int*)malloc(sizeof(int) * (size_t)batch);
        if (!h) return LLCK_E_BADARG;
        G.host = h;
        for (int b = 0; b < batch; ++b) {
            int nb = (m[b] + J_B - 1) / J_B;
            if (nb < 2) nb = 2;
            if (nb & 1) ++nb;
            h[b] = nb;
            if (nb > nbmax) nbmax = nb;
        }
        // pageable -> device cudaMemcpyAsync returns once the source has been staged: the host arrays may be reused right after the call
        CK(cudaMemcpyAsync(d_mv, m, sizeof(int) * batch, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_lv, l, sizeof(int) * batch, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_nbv, h, sizeof(int) * batch, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_soff, sig_offset, sizeof(long long) * batch, cudaMemcpyHostToDevice, st));
    }
    meta_zero_kernel<<<(batch + 255) / 256, 256, 0, st>>>(d_done, d_swoff, status, n_valid, d_hqrs, (int*)(ws + L.imbalance), batch);
    LLCK_LAUNCHED();
    CK(cudaGetLastError());

    // ---- SVD of U^{p-1}: bidiagonalisation, then divide and conquer on the real bidiagonal ----
    TICK();   // 0
    {
        dim3 grid(256, batch);
        hankel_init_kernel<<<grid, 256, 0, st>>>(bX, stride, ld, d_mv, d_nbv, (const cplx*)signals, d_soff, p - 1);
        LLCK_LAUNCHED();
        CK(cudaGetLastError());
    }
    const int max_sweeps = o.jacobi_max_sweeps > 0 ? o.jacobi_max_sweeps : 30;
    {
        // (1) U^{p-1} = Q B P^H, B real upper bidiagonal
        cplx* bQ = dbg ? mat(9) : mat(1);
        cplx* bPm = dbg ? mat(11) : mat(4);
        cplx* bReal = dbg ? mat(12) : mat(5);
        cplx* bLpre = dbg ? mat(6) : mat(2);
        cplx* bRpre = dbg ? mat(7) : mat(0);
        const long long pstride = (long long)ld * BD_NB;
        double* dws = (double*)(ws + L.dws); double* ews = (double*)(ws + L.ews);
        {
            // [V|X] lives in the (vp, yp) pair of panel buffers, [Y|U] in (vtp, wp): both pairs are contiguous in the workspace
            int rc = bidiag_driver(bX, bQ, bPm, stride, ld, d_mv, mmax, batch, (cplx*)(ws + L.vp), (cplx*)(ws + L.vtp),
                                   (cplx*)(ws + L.pan6), pstride, (cplx*)(ws + L.tws), (cplx*)(ws + L.pan7), dws, ews, st,
                                   (cplx*)(ws + L.ypart), o.cluster_size);
            if (rc) return rc;
        }
        TICK();   // 1: init + bidiagonalisation done
        // (2) SVD of the real bidiagonal B: divide and conquer.  Members it flags as numerically rank deficient (and every member
        //     with LLCK_SVD_JACOBI) are solved by the real block one-sided Jacobi.  The decision stays on the device: the Jacobi
        //     rounds are always enqueued and their CTAs exit at once for members that are done (no host read-back).
        const bool use_dc = (o.svd_mode == LLCK_SVD_DC);
        int* d_fallback = (int*)(ws + L.fallback);
        BdcParams bp;
        if (use_dc) {
            const int nb0 = dbg ? 14 : 6;
            bp.dws = dws; bp.ews = ews; bp.ld = ld; bp.mv = d_mv; bp.batch = batch;
            bp.ldq = 2 * ld; bp.qstride = 4 * stride; bp.xstride = 2 * stride;
            bp.Q[0] = (double*)mat(nb0); bp.Q[1] = (double*)mat(nb0 + 2); bp.X = (double*)mat(nb0 + 4);
            bdc_carve_vectors(bp, ws + L.bdcvec, batch, 2 * ld);
            bp.level = 0;
            int rc = bdc_driver(bp, mmax, st);
            if (rc) return rc;
            bdc_sv_kernel<<<batch, 256, 0, st>>>(bp, sing_vals, sv_stride, d_fallback);
            LLCK_LAUNCHED();
            dim3 grid(lmax, batch);
            bdc_gather_kernel<<<grid, 128, 0, st>>>(bp, d_lv, sing_vals, sv_stride, q, bLpre, bRpre, stride, ld, status, d_fallback, (int*)(ws + L.imbalance), 0);
            LLCK_LAUNCHED();
            fallback_to_done_kernel<<<(batch + 127) / 128, 128, 0, st>>>(d_fallback, (const int*)(ws + L.imbalance), d_done, batch);
            LLCK_LAUNCHED();
            CK(cudaGetLastError());
        }
        const int* d_only = use_dc ? d_fallback : nullptr;
        double* Xr = (double*)bReal;
        double* Vr = Xr + (long long)ld * ld;
        const long long rstride = 2 * stride;       // doubles per member
        {
            dim3 grid(256, batch);
            rsvd_init_kernel<<<grid, 256, 0, st>>>(Xr, Vr, rstride, ld, d_mv, d_nbv, dws, ews, d_only);
            LLCK_LAUNCHED();
            CK(cudaGetLastError());
        }
        CK(cudaFuncSetAttribute(rjacobi_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RJ_GRAM_SMEM));
        CK(cudaFuncSetAttribute(rjacobi_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RJE_SMEM));
        CK(cudaFuncSetAttribute(rjacobi_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RJ_UPD_SMEM));
        RJacobiParams rp;
        rp.X = Xr; rp.V = Vr; rp.stride = rstride; rp.ld = ld; rp.mv = d_mv; rp.nbv = d_nbv; rp.sweep_off = d_swoff; rp.done = d_done;
        rp.tol2 = 1e-28; rp.inner_sweeps = 1;
        rp.Jws = (double*)(ws + L.jws); rp.Gws = (double*)(ws + L.gws); rp.offws = (double*)(ws + L.offws); rp.skip = (int*)(ws + L.skip);
        rp.pairs_max = ((ld / J_B) + 1) / 2 + 1;
        // a member is converged when no pair exceeded 1e-6 (scaled) during a sweep: that sweep leaves <= ~1e-12
        const double conv = o.jacobi_conv > 0.0 ? o.jacobi_conv : 1e-6;
        // device-side loop: one graph launch whose WHILE node repeats the sweep until every member has converged (or max_sweeps)
        int* d_sweeps_left = (int*)(ws + L.scalars);
        bool looped = false;
        if (!(flags & LLCK_FLAG_NO_GRAPH)) {
            looped = launch_jacobi_while_graph(rp, nbmax, batch, max_sweeps, d_swoff, d_done, conv * conv, d_sweeps_left, st, graph_us) == cudaSuccess;
            if (looped) { LLCK_LAUNCHED(); jacobi_graph = 1; }
        }
        if (!looped) {
            // without conditional graph nodes: every sweep is enqueued; the CTAs of converged members exit at once
            for (int sweep = 0; sweep < max_sweeps; ++sweep) {
                CK(enqueue_jacobi_sweep(rp, nbmax, batch, d_swoff, d_done, conv * conv, st));
                llck_launch_count += 3 * (nbmax - 1) + 1;
            }
        }
        mark_unconverged_kernel<<<(batch + 127) / 128, 128, 0, st>>>(d_done, status, batch);
        LLCK_LAUNCHED();
        // singular values, truncation/scaling of the Jacobi members
        int npow2 = 64;
        while (npow2 < ld) npow2 <<= 1;
        rsvd_finalize_kernel<<<batch, 256, npow2 * 12, st>>>(Xr, rstride, ld, d_mv, d_nbv, sing_vals, sv_stride, d_perm, npow2, d_only);
        LLCK_LAUNCHED();
        {
            dim3 grid(lmax, batch);
            rsvd_gather_kernel<<<grid, 128, 0, st>>>(Xr, Vr, rstride, ld, d_mv, d_lv, sing_vals, sv_stride, d_perm, q, bLpre, bRpre, stride, status, 0, d_only);
            LLCK_LAUNCHED();
            CK(cudaGetLastError());
        }
        TICK();   // 2: SVD of the bidiagonal done
        // (3) back-multiplication by Q and P
        GemmParams g = gemm_params_zero();          // Lt = Q * Lpre
        g.A = bQ; g.strideA = stride; g.lda = ld; g.B = bLpre; g.strideB = stride; g.ldb = ld; g.C = bLt; g.strideC = stride; g.ldc = ld;
        g.Mv = d_mv; g.Nv = d_lv; g.Kv = d_mv;
        CK(zgemm_batched(A_NORMAL, g, mmax, lmax, mmax, batch, st, false, true));        // Lpre / Rpre are real
        g.A = bPm; g.B = bRpre; g.C = bRs;            // Rs = P * Rpre
        CK(zgemm_batched(A_NORMAL, g, mmax, lmax, mmax, batch, st, false, true));
        if (dbg) {      // X_dbg = Q * X[:,perm] (= L Sigma), V_dbg = P * V[:,perm] (= R) for the stage checker
            CK(cudaMemsetAsync(mat(0), 0, sizeof(cplx) * batch * stride, st));
            CK(cudaMemsetAsync(mat(1), 0, sizeof(cplx) * batch * stride, st));
            dim3 grid(mmax, batch);
            rsvd_gather_kernel<<<grid, 128, 0, st>>>(Xr, Vr, rstride, ld, d_mv, d_lv, sing_vals, sv_stride, d_perm, q, bLpre, bRpre, stride, status, 1, d_only);
            if (use_dc) bdc_gather_kernel<<<grid, 128, 0, st>>>(bp, d_lv, sing_vals, sv_stride, q, bLpre, bRpre, stride, ld, status, d_fallback, (int*)(ws + L.imbalance), 1);
            CK(cudaGetLastError());
            g.A = bQ; g.B = bLpre; g.C = mat(0); g.Nv = d_mv;
            CK(zgemm_batched(A_NORMAL, g, mmax, mmax, mmax, batch, st));
            g.A = bPm; g.B = bRpre; g.C = mat(1);
            CK(zgemm_batched(A_NORMAL, g, mmax, mmax, mmax, batch, st));
        }
    }
    TICK();   // 3: back-multiplication done
    // ---- reduced operator ----
    GemmParams gp = gemm_params_zero();
    gp.sig = (const cplx*)signals; gp.sig_off = d_soff;
    // T1 = U^p * Rs
    gp.A = nullptr; gp.strideA = 0; gp.lda = 0;
    gp.B = bRs; gp.strideB = stride; gp.ldb = ld;
    gp.C = bT1; gp.strideC = stride; gp.ldc = ld;
    gp.Mv = d_mv; gp.Nv = d_lv; gp.Kv = d_mv; gp.shift = p;
    CK(zgemm_batched(A_HANKEL, gp, mmax, lmax, mmax, batch, st));
    // Ured = Lt^H * T1
    cplx* bUred = bH;
    gp.A = bLt; gp.strideA = stride; gp.lda = ld;
    gp.B = bT1; gp.strideB = stride; gp.ldb = ld;
    gp.C = bUred; gp.strideC = stride; gp.ldc = ld;
    gp.Mv = d_lv; gp.Nv = d_lv; gp.Kv = d_mv; gp.shift = 0;
    CK(zgemm_batched(A_CONJT, gp, lmax, lmax, mmax, batch, st));
    if (dbg) CK(cudaMemcpyAsync(mat(5), bUred, sizeof(cplx) * batch * stride, cudaMemcpyDeviceToDevice, st));
    TICK();   // 4: T1 + Ured done
    // ---- eigen-decomposition of Ured ----
    {
        // blocked (compact-WY) Hessenberg reduction: panel kernel + DMMA trailing updates
        cplx* Vp = (cplx*)(ws + L.vp); cplx* Yp = (cplx*)(ws + L.yp); cplx* VTp = (cplx*)(ws + L.vtp);
        cplx* Wp = (cplx*)(ws + L.wp); cplx* Tws = (cplx*)(ws + L.tws);
        const long long pstride = (long long)ld * HB_NB;
        size_t sm = (size_t)(2 * ld + HB_NB * HB_NB + 4 * HB_NB) * 16 + 512;
        CK(cudaFuncSetAttribute(hess_panel_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        CK(cudaFuncSetAttribute(hess_panel_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        const int hp_csize = (batch <= 74) ? pick_cluster_size(hess_panel_kernel<1>, batch, E_THREADS, sm, LLCK_MAX_CLUSTER, o.cluster_size) : 1;
        const bool hp_two = panels_two_per_sm(batch, sm);
        GemmParams hp = gemm_params_zero();
        int k0_last = 0;
        for (int k0 = 0; k0 + 2 < lmax; k0 += HB_NB) {
            k0_last = k0;
            CK(launch_clustered(hp_two ? hess_panel_kernel<2> : hess_panel_kernel<1>, batch, hp_csize, E_THREADS, sm, st, bH, stride, ld, d_lv, k0, Vp, Yp, VTp, pstride, Tws, pstride,
                                (cplx*)(ws + L.ypart), hp_csize));
            const int e = k0 + HB_NB;
            const int r0 = k0 + 1;
            // rows 0..k0 were left out of the panel kernel:  Y[0:r0, :] = A[0:r0, r0:] * (V T)[r0:, :]
            hp = gemm_params_zero();
            hp.A = bH + (long long)ld * r0; hp.strideA = stride; hp.lda = ld;
            hp.B = VTp + r0; hp.strideB = pstride; hp.ldb = ld;
            hp.C = Yp; hp.strideC = pstride; hp.ldc = ld;
            hp.Mc = r0; hp.Nc = HB_NB; hp.Kv = d_lv; hp.Kc = -r0;
            CK(zgemm_batched(A_NORMAL, hp, r0, HB_NB, lmax - r0, batch, st));
            // ... and the panel's own columns above the panel:  A[0:r0, r0:e] -= Y[0:r0, :] * V[r0:e, :]^H
            hp = gemm_params_zero();
            hp.A = Yp; hp.strideA = pstride; hp.lda = ld;
            hp.B = Vp + r0; hp.strideB = pstride; hp.ldb = ld;
            hp.C = bH + (long long)ld * r0; hp.strideC = stride; hp.ldc = ld;
            hp.Mc = r0; hp.Nc = HB_NB - 1; hp.Kc = HB_NB; hp.accum = 1;
            CK(zgemm_batched(A_NORMAL, hp, r0, HB_NB - 1, HB_NB, batch, st, true));
            if (e >= lmax) continue;
            // A[:, e:] -= Y * V[e:, :]^H
            hp = gemm_params_zero();
            hp.A = Yp; hp.strideA = pstride; hp.lda = ld;
            hp.B = Vp + e; hp.strideB = pstride; hp.ldb = ld;
            hp.C = bH + (long long)ld * e; hp.strideC = stride; hp.ldc = ld;
            hp.Mv = d_lv; hp.Nv = d_lv; hp.Nc = -e; hp.Kc = HB_NB; hp.accum = 1;
            CK(zgemm_batched(A_NORMAL, hp, lmax, lmax - e, HB_NB, batch, st, true));
            // W = (V T)[k0+1:, :]^H * A[k0+1:, e:]
            hp = gemm_params_zero();
            hp.A = VTp + (k0 + 1); hp.strideA = pstride; hp.lda = ld;
            hp.B = bH + (k0 + 1) + (long long)ld * e; hp.strideB = stride; hp.ldb = ld;
            hp.C = Wp; hp.strideC = pstride; hp.ldc = HB_NB;
            hp.Mc = HB_NB; hp.Nv = d_lv; hp.Nc = -e; hp.Kv = d_lv; hp.Kc = -(k0 + 1);
            CK(zgemm_batched(A_CONJT, hp, HB_NB, lmax - e, lmax - k0 - 1, batch, st));
            // A[k0+1:, e:] -= V[k0+1:, :] * W
            hp = gemm_params_zero();
            hp.A = Vp + (k0 + 1); hp.strideA = pstride; hp.lda = ld;
            hp.B = Wp; hp.strideB = pstride; hp.ldb = HB_NB;
            hp.C = bH + (k0 + 1) + (long long)ld * e; hp.strideC = stride; hp.ldc = ld;
            hp.Mv = d_lv; hp.Mc = -(k0 + 1); hp.Nv = d_lv; hp.Nc = -e; hp.Kc = HB_NB; hp.accum = 1;
            CK(zgemm_batched(A_NORMAL, hp, lmax - k0 - 1, lmax - e, HB_NB, batch, st));
        }
        // Q = P_0 ... P_{n-3}: blocked backward accumulation
        {
            dim3 grid(128, batch);
            set_identity_kernel<<<grid, 256, 0, st>>>(bZ, stride, ld, d_lv);
            LLCK_LAUNCHED();
            CK(cudaGetLastError());
        }
        if (lmax > 2) {
            for (int k0 = k0_last; k0 >= 0; k0 -= HB_NB) {
                hess_qpanel_kernel<<<batch, E_THREADS, 0, st>>>(bH, stride, ld, d_lv, k0, Vp, VTp, pstride, Tws, pstride);
                LLCK_LAUNCHED();
                CK(cudaGetLastError());
                // W = (V T^H)[k0+1:, :]^H * Q[k0+1:, k0+1:]
                hp = gemm_params_zero();
                hp.A = VTp + (k0 + 1); hp.strideA = pstride; hp.lda = ld;
                hp.B = bZ + (k0 + 1) + (long long)ld * (k0 + 1); hp.strideB = stride; hp.ldb = ld;
                hp.C = Wp; hp.strideC = pstride; hp.ldc = HB_NB;
                hp.Mc = HB_NB; hp.Nv = d_lv; hp.Nc = -(k0 + 1); hp.Kv = d_lv; hp.Kc = -(k0 + 1);
                CK(zgemm_batched(A_CONJT, hp, HB_NB, lmax - k0 - 1, lmax - k0 - 1, batch, st));
                // Q[k0+1:, k0+1:] -= V[k0+1:, :] * W
                hp = gemm_params_zero();
                hp.A = Vp + (k0 + 1); hp.strideA = pstride; hp.lda = ld;
                hp.B = Wp; hp.strideB = pstride; hp.ldb = HB_NB;
                hp.C = bZ + (k0 + 1) + (long long)ld * (k0 + 1); hp.strideC = stride; hp.ldc = ld;
                hp.Mv = d_lv; hp.Mc = -(k0 + 1); hp.Nv = d_lv; hp.Nc = -(k0 + 1); hp.Kc = HB_NB; hp.accum = 1;
                CK(zgemm_batched(A_NORMAL, hp, lmax - k0 - 1, lmax - k0 - 1, HB_NB, batch, st));
            }
        }
        {
            dim3 grid(128, batch);
            clear_below_subdiag_kernel<<<grid, 256, 0, st>>>(bH, stride, ld, d_lv);
            LLCK_LAUNCHED();
            CK(cudaGetLastError());
        }
        if (dbg) {
            CK(cudaMemcpyAsync(mat(6), bH, sizeof(cplx) * batch * stride, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(mat(7), bZ, sizeof(cplx) * batch * stride, cudaMemcpyDeviceToDevice, st));
        }
        TICK();   // 5: hessenberg done
        CK(cudaFuncSetAttribute(hqr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HQR_SMEM_BYTES));
        // AED window: measured on B200 (tools/aed_sweep.py, hqr ms at window 24 / 28 / 32 with nibble 60): l = 1024: 589 / 576 / 561,
        // l = 768: 297 / 287 / 282, l = 640: 388 / 375 / 375, l = 512: 470 / 466 / 475, l = 384: 264 / 267 / 279, l = 256: 260 / 274 / 295
        const int aed_nw = o.aed_window > 0 ? o.aed_window : (lmax > 704 ? 32 : (lmax > 448 ? 28 : 24));
        // small batches: a thread-block cluster of csize CTAs per member shares the strip GEMMs (the window chase / AED run redundantly in
        // every CTA of the cluster); csize = largest power of two that still gives every cluster its own SMs
        const int csize = (batch <= 74) ? pick_cluster_size(hqr_kernel, batch, E_THREADS, HQR_SMEM_BYTES, LLCK_MAX_CLUSTER, o.cluster_size) : 1;
        CK(launch_clustered(hqr_kernel, batch, csize, E_THREADS, HQR_SMEM_BYTES, st, bH, bZ, stride, ld, d_lv, status, d_hqrs,
                            (long long*)o.hqr_profile, 1, aed_nw, (int)o.aed_nibble, csize));
        TICK();   // 6: hqr done
        {
            dim3 gz(128, batch);
            trevc_zero_kernel<<<gz, 256, 0, st>>>(bXev, stride, ld, d_lv);
            LLCK_LAUNCHED();
            CK(cudaGetLastError());
            for (int j0 = ((lmax - 1) / TV_NB) * TV_NB; j0 >= 0; j0 -= TV_NB) {
                dim3 gd((lmax - j0 + 127) / 128, batch);
                trevc_diag_kernel<<<gd, 128, 0, st>>>(bH, bXev, stride, ld, d_lv, j0);
                LLCK_LAUNCHED();
                CK(cudaGetLastError());
                if (j0 > 0) {
                    // X[0:j0, j0:n] -= T[0:j0, j0:j0+32] * X[j0:j0+32, j0:n]
                    GemmParams tp = gemm_params_zero();
                    tp.A = bH + (long long)ld * j0; tp.strideA = stride; tp.lda = ld;
                    tp.B = bXev + j0 + (long long)ld * j0; tp.strideB = stride; tp.ldb = ld;
                    tp.C = bXev + (long long)ld * j0; tp.strideC = stride; tp.ldc = ld;
                    tp.Mc = j0; tp.Nv = d_lv; tp.Nc = -j0; tp.Kv = d_lv; tp.Kc = -j0; tp.Kcap = TV_NB; tp.accum = 1;
                    CK(zgemm_batched(A_NORMAL, tp, j0, lmax - j0, TV_NB, batch, st));
                }
            }
            dim3 gn((lmax + 7) / 8, batch);
            trevc_normalize_kernel<<<gn, 256, 0, st>>>(bXev, stride, ld, d_lv);
            LLCK_LAUNCHED();
            CK(cudaGetLastError());
        }
    }
    TICK();   // 7: trevc done
    // P = Z * Xev
    gp.A = bZ; gp.strideA = stride; gp.lda = ld;
    gp.B = bXev; gp.strideB = stride; gp.ldb = ld;
    gp.C = bP; gp.strideC = stride; gp.ldc = ld;
    gp.Mv = d_lv; gp.Nv = d_lv; gp.Kv = d_lv; gp.triB = 1;      // Xev is upper triangular
    CK(zgemm_batched(A_NORMAL, gp, lmax, lmax, lmax, batch, st));
    gp.triB = 0;
    // B = Rs * P
    gp.A = bRs; gp.B = bP; gp.C = bB;
    gp.Mv = d_mv; gp.Nv = d_lv; gp.Kv = d_lv;
    CK(zgemm_batched(A_NORMAL, gp, mmax, lmax, lmax, batch, st));
    // W = U0 * B
    gp.A = nullptr; gp.B = bB; gp.C = bW;
    gp.Mv = d_mv; gp.Nv = d_lv; gp.Kv = d_mv; gp.shift = 0;
    CK(zgemm_batched(A_HANKEL, gp, mmax, lmax, mmax, batch, st));
    TICK();   // 8: back-transform GEMMs done
    // ---- amplitudes / line list ----
    {
        dim3 grid((lmax + 7) / 8, batch);
        epilogue_kernel<<<grid, 256, 0, st>>>(bB, bW, bH, stride, ld, d_mv, d_lv, dwell, line_lists, ll_stride,
                                              (cplx*)mu_out, (cplx*)d_out, mu_stride, n_valid, status);
        LLCK_LAUNCHED();
        CK(cudaGetLastError());
    }
    TICK();   // 9: epilogue done
    if (info) {
        for (int i = 0; i < 16; ++i) info[i] = 0;
        info[2] = ld; info[3] = nbmax;
        info[13] = llck_launch_count;                 // kernels enqueued by this call, counted at the launch sites
        info[0] = graph_us[0] + graph_us[1] + graph_us[2] + graph_us[3];      // host microseconds spent building / launching the Jacobi loop graph
        info[14] = jacobi_graph;                      // 1: the Jacobi sweeps ran as a device-side WHILE graph node, 0: enqueued unconditionally
    }
    if (timing) {
        // diagnostic mode: the only path that waits for the stream.  info[1] = max QR sweeps over the batch,
        // info[4..12] = stage durations in microseconds
        int* h = (int*)malloc(sizeof(int) * (size_t)batch);
        if (h) {
            cudaError_t e3 = cudaMemcpyAsync(h, d_hqrs, sizeof(int) * batch, cudaMemcpyDeviceToHost, st);
            if (e3 == cudaSuccess) e3 = cudaStreamSynchronize(st);
            int h_maxs = 0;
            if (e3 == cudaSuccess) for (int b = 0; b < batch; ++b) if (h[b] > h_maxs) h_maxs = h[b];
            free(h);
            if (e3 != cudaSuccess) return -(int)e3;
            info[1] = h_maxs;
        } else {
            CK(cudaStreamSynchronize(st));
        }
        for (int i = 0; i + 1 < ntev; ++i) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, G.tev[i], G.tev[i + 1]));
            info[4 + i] = (int32_t)(ms * 1000.0f);
        }
    }
#undef TICK
    return 0;
}

}  // extern "C"
