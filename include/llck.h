/* llck.h -- C ABI of the B200-native batched KBDM solver (libllck.so).
 *
 * The reference (danilomendesdias/llckbdm) is pure Python and has no FFI; the drop-in boundary is its
 * Python function surface.  This header is what a binding of that surface calls:
 *
 *   llck_kbdm_batched   replaces the body of  llckbdm/kbdm.py:19-92  (kbdm: Hankel U matrices :95-130,
 *                       SVD-reduced GEP :133-212, eigenvector normalisation :215-240, line list :71-92)
 *                       for a whole ensemble at once, i.e. the serial loop of
 *                       llckbdm/sampling.py:52-70 (sample_kbdm) becomes ONE call.
 *
 * Conventions: plain pointers and sizes only; returns 0 on success, a negative cudaError_t on a CUDA
 * failure, or a positive LLCK_E_* code on bad arguments; never throws; allocates no device memory (the
 * caller owns every buffer, including the workspace); all device work is issued on `stream` and the
 * calls are STREAM-ORDERED AND ASYNCHRONOUS: they return as soon as the work is enqueued and never
 * wait for the stream (exceptions, documented at the entry: LLCK_FLAG_TIMING and the *_test stage
 * entries).  Every data-dependent decision (which members need the Jacobi SVD, Jacobi convergence,
 * QR deflation) is taken on the device.  Host arrays passed to a call (m, l, sig_offset, sig_len) are
 * staged before the call returns and may be reused at once.  No environment variable or other
 * hidden state changes the behaviour of the library: every knob is an explicit argument (flags, llck_options).
 * Per-member numerical failures are reported through status[] (the Python wrapper turns them into
 * numpy.linalg.LinAlgError like np.linalg.inv / scipy.linalg.eig would, kbdm.py:186,192).
 *
 * Pointers marked [dev] are device pointers on the current CUDA device, [host] are host pointers.
 * Complex values are interleaved (re, im) float64 pairs == numpy complex128.
 */
#ifndef LLCK_H
#define LLCK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LLCK_VERSION 200

/* argument errors (positive return values) */
#define LLCK_E_BADARG 1
#define LLCK_E_WORKSPACE 2           /* workspace_bytes < llck_workspace_bytes(batch, ld, flags) */
#define LLCK_E_SHORT_SIGNAL 3        /* sig_len[b] < 2 m[b] + p - 1: the Hankel matrices would read past the member's FID (kbdm.py:59-62) */
#define LLCK_E_TOO_LARGE 4           /* max m > LLCK_M_MAX (llck_kbdm_batched) */

/* largest supported Hankel dimension (the panel kernels keep 2 vectors of the leading dimension in shared memory) */
#define LLCK_M_MAX 2048

/* per-member status[] values */
#define LLCK_STATUS_OK 0
#define LLCK_STATUS_QR_NOCONV 1      /* Hessenberg-QR did not converge (scipy.linalg.eig -> LinAlgError) */
#define LLCK_STATUS_SINGULAR 2       /* zero singular value among the l kept (np.linalg.inv -> LinAlgError, kbdm.py:186) */
#define LLCK_STATUS_NONFINITE 3      /* non-finite pole or amplitude */
#define LLCK_STATUS_SVD_NOCONV 4     /* Jacobi SVD (fallback for rank-deficient members) hit the sweep limit */

/* flags */
#define LLCK_FLAG_DEBUG_KEEP 1       /* keep every intermediate in its own workspace matrix (tests only) */
#define LLCK_FLAG_TIMING 2           /* diagnostic: record CUDA events between stages, WAIT for the stream, return stage durations (us)
                                        in info[4..12] and the maximum number of QR sweeps in info[1] */

#define LLCK_FLAG_NO_GRAPH 4         /* do not use a CUDA graph WHILE node for the Jacobi sweeps: enqueue all of them (their CTAs exit
                                        at once for converged members); for callers that are themselves capturing the stream */

/* SVD back end for the real bidiagonal (llck_options.svd_mode) */
#define LLCK_SVD_DC 0                /* divide and conquer; members it flags as numerically rank deficient fall back to Jacobi (default) */
#define LLCK_SVD_JACOBI 1            /* block one-sided Jacobi for every member */

/* Tuning knobs of llck_kbdm_batched.  Zero-initialise, set struct_size = sizeof(llck_options), change what you need; a NULL
 * pointer means all defaults.  (Fields are only ever appended; struct_size tells the library which ones the caller knows.) */
typedef struct llck_options {
    int32_t struct_size;
    int32_t svd_mode;            /* LLCK_SVD_* */
    int32_t cluster_size;        /* CTAs per member of the one-CTA-per-member kernels for batches <= 74: 0 = auto, else 1, 2, 4 or 8 */
    int32_t aed_window;          /* aggressive-early-deflation window of the multishift QR, 8..48; 0 = default (24 / 28 / 32 for l <= 448 / <= 704 / larger) */
    int32_t aed_nibble;          /* percent of the window that must deflate to skip the sweep (LAPACK NIBBLE); 0 = default (60) */
    int32_t jacobi_max_sweeps;   /* sweeps enqueued for the Jacobi SVD (converged members exit at once); 0 = default (30) */
    double  jacobi_conv;         /* scaled off-diagonal threshold that ends a member's Jacobi iteration; 0 = default (1e-6) */
    void*   hqr_profile;         /* [dev] int64 [batch][10], optional: clock64 phase split of hqr_kernel per member (profiling) */
} llck_options;

int llck_version(void);

/* Waits for and releases the CUDA graphs earlier llck_kbdm_batched calls launched (see there).  Optional: every call releases
 * the graphs of previous calls that have completed; call this before unloading the library or tearing the context down. */
int llck_release_resources(void);

/* Leading dimension used for every per-member matrix: round_up(max m, 64). */
int llck_leading_dim(int m_max);

/* Workspace size in bytes for `batch` members with leading dimension ld = llck_leading_dim(max m).
 * Pure function. flags: 0 or LLCK_FLAG_DEBUG_KEEP. */
size_t llck_workspace_bytes(int batch, int ld, int flags);

/* Byte offset of debug matrix `which` (0..13) of member 0 inside the workspace (LLCK_FLAG_DEBUG_KEEP layout);
 * member b is at offset + b * ld*ld*16.
 * 0:X(=L*S) 1:V 2:Rs 3:Lt 4:T1 5:Ured 6:Hhess 7:Qhess 8:T 9:Z 10:Xev 11:P 12:B 13:W */
size_t llck_debug_offset(int batch, int ld, int which);

/* Batched KBDM solve.  Member b uses the FID  signals[sig_offset[b] ... sig_offset[b] + sig_len[b])  (it reads the first
 * 2*m[b]+p-1 points; a shorter FID is LLCK_E_SHORT_SIGNAL, nothing is launched), Hankel dimension m[b] <= LLCK_M_MAX, kept rank
 * l[b] (1 <= l <= m), shift p >= 1, Tikhonov q >= 0 (kbdm.py:19).  signals [dev] complex; sig_offset, sig_len [host] int64 [batch]
 * (in complex elements); m, l [host] int32 [batch] -- sizes stay on the host because they set the launch geometry.
 *
 *   line_lists [dev]  float64 [batch][ll_stride]   rows (A, T2, F, PH), l[b] rows per member, eig order (kbdm.py:88-92)
 *   mu_out     [dev]  complex [batch][mu_stride]   raw poles (optional, may be NULL)
 *   d_out      [dev]  complex [batch][mu_stride]   complex amplitudes D_k (optional, may be NULL)
 *   sing_vals  [dev]  float64 [batch][sv_stride]   all m[b] singular values, descending (kbdm.py:68,207)
 *   n_valid    [dev]  int32   [batch]              rows passing filter_samples (A>1e-6 and T2>0, sampling.py:92-95)
 *   status     [dev]  int32   [batch]              LLCK_STATUS_*
 *   opts       [host] llck_options (optional, NULL = defaults)
 *   info       [host] int32   [16] (optional)      [0]=host microseconds spent building and launching the Jacobi loop graph, [2]=ld,
 *                                                  [3]=Jacobi column blocks of the largest member, [13]=kernels enqueued by
 *                                                  this call (counted at the launch sites; a graph launch counts once),
 *                                                  [14]=1 if the Jacobi sweeps ran as a device-side WHILE graph node; with
 *                                                  LLCK_FLAG_TIMING also [1]=max QR multishift sweeps and [4..12]=stage durations in us
 *                                                  (init + bidiagonalisation, SVD of the bidiagonal, back-multiplication, T1+Ured,
 *                                                  hessenberg, hqr, trevc, P+B+W, epilogue)
 * Workspace: llck_workspace_bytes(batch, ld, flags) = 11 ld x ld complex matrices per member (6 pipeline + 5 for the
 * divide-and-conquer SVD of the bidiagonal) + panel / bookkeeping vectors.  Batches of <= 74 members run the one-CTA-per-member
 * kernels as thread-block clusters of 2/4/8 CTAs per member.
 * Asynchronous: returns once the launch sequence is enqueued on `stream` (LLCK_FLAG_TIMING makes it wait).  The Jacobi fallback
 * for members the divide-and-conquer SVD flags is ONE CUDA-graph launch whose conditional WHILE node repeats the sweep on the
 * device until every member has converged (at most llck_options.jacobi_max_sweeps times; zero times when no member was flagged),
 * so no decision needs a device-to-host read-back.  (A small graph is built inside the call -- host-side objects, no device
 * memory -- and, because an executable graph cannot be destroyed while in flight without blocking, released by a later call or
 * by llck_release_resources() once its launch has completed: the library's only process-wide state.)
 */
int llck_kbdm_batched(const void* signals, const int64_t* sig_offset, const int64_t* sig_len, const int32_t* m, const int32_t* l,
                      int32_t p, double q, double dwell, int32_t batch,
                      double* line_lists, int64_t ll_stride,
                      void* mu_out, void* d_out, int64_t mu_stride,
                      double* sing_vals, int64_t sv_stride,
                      int32_t* n_valid, int32_t* status,
                      void* workspace, size_t workspace_bytes, int32_t flags, const llck_options* opts,
                      void* stream, int32_t* info);

/* Stage-level entry (tests / profiling): one complex GEMM  C = opA(A) * B  through the production kernel.
 * amode: 0 normal, 1 conj-transpose (A stored K x M), 2 implicit Hankel (A[i,k] = sig[i+k+shift]). All [dev].
 * Test helper: allocates two small device scratch arrays and waits for the stream. */
int llck_zgemm(int32_t amode, const void* A, int32_t lda, const void* B, int32_t ldb, void* C, int32_t ldc,
               int32_t M, int32_t N, int32_t K, const void* sig, int32_t shift, void* stream);

/* Stage-level entry (tests): blocked Householder bidiagonalisation A = Q B P^H of one m x m matrix (column-major, ld).
 * A [dev] is overwritten by the reflectors; d_out[m], e_out[m-1] [dev] = the REAL diagonal / superdiagonal of B;
 * Q, P [dev] ld x m complex.  Allocates its own scratch (test helper, not part of the hot path). */
int llck_bidiag_test(void* A, int32_t m, int32_t ld, double* d_out, double* e_out, void* Q, void* P, void* stream);

/* Frequency-domain RMSE of `batch` candidate line lists against one FID -- replaces the scoring loop of
 * llckbdm/min_rmse_kbdm.py:33-41 (calculate_freq_domain_rmse, llckbdm/metrics.py:7-17: multi_fid synthesis on
 * t = n*dwell, fft/sqrt(N) of data and model, RMSE of the REAL parts).  Stream-ordered, asynchronous, allocates nothing.
 *   data        device complex128 [N]
 *   line_lists  device float64 [batch][ll_stride], rows (A, T2, F, PH) as written by llck_kbdm_batched
 *   n_rows      device int32 [batch]: rows of each candidate
 *   filter      1: skip rows failing  A > amplitude_tol and T2 > 0  (filter_samples, llckbdm/sampling.py:75-97)
 *   rmse_out    device float64 [batch]; +inf for a candidate without valid rows (min_rmse_kbdm.py:36-37)
 * Any N: FIDs of up to 12800 points keep the whole model in shared memory, longer ones are tiled over n.             */
int llck_rmse_batched(const void* data, int32_t N, double dwell, const double* line_lists, int64_t ll_stride,
                      const int32_t* n_rows, int32_t batch, int32_t filter, double amplitude_tol, double* rmse_out, void* stream);

/* Silhouette coefficients of `nclusterings` labelings of the same n points -- replaces sklearn.metrics.silhouette_samples as
 * called by llckbdm/llckbdm.py:291 on the 4-dimensional line-list features (llckbdm.py:202-230); Euclidean metric, every label
 * value (HDBSCAN's noise label included) is a cluster, singleton clusters score 0.  Stream-ordered, asynchronous.
 *   X           device float64 [n][4]
 *   order       device int32 [nclusterings][n]    point indices sorted by label
 *   seg         device int32 [nclusterings][n+1]  start offset of each cluster in `order` (nseg+1 entries used)
 *   nseg        device int32 [nclusterings]       number of clusters
 *   cluster_of  device int32 [nclusterings][n]    cluster index (0..nseg-1) of the point at each sorted position
 *   out         device float64 [nclusterings][n]  silhouette of every point, indexed by ORIGINAL point index              */
int llck_silhouette_batched(const double* X, int32_t n, const int32_t* order, const int32_t* seg, const int32_t* nseg,
                            const int32_t* cluster_of, int32_t nclusterings, double* out, void* stream);

/* Pooled filtered line lists + clustering features of a solved ensemble -- replaces the host sequence of llckbdm/llckbdm.py:94-98
 * (np.concatenate of the members' line lists, filter_samples of llckbdm/sampling.py:75-97, _transform_line_lists of
 * llckbdm/llckbdm.py:202-230) by one launch that reads the solver's output buffer.  Rows keep member order and row order.
 *   line_lists  device float64 [batch][ll_stride] as written by llck_kbdm_batched;  n_rows device int32 [batch] (= l)
 *   offset      device int64 [batch]: first output row of each member = exclusive prefix sum of the members' valid-row counts
 *               (the n_valid output of llck_kbdm_batched)
 *   samples     device float64 [total][4]: kept rows (A, T2, F, PH);   features  device float64 [total][4]: (Re mu, Im mu, A, 0),
 *               mu = exp(i dwell (2 pi F + i/T2)).  Stream-ordered, asynchronous.                                              */
int llck_pool_features(const double* line_lists, int64_t ll_stride, const int32_t* n_rows, const int64_t* offset, int32_t batch,
                       double dwell, double amplitude_tol, double* samples, double* features, void* stream);

/* The two O(n^2) stages of the HDBSCAN fits of llckbdm/llckbdm.py:104-116 (one fit per min_samples value on the SAME points),
 * bit-compatible with sklearn.cluster.HDBSCAN's Euclidean Prim path (the stand-in for the un-vendored `hdbscan` package):
 * llck_hdbscan_core_distances: core[k-1][i] = distance of point i to its k-th nearest neighbour (itself included), k = 1..kmax
 *                              (replaces one KD-tree kneighbors query per fit);  X device float64 [n][4], core device [kmax][n].
 * llck_hdbscan_mst:            Prim's spanning tree of the mutual-reachability graph max(core[a], core[b], |a - b|) for `nfits`
 *                              fits at once (replaces mst_from_data_matrix per fit); core_row[f] selects the row of `core`
 *                              (= min_samples - 1); edges in insertion order into mst_src / mst_dst / mst_w [nfits][n-1];
 *                              min_reach [nfits][n] and cur_src [nfits][n] are scratch.  n <= 131072.  Stream-ordered, asynchronous.
 *                              Up to 45,056 points a fit runs on a thread-block cluster of 8 CTAs with its points resident in
 *                              shared memory and registers (no global traffic per step); larger sets, or LLCK_MST_SINGLE_CTA,
 *                              use one CTA per fit that streams the points from L2.  Both give bit-identical edges.        */
int llck_hdbscan_core_distances(const double* X, int32_t n, int32_t kmax, double* core, void* stream);
#define LLCK_MST_SINGLE_CTA 1         /* llck_hdbscan_mst flags: always use the one-CTA-per-fit kernel */
#define LLCK_MST_DIM3 2               /* the caller guarantees that all points share their 4th coordinate (the LLC-KBDM features do: it is 0):
                                         the cluster kernel leaves it out of the distances -- bit-identical, fewer FP64 instructions */
int llck_hdbscan_mst(const double* X, int32_t n, const double* core, const int32_t* core_row, int32_t nfits,
                     double* min_reach, int32_t* cur_src, int64_t* mst_src, int64_t* mst_dst, double* mst_w, int32_t flags, void* stream);

/* HOST function: flat cluster labels of `nfits` HDBSCAN fits from their spanning trees -- the rest of every fit of
 * llckbdm/llckbdm.py:280-283 after the spanning tree (single-linkage dendrogram, condensed tree, stabilities, excess-of-mass
 * selection, labelling) with the clusterer's defaults (min_cluster_size 5 -> pass 5, "eom", allow_single_cluster False,
 * cluster_selection_epsilon 0).  Label for label what sklearn.cluster.HDBSCAN (the stand-in for the un-vendored `hdbscan`
 * package) returns for the same sorted edges: rows and floating-point sums are formed in its order.  The fits run on
 * `nthreads` host threads (0 = all hardware threads).  All pointers [host].
 *   mst_src, mst_dst  int64 [nfits][n-1], mst_w float64 [nfits][n-1]: edges as written by llck_hdbscan_mst
 *   order             int64 [nfits][n-1] (optional): permutation sorting each fit's edges by weight, ascending -- pass the
 *                     clusterer's own argsort so that equal weights are processed in its order; NULL = already sorted
 *   labels            int32 [nfits][n] out: 0..k-1, -1 = noise                                                              */
int llck_hdbscan_labels(const int64_t* mst_src, const int64_t* mst_dst, const double* mst_w, const int64_t* order, int32_t n, int32_t nfits,
                        int32_t min_cluster_size, int32_t nthreads, int32_t* labels);

/* Batched FID synthesis -- llckbdm/sig_gen.py:57-71 (multi_fid) for `batch` parameter sets at once on t_n = n * dwell
 * (the benchmark inputs of configs C4 / C5 and the residual model of llckbdm/llckbdm.py:177).
 *   params  device float64 [batch][pstride], rows (A, T2, F, PH);  n_rows device int32 [batch];  out device complex128 [batch][N] */
int llck_multi_fid_batched(const double* params, int64_t pstride, const int32_t* n_rows, int32_t batch, int32_t N, double dwell,
                           void* out, void* stream);

/* Stage entry (tests): divide-and-conquer SVD of `batch` real upper-bidiagonal matrices (second half of the replacement of
 * scipy.linalg.svd, llckbdm/kbdm.py:166).  d, e: device [batch][ld] (diagonal m, super-diagonal m-1); m: host [batch];
 * ld multiple of 64.  Outputs (device): sing_vals [batch][ld] descending, Us = U*diag(s) and V as complex128 [batch][ld*ld]
 * column-major with zero imaginary parts, fallback [batch] = 1 for members the solver leaves to the Jacobi path
 * (numerically rank deficient).  Synchronous. */
int llck_bdc_test(const double* d, const double* e, const int32_t* m, int32_t batch, int32_t ld,
                  double* sing_vals, void* Us, void* V, int32_t* fallback, void* stream);

#ifdef __cplusplus
}
#endif
#endif
