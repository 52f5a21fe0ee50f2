// The two O(n^2) stages of the HDBSCAN fits of the LLC-KBDM clustering loop (reference llckbdm/llckbdm.py:104-116 runs one
// fit per min_samples value 1..M-1 on the SAME points), bit-compatible with the host clusterer that stands in for the
// un-vendored `hdbscan` package here (sklearn.cluster.HDBSCAN, euclidean metric, Prim path):
//   core distances : distance to the k-th nearest neighbour, k = 1..K in ONE brute-force pass (sklearn: one KD-tree query per fit)
//   spanning tree  : Prim's algorithm on the mutual-reachability graph exactly as sklearn's mst_from_data_matrix runs it (start at
//                    node 0, strict-less updates, lowest index wins ties), one CTA per min_samples value, all fits concurrently
// Distances are sqrt(((dx^2 + dy^2) + dz^2) + dw^2) with separately rounded multiplies and adds (the host code is compiled
// without FMA), so the edge lists are identical to the host's and the host's tree condensation yields identical labels.
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>

#define HDB_KMAX 128
#define HDB_TILE 256

__device__ __forceinline__ double hdb_rdist(const double4 a, const double4 b) {
    const double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
    double d = __dmul_rn(dx, dx);
    d = __dadd_rn(d, __dmul_rn(dy, dy));
    d = __dadd_rn(d, __dmul_rn(dz, dz));
    d = __dadd_rn(d, __dmul_rn(dw, dw));
    return d;
}

// the same with the last coordinate left out: bit-identical when all points share it (dw = 0 exactly, d + 0 = d); the LLC-KBDM
// features are (Re mu, Im mu, A, 0) (reference llckbdm.py:219 zeroes the phase feature), and Prim's inner loop is FP64-issue bound
__device__ __forceinline__ double hdb_rdist3(const double4 a, const double4 b) {
    const double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    double d = __dmul_rn(dx, dx);
    d = __dadd_rn(d, __dmul_rn(dy, dy));
    d = __dadd_rn(d, __dmul_rn(dz, dz));
    return d;
}

// core[k-1][i] = distance from point i to its k-th nearest neighbour (itself included), k = 1..K
__global__ void __launch_bounds__(128) hdb_core_kernel(const double* __restrict__ X, int n, int K, double* __restrict__ core) {
    __shared__ double4 tile[HDB_TILE];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    const double4 xi = live ? reinterpret_cast<const double4*>(X)[i] : make_double4(0, 0, 0, 0);
    double best[HDB_KMAX];                       // ascending squared distances (local memory)
    for (int k = 0; k < K; ++k) best[k] = INFINITY;
    double worst = INFINITY;
    for (int t0 = 0; t0 < n; t0 += HDB_TILE) {
        __syncthreads();
        for (int q = threadIdx.x; q < HDB_TILE; q += blockDim.x)
            if (t0 + q < n) tile[q] = reinterpret_cast<const double4*>(X)[t0 + q];
        __syncthreads();
        const int cnt = min(HDB_TILE, n - t0);
        if (live) {
            for (int q = 0; q < cnt; ++q) {
                const double d = hdb_rdist(xi, tile[q]);
                if (d < worst) {
                    int k = K - 1;
                    while (k > 0 && best[k - 1] > d) { best[k] = best[k - 1]; --k; }
                    best[k] = d;
                    worst = best[K - 1];
                }
            }
        }
    }
    if (live) for (int k = 0; k < K; ++k) core[(long long)k * n + i] = __dsqrt_rn(best[k]);
}

// One CTA per fit f: Prim on max(core_f[a], core_f[b], dist(a, b)).  Edge i of the tree = (src[i], dst[i], w[i]) in insertion order.
// min_reach / cur_src: per-fit scratch [nfits][n].
#define HDB_PRIM_THREADS 1024
__global__ void __launch_bounds__(HDB_PRIM_THREADS) hdb_prim_kernel(const double* __restrict__ X, int n, const double* __restrict__ core,
                                                                    const int* __restrict__ core_row, double* __restrict__ min_reach,
                                                                    int* __restrict__ cur_src, long long* __restrict__ mst_src,
                                                                    long long* __restrict__ mst_dst, double* __restrict__ mst_w) {
    __shared__ double rv[32];
    __shared__ int rj[32], rs[32];
    __shared__ int s_cur;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* cr = core + (long long)core_row[f] * n;
    double* mr = min_reach + (long long)f * n;
    int* cs = cur_src + (long long)f * n;
    for (int j = tid; j < n; j += HDB_PRIM_THREADS) { mr[j] = INFINITY; cs[j] = 1; }      // np.full(inf), np.ones
    // visited flags of this thread's own points j = tid + 1024 * q live in registers (n <= 1024 * 128)
    unsigned long long vis0 = 0ull, vis1 = 0ull;
    int cur = 0;
    __syncthreads();
    for (int i = 0; i < n - 1; ++i) {
        if ((cur & (HDB_PRIM_THREADS - 1)) == tid) {
            const int q = cur >> 10;
            if (q < 64) vis0 |= 1ull << q; else vis1 |= 1ull << (q - 64);
        }
        const double4 xc = reinterpret_cast<const double4*>(X)[cur];
        const double cd = cr[cur];
        double bv = 1.7976931348623157e308;      // DBL_MAX: candidates must be strictly below it
        int bj = 0, bs = 0;
        for (int j = tid, q = 0; j < n; j += HDB_PRIM_THREADS, ++q) {
            const bool visited = (q < 64) ? ((vis0 >> q) & 1ull) : ((vis1 >> (q - 64)) & 1ull);
            if (visited) continue;
            double m = mr[j];
            int src = cs[j];
            // skip the coordinate load, distance and square root when they cannot lower m (see hdb_prim_cluster_kernel)
            const double lb = fmax(cd, cr[j]);
            if (lb < m) {
                const double d2 = hdb_rdist(xc, reinterpret_cast<const double4*>(X)[j]);
                if (d2 < m * m * 1.0000000000000009) {
                    const double mrd = fmax(lb, __dsqrt_rn(d2));
                    if (mrd < m) { m = mrd; src = cur; mr[j] = m; cs[j] = cur; }
                }
            }
            if (m < bv) { bv = m; bj = j; bs = src; }            // ascending j within a thread: strict < keeps the lowest index
        }
        // lexicographic (value, index) minimum over the CTA == the sequential scan's "first strictly smaller" rule
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o), os = __shfl_xor_sync(0xffffffffu, bs, o);
            if (ov < bv || (ov == bv && ov < 1.7976931348623157e308 && oj < bj)) { bv = ov; bj = oj; bs = os; }
        }
        if (lane == 0) { rv[warp] = bv; rj[warp] = bj; rs[warp] = bs; }
        __syncthreads();
        if (warp == 0) {
            bv = rv[lane]; bj = rj[lane]; bs = rs[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, o), os = __shfl_xor_sync(0xffffffffu, bs, o);
                if (ov < bv || (ov == bv && ov < 1.7976931348623157e308 && oj < bj)) { bv = ov; bj = oj; bs = os; }
            }
            if (lane == 0) {
                mst_src[(long long)f * (n - 1) + i] = bs;
                mst_dst[(long long)f * (n - 1) + i] = bj;
                mst_w[(long long)f * (n - 1) + i] = bv;
                s_cur = bj;
            }
        }
        __syncthreads();
        cur = s_cur;
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Prim's algorithm with one thread-block CLUSTER of HDB_CS CTAs per fit: the kernel above re-reads every unvisited point
// (coordinates, core distance, running minimum, source: 52 B) from L2 in each of the n-1 steps -- 99 fits x 43k points is
// L2-bandwidth bound.  Here every CTA keeps its share of the points in shared memory (coordinates) and registers (core distance,
// running minimum, source, visited bit), so a step touches no global memory except the edge it appends:
//   per step: update own points against the current node -> CTA-wide lexicographic (value, index) minimum (warp level: integer
//   redux on the bit pattern of the non-negative distances) -> the CTA's candidate (value, index, source, coordinates and core
//   distance of the candidate) is stored into every peer's shared memory (distributed shared memory) -> ONE cluster barrier ->
//   every CTA picks the winner of the HDB_CS candidates redundantly.
// Same arithmetic and the same tie rule as hdb_prim_kernel: the edge lists are bit-identical.
// Limits: n <= HDB_CS * HDB_CT * HDB_PT_MAX points (45,056); larger sets use hdb_prim_kernel.
// dynamic smem: pt * HDB_CT * 32 B (coordinates) + 2 * HDB_CS candidate records
#define HDB_CS 8
#define HDB_CT 512           // threads per CTA
#define HDB_PT_MAX 11        // own points per thread
struct __align__(16) HdbCand { double v; int j; int s; double4 x; double cd; double pad; };

// lexicographic (value, index) minimum over a warp for non-negative doubles: their bit patterns order like the values
__device__ __forceinline__ void hdb_warp_argmin(double& bv, int& bj, int& bs, double& bc) {
    const unsigned long long key = (unsigned long long)__double_as_longlong(bv);
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    const bool is_min = (hi == mhi) && (lo == mlo);
    const int mj = __reduce_min_sync(0xffffffffu, is_min ? bj : 0x7fffffff);
    const unsigned who = __ballot_sync(0xffffffffu, is_min && bj == mj);
    const int srcl = __ffs(who) - 1;
    bv = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
    bj = mj;
    bs = __shfl_sync(0xffffffffu, bs, srcl);
    bc = __shfl_sync(0xffffffffu, bc, srcl);
}

template <bool DIM3>
__global__ void __launch_bounds__(HDB_CT, 1) hdb_prim_cluster_kernel(const double* __restrict__ X, int n, const double* __restrict__ core,
                                                                     const int* __restrict__ core_row, int pt,
                                                                     long long* __restrict__ mst_src, long long* __restrict__ mst_dst,
                                                                     double* __restrict__ mst_w) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double4* xs = reinterpret_cast<double4*>(smem_raw);                                    // [pt][HDB_CT] own coordinates
    HdbCand* recs = reinterpret_cast<HdbCand*>(xs + (size_t)pt * HDB_CT);                   // [2][HDB_CS]
    constexpr int NW = HDB_CT / 32;
    __shared__ double rv[NW], rc[NW];
    __shared__ int rj[NW], rs[NW];
    __shared__ int s_win;
    const int f = blockIdx.x / HDB_CS, crank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* cr = core + (long long)core_row[f] * n;
    const int per_cta = pt * HDB_CT;
    const int base = crank * per_cta;                       // this CTA owns points [base, base + per_cta)
    double mr[HDB_PT_MAX], crj[HDB_PT_MAX];
    int cs[HDB_PT_MAX];
    unsigned live = 0u;                                      // bit q: own point q exists and is not visited yet
#pragma unroll
    for (int q = 0; q < HDB_PT_MAX; ++q) {
        const int j = base + tid + HDB_CT * q;
        mr[q] = INFINITY; cs[q] = 1; crj[q] = 0.0;
        if (q < pt && j < n) {
            xs[tid + HDB_CT * q] = reinterpret_cast<const double4*>(X)[j];
            crj[q] = cr[j];
            if (j != 0) live |= 1u << q;                     // node 0 is the start: visited
        }
    }
    double4 xc = reinterpret_cast<const double4*>(X)[0];
    double cd = cr[0];
    int cur = 0;
    cluster.sync();
    for (int i = 0; i < n - 1; ++i) {
        double bv = 1.7976931348623157e308, bc = 0.0;        // DBL_MAX: candidates must be strictly below it
        int bj = 0x7ffffffe, bs = 0;
#pragma unroll
        for (int q = 0; q < HDB_PT_MAX; ++q) {
            if ((live >> q) & 1u) {
                // mutual reachability max(cd, cr_j, |x_c - x_j|) replaces mr[q] only if it is strictly smaller: when a core distance
                // already reaches mr[q], or the SQUARED distance clearly exceeds mr[q]^2, nothing changes and the (expensive, FP64
                // pipe bound) distance / square root are skipped -- results are unchanged, the margin covers the roundings
                const double m_old = mr[q], lb = fmax(cd, crj[q]);
                if (lb < m_old) {
                    const double d2 = DIM3 ? hdb_rdist3(xc, xs[tid + HDB_CT * q]) : hdb_rdist(xc, xs[tid + HDB_CT * q]);
                    if (d2 < m_old * m_old * 1.0000000000000009) {
                        const double mrd = fmax(lb, __dsqrt_rn(d2));
                        if (mrd < m_old) { mr[q] = mrd; cs[q] = cur; }
                    }
                }
                if (mr[q] < bv) { bv = mr[q]; bj = base + tid + HDB_CT * q; bs = cs[q]; bc = crj[q]; }   // ascending j: strict < keeps the lowest index
            }
        }
        hdb_warp_argmin(bv, bj, bs, bc);
        if (lane == 0) { rv[warp] = bv; rj[warp] = bj; rs[warp] = bs; rc[warp] = bc; }
        __syncthreads();
        if (warp == 0) {
            // second level over the NW warp results, then lanes 0..HDB_CS-1 store the CTA's record into the peers' shared memory
            bv = lane < NW ? rv[lane] : 1.7976931348623157e308;
            bj = lane < NW ? rj[lane] : 0x7ffffffe;
            bs = lane < NW ? rs[lane] : 0;
            bc = lane < NW ? rc[lane] : 0.0;
            hdb_warp_argmin(bv, bj, bs, bc);
            if (lane < HDB_CS) {
                const bool have = bv < 1.7976931348623157e308;
                HdbCand rec;
                rec.v = bv; rec.j = bj; rec.s = bs; rec.pad = 0.0;
                rec.x = have ? xs[bj - base] : make_double4(0, 0, 0, 0);
                rec.cd = bc;
                HdbCand* peer = cluster.map_shared_rank(recs, lane);
                peer[(i & 1) * HDB_CS + crank] = rec;          // double buffered by step parity
            }
        }
        cluster.sync();
        // winner of the HDB_CS candidates (lexicographic, like the in-CTA reduction): warp 0 picks it, everybody reads the result
        if (warp == 0) {
            const HdbCand* rr = recs + (i & 1) * HDB_CS;
            double wv = lane < HDB_CS ? rr[lane].v : 1.7976931348623157e308, wc = 0.0;
            int wj = lane < HDB_CS ? rr[lane].j : 0x7ffffffe, wr = lane;
            hdb_warp_argmin(wv, wj, wr, wc);                  // wr: the winning record (carried like the source index)
            if (lane == 0) { s_win = wr; }
        }
        __syncthreads();
        const HdbCand* win = recs + (i & 1) * HDB_CS + s_win;
        xc = win->x; cd = win->cd; cur = win->j;
        const int wj = cur;
        if (crank == 0 && tid == 0) {
            mst_src[(long long)f * (n - 1) + i] = win->s;
            mst_dst[(long long)f * (n - 1) + i] = wj;
            mst_w[(long long)f * (n - 1) + i] = win->v;
        }
        const int wl = wj - base;
        if (wl >= 0 && wl < per_cta && (wl & (HDB_CT - 1)) == tid) live &= ~(1u << (wl / HDB_CT));
    }
    cluster.sync();          // no CTA may exit while a peer can still store into its shared memory
}
