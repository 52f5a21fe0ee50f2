// Host half of the HDBSCAN fits of reference llckbdm/llckbdm.py:104-116 (one fit per min_samples value on the same points):
// from a spanning tree of the mutual-reachability graph (device: hdbscan_mst.cuh) to flat cluster labels --
// single-linkage dendrogram, condensed tree (min_cluster_size), cluster stabilities, excess-of-mass selection, labelling --
// with the clusterer's defaults (min_cluster_size 5, "eom", allow_single_cluster False, cluster_selection_epsilon 0, no
// max_cluster_size).  The stand-in clusterer (sklearn.cluster.HDBSCAN, derived from the `hdbscan` package the reference imports)
// does this in interpreted / list-based code, one fit at a time; here all fits of an llc_kbdm call run on host threads, O(n) each.
//
// Label-for-label agreement needs the SAME floating-point sums: the condensed-tree rows are produced in the clusterer's order
// (breadth-first over the dendrogram, left before right) and the stabilities are accumulated row by row in that order.
#pragma once
#include <stdint.h>
#include <limits>
#include <thread>
#include <vector>

namespace llck_hdb {

struct CondRow { int64_t parent, child; double lambda; int64_t size; };

// labels of one fit.  src/dst/w: the n-1 spanning-tree edges; order: permutation that sorts them by weight (may be null = sorted)
static void labels_one_fit(const int64_t* src, const int64_t* dst, const double* w, const int64_t* order, int64_t n,
                           int64_t min_cluster_size, int32_t* labels) {
    const int64_t ne = n - 1;
    if (n <= 0) return;
    if (ne <= 0) { labels[0] = -1; return; }
    // ---- single-linkage dendrogram: union-find over the edges in ascending weight; node n+i is created by edge i ----
    std::vector<int64_t> uf_parent(2 * n - 1, -1), uf_size(2 * n - 1, 0);
    for (int64_t i = 0; i < n; ++i) uf_size[i] = 1;
    std::vector<int64_t> left(ne), right(ne), csize(ne);
    std::vector<double> value(ne);
    auto find = [&](int64_t x) {
        int64_t r = x;
        while (uf_parent[r] != -1) r = uf_parent[r];
        while (uf_parent[x] != -1 && uf_parent[x] != r) { int64_t nx = uf_parent[x]; uf_parent[x] = r; x = nx; }
        return r;
    };
    int64_t next_label = n;
    for (int64_t i = 0; i < ne; ++i) {
        const int64_t e = order ? order[i] : i;
        const int64_t a = find(src[e]), b = find(dst[e]);
        left[i] = a; right[i] = b; value[i] = w[e]; csize[i] = uf_size[a] + uf_size[b];
        uf_parent[a] = next_label; uf_parent[b] = next_label; uf_size[next_label] = csize[i];
        ++next_label;
    }
    // ---- condensed tree ----
    const int64_t root = 2 * ne;
    std::vector<int64_t> bfs; bfs.reserve(2 * n - 1);
    auto bfs_from = [&](int64_t start, std::vector<int64_t>& out) {      // breadth-first, left before right
        out.clear();
        out.push_back(start);
        for (size_t h = 0; h < out.size(); ++h) {
            const int64_t x = out[h];
            if (x >= n) { out.push_back(left[x - n]); out.push_back(right[x - n]); }
        }
    };
    bfs_from(root, bfs);
    std::vector<int64_t> relabel(root + 1, 0);
    std::vector<uint8_t> ignore(root + 1, 0);
    std::vector<CondRow> rows; rows.reserve(n + 64);
    std::vector<int64_t> sub;
    relabel[root] = n;
    int64_t next_cluster = n + 1;
    const double INF = std::numeric_limits<double>::infinity();
    auto spill = [&](int64_t from, int64_t parent_label, double lambda) {      // every point below `from` leaves the parent at lambda
        bfs_from(from, sub);
        for (int64_t s : sub) {
            if (s < n) rows.push_back({parent_label, s, lambda, 1});
            ignore[s] = 1;
        }
    };
    for (int64_t node : bfs) {
        if (ignore[node] || node < n) continue;
        const int64_t l = left[node - n], r = right[node - n];
        const double d = value[node - n];
        const double lambda = d > 0.0 ? 1.0 / d : INF;
        const int64_t lc = l >= n ? csize[l - n] : 1, rc = r >= n ? csize[r - n] : 1;
        if (lc >= min_cluster_size && rc >= min_cluster_size) {
            relabel[l] = next_cluster++;
            rows.push_back({relabel[node], relabel[l], lambda, lc});
            relabel[r] = next_cluster++;
            rows.push_back({relabel[node], relabel[r], lambda, rc});
        } else if (lc < min_cluster_size && rc < min_cluster_size) {
            spill(l, relabel[node], lambda);
            spill(r, relabel[node], lambda);
        } else if (lc < min_cluster_size) {
            relabel[r] = relabel[node];
            spill(l, relabel[node], lambda);
        } else {
            relabel[l] = relabel[node];
            spill(r, relabel[node], lambda);
        }
    }
    // ---- stabilities (row order = summation order) ----
    const int64_t nclusters = next_cluster - n;          // cluster ids n .. next_cluster-1, id order is a topological order
    std::vector<double> births(next_cluster, std::numeric_limits<double>::quiet_NaN());
    for (const CondRow& c : rows) births[c.child] = c.lambda;
    births[n] = 0.0;
    std::vector<double> stability(nclusters, 0.0);
    for (const CondRow& c : rows) stability[c.parent - n] += (c.lambda - births[c.parent]) * (double)c.size;
    // ---- excess of mass: cluster tree = rows with size > 1; every cluster has no or two child clusters ----
    std::vector<int64_t> kid0(nclusters, -1), kid1(nclusters, -1), up(nclusters, -1);
    for (const CondRow& c : rows) {
        if (c.size > 1) {
            const int64_t pi = c.parent - n, ci = c.child - n;
            if (kid0[pi] < 0) kid0[pi] = ci; else kid1[pi] = ci;
            up[ci] = pi;
        }
    }
    std::vector<uint8_t> is_cluster(nclusters, 1);
    is_cluster[0] = 0;                                   // allow_single_cluster = False: the root is never selected
    std::vector<int64_t> stack;
    for (int64_t c = nclusters - 1; c >= 1; --c) {
        double subtree = 0.0;
        if (kid0[c] >= 0) subtree = stability[kid0[c]];
        if (kid1[c] >= 0) subtree = subtree + stability[kid1[c]];
        if (subtree > stability[c]) {
            is_cluster[c] = 0;
            stability[c] = subtree;
        } else {
            stack.clear();
            if (kid0[c] >= 0) stack.push_back(kid0[c]);
            if (kid1[c] >= 0) stack.push_back(kid1[c]);
            while (!stack.empty()) {
                const int64_t x = stack.back(); stack.pop_back();
                is_cluster[x] = 0;
                if (kid0[x] >= 0) stack.push_back(kid0[x]);
                if (kid1[x] >= 0) stack.push_back(kid1[x]);
            }
        }
    }
    // ---- labelling: a point belongs to the nearest selected cluster above it; reaching the root means noise ----
    std::vector<int32_t> label_of(nclusters, -1);
    int32_t nl = 0;
    for (int64_t c = 0; c < nclusters; ++c) if (is_cluster[c]) label_of[c] = nl++;
    for (int64_t c = 1; c < nclusters; ++c)              // ids are topologically ordered: parents come first
        if (!is_cluster[c]) label_of[c] = (up[c] >= 0) ? label_of[up[c]] : -1;
    for (int64_t i = 0; i < n; ++i) labels[i] = -1;
    for (const CondRow& c : rows)
        if (c.size == 1 && c.child < n) labels[c.child] = label_of[c.parent - n];
}

static void labels_all_fits(const int64_t* src, const int64_t* dst, const double* w, const int64_t* order, int64_t n, int nfits,
                            int64_t min_cluster_size, int nthreads, int32_t* labels) {
    const int64_t ne = n > 0 ? n - 1 : 0;
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nfits) nthreads = nfits;
    auto work = [&](int t) {
        for (int f = t; f < nfits; f += nthreads)
            labels_one_fit(src + (int64_t)f * ne, dst + (int64_t)f * ne, w + (int64_t)f * ne, order ? order + (int64_t)f * ne : nullptr,
                           n, min_cluster_size, labels + (int64_t)f * n);
    };
    if (nthreads == 1) { work(0); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) pool.emplace_back(work, t);
    for (auto& th : pool) th.join();
}

}  // namespace llck_hdb
