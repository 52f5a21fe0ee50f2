"""Drop-in for reference llckbdm/llckbdm.py (LLC-KBDM driver).

The ensemble of KBDM solves (79 % of the reference's wall time) runs on the GPU
through ``sampling.sample_kbdm_pooled``.  The clustering stage ("next" row f-1 of the scope table) is accelerated too: the
silhouette coefficients of all clusterings (one batched ``llck_silhouette_batched`` launch instead of M-1
``sklearn.metrics.silhouette_samples`` calls, llckbdm.py:291), the min-RMSE scoring of the cluster averages
(``llck_rmse_batched``, llckbdm.py:120), the two O(n^2) stages of the HDBSCAN fits (core distances and Prim's spanning
tree of the mutual-reachability graph, for all min_samples values in one launch each) and the rest of every fit (dendrogram,
condensed tree, excess-of-mass selection: ``llck_hdbscan_labels``, all fits on native host threads).

Clusterer contract.  The reference imports the un-vendored ``hdbscan`` package (llckbdm.py:3) and calls
``hdbscan.HDBSCAN(min_samples=k)`` with every other parameter at its default (min_cluster_size=5, euclidean, EOM,
allow_single_cluster=False).  ``hdbscan`` does NOT count the point itself among its ``min_samples`` neighbours;
``sklearn.cluster.HDBSCAN`` (derived from it, the stand-in here) does, so the reference's ``min_samples=k`` is
``sklearn.cluster.HDBSCAN(min_samples=k+1)`` and, on the device, the distance to the (k+1)-th nearest neighbour including the
point itself.  ``k`` is clipped to n-1 like ``hdbscan`` does.  The device path is the default whether or not ``hdbscan`` is
importable; it computes the EXACT minimum spanning tree, while ``hdbscan``'s default ``approx_min_span_tree=True`` Boruvka may
return a slightly different tree -- label parity is therefore pinned against sklearn's exact-tree implementation (bit-identical
edges, identical labels), not against the absent package.  ``CLUSTER_BACKEND = "host"`` runs the plain library fits
(``hdbscan`` if importable, else sklearn with k+1) instead.
"""
import logging
import os
from concurrent.futures import ThreadPoolExecutor

import attr
import numpy as np

from . import _native
from .ensemble import (group_labels, hdbscan_msts_device, score_rmse_device, silhouette_samples_device, solve_pooled,
                       to_device_complex, _require_cuda)
from .kbdm import check_finite, raise_for_status, resolve_m_l
from .metrics import calculate_freq_domain_rmse  # noqa: F401
from .min_rmse_kbdm import min_rmse_kbdm  # noqa: F401
from .sampling import filter_samples, sample_kbdm, sample_kbdm_pooled  # noqa: F401
from .sig_gen import _validate_parameters, gen_t_freq_arrays, multi_fid, multi_fid_batched_device  # noqa: F401

logger = logging.getLogger(__name__)

try:  # pragma: no cover - depends on the environment
    from hdbscan import HDBSCAN as _HDBSCAN
    _HDBSCAN_COUNTS_SELF = False
except Exception:  # noqa: BLE001
    from sklearn.cluster import HDBSCAN as _HDBSCAN
    _HDBSCAN_COUNTS_SELF = True

# "device": spanning trees on the GPU + native labelling (default); "host": the clustering library's own fits
CLUSTER_BACKEND = "device"
# worker threads / processes of the host halves (0 = all cores)
CLUSTER_JOBS = 0
MIN_CLUSTER_SIZE = 5           # hdbscan.HDBSCAN default, never changed by the reference
# wall-clock seconds of the stages of the last llc_kbdm call (profiling aid: solve+pool, spanning trees, labelling, silhouettes+averages, rmse)
LAST_STAGE_SECONDS = {}


@attr.s(auto_attribs=True)
class LlcKbdmResult:
    line_list: np.ndarray = np.array([])
    rmse: float = None
    silhouette: np.ndarray = np.array([])


@attr.s(auto_attribs=True)
class IterativeLlcKbdmResult:
    line_list: np.ndarray = np.array([])
    line_lists: np.ndarray = np.array([])
    rmse: float = None
    silhouettes: np.ndarray = np.array([])


@attr.s(auto_attribs=True)
class ClusteringResult:
    num_clusters: int = 0
    labels: np.ndarray = np.array([])
    clustered: np.ndarray = np.array([])
    non_clustered: np.ndarray = np.array([])
    summarized_line_list: np.ndarray = np.array([])
    clustered_silhouettes: np.ndarray = np.array([])


def _resolve_members(n_points, m_range, l, p):
    ms, ls = [], []
    for m in m_range:
        logger.info(f'Computing KBDM with m = {m}')
        mm, ll_ = resolve_m_l(n_points, m, l, p)
        ms.append(mm)
        ls.append(ll_)
    return ms, ls


def _llc_kbdm_device(sig_dev, dwell, ms, ls, p, q):
    """``llc_kbdm`` on an FID that is already on the device (torch complex128 CUDA tensor): batched solves, pooling + filter +
    feature transform, spanning trees, silhouettes and the min-RMSE selection all read it there; only line lists, labels and
    silhouettes (kilobytes) cross to the host."""
    torch = _require_cuda()
    import time
    t0 = time.perf_counter()
    LAST_STAGE_SECONDS.clear()
    if q > 0:
        logger.debug('Using Tikhonov Regularization with q=%f', q)
    # sampling + pooling + filter + feature transform (reference llckbdm.py:76-98) in one device pass
    samples, features, status = solve_pooled(sig_dev, ms, ls, p, q, dwell)
    LAST_STAGE_SECONDS["solve_pool"] = time.perf_counter() - t0
    for k, mm in enumerate(ms):
        raise_for_status(int(status[k]), mm)
    if len(samples) == 0:
        return LlcKbdmResult()
    # HDBSCAN for min_samples = 1..M-1 (reference llckbdm.py:104-116): all fits at once
    t1 = time.perf_counter()
    labelings = _fit_all(features, list(range(1, len(ms))))
    t2 = time.perf_counter()
    results = _results_from_labelings(samples, features, labelings)
    t3 = time.perf_counter()
    LAST_STAGE_SECONDS.update(hdbscan_fits=t2 - t1, silhouettes_averages=t3 - t2, pooled_points=int(len(samples)))
    if not results:
        return LlcKbdmResult()
    # min-RMSE selection over the cluster averages (reference llckbdm.py:120-124 -> min_rmse_kbdm.py:33-55) on the device
    cands = [r.summarized_line_list for r in results]
    for cand in cands:                    # the reference's multi_fid validates every row (sig_gen.py:140-169)
        for row in cand:
            _validate_parameters(*row)
    rows = np.array([len(c) for c in cands], dtype=np.int32)
    packed = np.zeros((len(cands), max(1, int(rows.max())), 4))
    for i, c in enumerate(cands):
        packed[i, :rows[i]] = c
    dev = sig_dev.device
    with torch.cuda.device(dev):
        rmses = score_rmse_device(sig_dev, dwell, torch.from_numpy(packed).to(dev), torch.from_numpy(rows).to(dev),
                                  filter_rows=False).cpu().numpy()
    for i, rmse in enumerate(rmses):
        logger.debug('RMSE for sample #%d: %f', i, rmse)
    k = int(np.argmin(rmses))
    LAST_STAGE_SECONDS["rmse_selection"] = time.perf_counter() - t3
    return LlcKbdmResult(line_list=cands[k], rmse=float(rmses[k]), silhouette=np.array(results[k].clustered_silhouettes))


def llc_kbdm(data, dwell, m_range, p=1, l=None, q=0.0):
    """Line-List-Clustering KBDM (reference llckbdm.py:41-141): sample KBDM over m_range (GPU, batched), pool and
    filter the line lists, cluster them with HDBSCAN for min_samples = 1..M-1, average each cluster, and return the
    clustering whose averaged line list has the smallest frequency-domain RMSE."""
    if len(m_range) < 2:
        raise ValueError("size of 'm_range' must be greater than 2.")
    torch = _require_cuda()
    ms, ls = _resolve_members(data.size, m_range, l, p)
    for mm in ms:
        check_finite(data, mm, p)
    dev = torch.device("cuda", torch.cuda.current_device())
    return _llc_kbdm_device(to_device_complex(np.asarray(data).ravel(), dev), dwell, ms, ls, p, q)


def iterative_llc_kbdm(data, dwell, m_range, p=1, l=None, q=0.0, max_iterations=5, silhouette_threshold=0.6):
    """Residual-iteration variant (reference llckbdm.py:144-199).  The FID goes to the device ONCE; the running estimate, the
    residual ``data - estimate`` (llckbdm.py:163), the model of the lines kept in each iteration (``multi_fid``, llckbdm.py:177)
    and the final RMSE (llckbdm.py:192) are all computed there -- nothing of FID length returns to the host."""
    if max_iterations < 1:
        raise ValueError("'max_iterations must be greater than zero")
    if len(m_range) < 2:
        raise ValueError("size of 'm_range' must be greater than 2.")
    torch = _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    n_points = len(data)
    ms, ls = _resolve_members(n_points, m_range, l, p)
    need = max(2 * mm + p - 1 for mm in ms)
    check_finite(data, (need - p + 1) // 2, p)
    data_dev = to_device_complex(np.asarray(data).ravel(), dev)
    estimate_dev = torch.zeros_like(data_dev)
    line_lists, silhouettes = [], []
    n_peaks = 0
    thresholds = np.linspace(silhouette_threshold, 0, max_iterations)
    for it in range(max_iterations):
        print(f'Iteration #{it}')
        resid_dev = data_dev - estimate_dev
        if not bool(torch.isfinite(torch.view_as_real(resid_dev[:need])).all().item()):
            raise ValueError("array must not contain infs or NaNs")
        res = _llc_kbdm_device(resid_dev, dwell, ms, ls, p, q)
        if len(res.line_list) == 0:
            logging.info('No more peaks can be fitted. Stopping.')
            break
        keep = np.nonzero(res.silhouette > np.percentile(res.silhouette, thresholds[it]))
        line_list = res.line_list[keep]
        if len(line_list):
            estimate_dev += multi_fid_batched_device([line_list], n_points, dwell, device=dev)[0]
        line_lists.append(line_list)
        silhouettes.append(res.silhouette[keep])
        n_peaks += len(line_list)
        print(f'Found {len(line_list)} peaks. Total: {n_peaks} peaks.')
    if not line_lists:
        return IterativeLlcKbdmResult()
    line_list = np.concatenate(line_lists)
    with torch.cuda.device(dev):
        rows = torch.tensor([len(line_list)], dtype=torch.int32, device=dev)
        cand = torch.from_numpy(np.ascontiguousarray(line_list.reshape(1, -1, 4))).to(dev) if len(line_list) else \
            torch.zeros((1, 1, 4), dtype=torch.float64, device=dev)
        rmse = float(score_rmse_device(estimate_dev, dwell, cand, rows, filter_rows=False).cpu().numpy()[0])
    ragged_ll = np.empty(len(line_lists), dtype=object)
    ragged_sil = np.empty(len(silhouettes), dtype=object)
    for i, (a, s_) in enumerate(zip(line_lists, silhouettes)):
        ragged_ll[i], ragged_sil[i] = a, s_
    return IterativeLlcKbdmResult(line_list=line_list, line_lists=ragged_ll, silhouettes=ragged_sil, rmse=rmse)


def _transform_line_lists(line_lists, dwell):
    """(A, T2, F, PH) -> (Re mu, Im mu, A, 0) with mu = exp(i dwell (2 pi F + i/T2))  (reference llckbdm.py:202-230;
    the phase feature is zeroed there)."""
    A, T2, F = line_lists[:, 0], line_lists[:, 1], line_lists[:, 2]
    mu = np.exp(1j * dwell * (2 * np.pi * F + 1j / T2))
    return np.column_stack((mu.real, mu.imag, A, line_lists[:, 3] * 0))


def _inverse_transform_line_lists(transformed_line_lists, dwell):
    """Inverse of ``_transform_line_lists`` (reference llckbdm.py:233-261)."""
    mu = transformed_line_lists[:, 0] + 1j * transformed_line_lists[:, 1]
    omega = -1j * np.log(mu) / dwell
    return np.column_stack((transformed_line_lists[:, 2], 1. / omega.imag, omega.real / (2 * np.pi),
                            transformed_line_lists[:, 3]))


def _library_min_samples(min_samples, n):
    """The reference's ``hdbscan.HDBSCAN(min_samples=k)`` in the installed clusterer's convention: ``hdbscan`` clips k to n-1 and
    does not count the point itself; sklearn counts it (k+1)."""
    k = max(1, min(int(min_samples), n - 1))
    return k + 1 if _HDBSCAN_COUNTS_SELF else k


def _fit_one(features, min_samples):
    """One library fit with the reference's parameters (llckbdm.py:280-283)."""
    kw = {"copy": True} if _HDBSCAN_COUNTS_SELF else {}
    model = _HDBSCAN(min_samples=_library_min_samples(min_samples, len(features)), **kw)
    model.fit(features)
    return np.asarray(model.labels_)


def _labels_from_msts(src, dst, w):
    """Labels of all fits from their spanning trees (int64 [F, n-1] x2, float64 [F, n-1]): the edges are sorted with the
    clusterer's own call (``np.argsort`` of the weights, sklearn ``_process_mst``) so that equal weights keep its order -- on
    threads, numpy releases the GIL -- and the trees are condensed and labelled by ``llck_hdbscan_labels`` on native threads."""
    lib = _native.load()
    F, ne = w.shape
    n = ne + 1
    jobs = CLUSTER_JOBS or (os.cpu_count() or 1)
    w = np.ascontiguousarray(w, dtype=np.float64)
    if F > 1 and jobs > 1:
        with ThreadPoolExecutor(max_workers=min(jobs, F)) as pool:
            order = np.stack(list(pool.map(np.argsort, w)))
    else:
        order = np.stack([np.argsort(row) for row in w])
    order = np.ascontiguousarray(order, dtype=np.int64)
    src = np.ascontiguousarray(src, dtype=np.int64)
    dst = np.ascontiguousarray(dst, dtype=np.int64)
    labels = np.empty((F, n), dtype=np.int32)
    rc = lib.llck_hdbscan_labels(src.ctypes.data, dst.ctypes.data, w.ctypes.data, order.ctypes.data, n, F, MIN_CLUSTER_SIZE,
                                 int(jobs), labels.ctypes.data)
    _native.check_rc(rc, "llck_hdbscan_labels")
    return [labels[f].astype(np.intp) for f in range(F)]


def _labels_from_mst(src, dst, w):
    """One fit (see ``_labels_from_msts``)."""
    return _labels_from_msts(np.asarray(src)[None, :], np.asarray(dst)[None, :], np.asarray(w, dtype=np.float64)[None, :])[0]


def _gpu_fit_supported(features, min_samples_list):
    """Sizes the device kernels handle (n <= 131072 points, k+1 <= 128 neighbours) and finite features; anything else keeps
    the plain library fits."""
    n = len(features)
    if CLUSTER_BACKEND != "device" or n < 2 or n > 131072 or not min_samples_list or min(min_samples_list) < 1:
        return False
    if min(max(min_samples_list), n - 1) + 1 > 128:
        return False
    return bool(np.isfinite(features).all())


def _fit_all(features, min_samples_list):
    """Labels of one HDBSCAN fit per min_samples value, in order (reference llckbdm.py:104-116 + :280-283).

    Device path (default): core distances and Prim's spanning tree of the mutual-reachability graph for all min_samples values at
    once (``ensemble.hdbscan_msts_device``; edges bit-identical to sklearn's exact Prim), then ``_labels_from_msts``.  Otherwise
    the library fits run on the host (in parallel processes when there are enough of them)."""
    n_fits = len(min_samples_list)
    n = len(features)
    if _gpu_fit_supported(features, min_samples_list):
        # neighbour counts INCLUDING the point itself (see the module docstring)
        import time
        ks = [min(int(ms), n - 1) + 1 for ms in min_samples_list]
        t0 = time.perf_counter()
        src, dst, w = hdbscan_msts_device(features, ks)
        t1 = time.perf_counter()
        labels = _labels_from_msts(src, dst, w)
        LAST_STAGE_SECONDS.update(spanning_trees=t1 - t0, labelling=time.perf_counter() - t1)
        return labels
    jobs = CLUSTER_JOBS or min(n_fits, os.cpu_count() or 1)
    if not (jobs > 1 and n_fits >= 8 and n >= 4000):
        return [_fit_one(features, ms) for ms in min_samples_list]
    from joblib import Parallel, delayed
    return Parallel(n_jobs=jobs, prefer="processes")(delayed(_fit_one)(features, ms) for ms in min_samples_list)


def _results_from_labelings(samples, features, labelings):
    """ClusteringResult per labeling with >= 1 cluster (reference llckbdm.py:285-321), silhouettes from one device launch.
    The points are grouped by label ONCE per labeling (``ensemble.group_labels``, a radix sort); the grouping feeds the
    silhouette kernel and replaces the reference's one ``labels == k`` scan per cluster."""
    keep = [np.asarray(lab) for lab in labelings if np.asarray(lab).max(initial=-1) >= 0]
    if not keep:
        return []
    groups = [group_labels(lab) for lab in keep]
    sil_all = silhouette_samples_device(features, keep, groups=groups)
    rates = samples.copy()
    rates[:, 1] = 1 / rates[:, 1]                      # T2 is averaged as a rate (reference llckbdm.py:345-349)
    results = []
    for labels, sil, (order, _cof, seg, values) in zip(keep, sil_all, groups):
        num_clusters = int(labels.max()) + 1             # HDBSCAN labels are 0..k-1 (+ -1 for noise)
        order = order.astype(np.intp)                    # ascending indices inside every group == np.nonzero(labels == k)
        first = int(np.searchsorted(values, 0))          # groups below are negative labels (noise)
        contiguous = len(values) - first == num_clusters
        if contiguous:
            bounds = seg[first:first + num_clusters + 1].astype(np.intp)
        else:                                            # label values with gaps: not produced by HDBSCAN
            bounds = np.searchsorted(labels[order], np.arange(num_clusters + 1))
        starts, counts = bounds[:-1], np.diff(bounds)
        clustered = [(order[bounds[k]:bounds[k + 1]],) for k in range(num_clusters)]
        if np.all(counts > 0):
            cluster_sil = np.add.reduceat(sil[order], starts) / counts
            summary = np.add.reduceat(rates[order], starts, axis=0) / counts[:, None]
            summary[:, 1] = 1 / summary[:, 1]
        else:                                            # keep the reference's per-cluster semantics
            cluster_sil = np.array([np.average(sil[c]) for c in clustered])
            summary = _summarize_clusters(samples=samples, clusters=clustered)
        noise = order[:seg[first]] if first > 0 and values[0] == -1 and first == 1 else np.nonzero(labels == -1)[0]
        results.append(ClusteringResult(num_clusters=num_clusters, labels=labels, clustered=clustered, non_clustered=(noise,),
                                        summarized_line_list=summary, clustered_silhouettes=np.asarray(cluster_sil)))
    return results


def _cluster_line_lists(samples, transformed_samples, min_samples):
    """One HDBSCAN fit + per-cluster mean silhouette + cluster averages (reference llckbdm.py:264-321)."""
    labels = _fit_one(transformed_samples, min_samples)
    res = _results_from_labelings(samples, transformed_samples, [labels])
    return res[0] if res else ClusteringResult(num_clusters=0, labels=labels)


def _summarize_clusters(samples, clusters, summarizer=np.average):
    """Average each cluster; T2 is averaged as a rate 1/T2 and inverted back (reference llckbdm.py:324-353)."""
    if summarizer is None:
        summarizer = np.average
    out = []
    for members in clusters:
        block = samples[members].copy()
        block[:, 1] = 1 / block[:, 1]
        row = summarizer(block, axis=0)
        row[1] = 1 / row[1]
        out.append(row)
    return np.array(out)
