"""Drop-in for reference llckbdm/llckbdm.py (LLC-KBDM driver).

The ensemble of KBDM solves (79 % of the reference's wall time) runs on the GPU
through ``sampling.sample_kbdm``.  Of the clustering stage ("next" row f-1 of the scope table) the two O(n^2)/O(M n K) loops
are on the GPU too: the silhouette coefficients of all clusterings (one batched ``llck_silhouette_batched`` launch instead of
M-1 ``sklearn.metrics.silhouette_samples`` calls, llckbdm.py:291) and the min-RMSE scoring of the cluster averages
(``llck_rmse_batched``, llckbdm.py:120), and so are the two O(n^2) stages of the HDBSCAN fits (core distances and Prim's spanning
tree of the mutual-reachability graph, for all min_samples values in one launch each, edge-for-edge identical to the host
clusterer's); the tree condensation / EOM selection stays the clusterer's own host code, so labels agree exactly.

Clusterer: the reference imports the un-vendored ``hdbscan`` package (llckbdm.py:3).  If it is
importable it is used; otherwise ``sklearn.cluster.HDBSCAN`` with the same defaults
(min_cluster_size=5, euclidean, EOM, allow_single_cluster=False) stands in.
"""
import logging
import os

import numpy as np

from .metrics import calculate_freq_domain_rmse
from .min_rmse_kbdm import min_rmse_kbdm
from .sampling import filter_samples, sample_kbdm, sample_kbdm_pooled  # noqa: F401
from .ensemble import hdbscan_msts_device, silhouette_samples_device
from .sig_gen import gen_t_freq_arrays, multi_fid, multi_fid_batched_device  # noqa: F401

logger = logging.getLogger(__name__)

try:  # pragma: no cover - depends on the environment
    from hdbscan import HDBSCAN as _HDBSCAN
except Exception:  # noqa: BLE001
    from sklearn.cluster import HDBSCAN as _HDBSCAN


class LlcKbdmResult:
    def __init__(self, line_list=None, rmse=None, silhouette=None):
        self.line_list = np.array([]) if line_list is None else line_list
        self.rmse = rmse
        self.silhouette = np.array([]) if silhouette is None else silhouette


class IterativeLlcKbdmResult:
    def __init__(self, line_list=None, line_lists=None, rmse=None, silhouettes=None):
        self.line_list = np.array([]) if line_list is None else line_list
        self.line_lists = np.array([]) if line_lists is None else line_lists
        self.rmse = rmse
        self.silhouettes = np.array([]) if silhouettes is None else silhouettes


class ClusteringResult:
    def __init__(self, num_clusters=0, labels=None, clustered=None, non_clustered=None, summarized_line_list=None,
                 clustered_silhouettes=None):
        self.num_clusters = num_clusters
        self.labels = np.array([]) if labels is None else labels
        self.clustered = [] if clustered is None else clustered
        self.non_clustered = np.array([]) if non_clustered is None else non_clustered
        self.summarized_line_list = np.array([]) if summarized_line_list is None else summarized_line_list
        self.clustered_silhouettes = np.array([]) if clustered_silhouettes is None else clustered_silhouettes


def llc_kbdm(data, dwell, m_range, p=1, l=None, q=0.0):
    """Line-List-Clustering KBDM (reference llckbdm.py:41-141): sample KBDM over m_range (GPU, batched), pool and
    filter the line lists, cluster them with HDBSCAN for min_samples = 1..M-1, average each cluster, and return the
    clustering whose averaged line list has the smallest frequency-domain RMSE."""
    if len(m_range) < 2:
        raise ValueError("size of 'm_range' must be greater than 2.")
    # sampling + pooling + filter + feature transform (reference llckbdm.py:76-98) in one device pass
    samples, features = sample_kbdm_pooled(data=data, dwell=dwell, m_range=m_range, p=p, l=l, q=q)
    if len(samples) == 0:
        return LlcKbdmResult()
    n_members = len(m_range)
    # HDBSCAN for min_samples = 1..M-1 (reference llckbdm.py:104-116): the fits are independent -> spread over the host cores;
    # the silhouettes of all clusterings are then ONE batched device launch instead of M-1 O(n^2) sklearn calls.
    labelings = _fit_all(features, list(range(1, n_members)))
    results = _results_from_labelings(samples, features, labelings)
    best = min_rmse_kbdm(data=data, dwell=dwell, samples=[r.summarized_line_list for r in results])
    if best is None:
        return LlcKbdmResult()
    return LlcKbdmResult(line_list=best.line_list, rmse=best.min_rmse,
                         silhouette=np.array(results[best.min_index].clustered_silhouettes))


def iterative_llc_kbdm(data, dwell, m_range, p=1, l=None, q=0.0, max_iterations=5, silhouette_threshold=0.6):
    """Residual-iteration variant (reference llckbdm.py:144-199)."""
    if max_iterations < 1:
        raise ValueError("'max_iterations must be greater than zero")
    estimate = np.zeros_like(data)
    line_lists, silhouettes = [], []
    t_array, _ = gen_t_freq_arrays(N=len(data), dwell=dwell)
    n_peaks = 0
    thresholds = np.linspace(silhouette_threshold, 0, max_iterations)
    for it in range(max_iterations):
        print(f'Iteration #{it}')
        res = llc_kbdm(data=data - estimate, dwell=dwell, m_range=m_range, p=p, l=l, q=q)
        if len(res.line_list) == 0:
            logging.info('No more peaks can be fitted. Stopping.')
            break
        keep = np.nonzero(res.silhouette > np.percentile(res.silhouette, thresholds[it]))
        line_list = res.line_list[keep]
        # residual model on the device (reference llckbdm.py:177 calls sig_gen.multi_fid on the host)
        estimate = estimate + multi_fid_batched_device([line_list], len(data), dwell)[0].cpu().numpy()
        line_lists.append(line_list)
        silhouettes.append(res.silhouette[keep])
        n_peaks += len(line_list)
        print(f'Found {len(line_list)} peaks. Total: {n_peaks} peaks.')
    if not line_lists:
        return IterativeLlcKbdmResult()
    line_list = np.concatenate(line_lists)
    rmse = calculate_freq_domain_rmse(data=estimate, params_est=line_list, dwell=dwell)
    ragged_ll = np.empty(len(line_lists), dtype=object)
    ragged_sil = np.empty(len(silhouettes), dtype=object)
    for i, (a, s) in enumerate(zip(line_lists, silhouettes)):
        ragged_ll[i], ragged_sil[i] = a, s
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


def _fit_one(features, min_samples):
    model = _HDBSCAN(min_samples=min_samples)
    model.fit(features)
    return np.asarray(model.labels_)


def _labels_from_mst(src, dst, w):
    """The host half of sklearn.cluster.HDBSCAN.fit after the spanning tree: sort the edges, single-linkage tree, condensed tree,
    EOM selection (``_process_mst`` + ``tree_to_labels`` with the estimator's defaults)."""
    from sklearn.cluster._hdbscan._linkage import MST_edge_dtype, make_single_linkage
    from sklearn.cluster._hdbscan._tree import tree_to_labels
    mst = np.empty(len(w), dtype=MST_edge_dtype)
    mst["current_node"], mst["next_node"], mst["distance"] = src, dst, w
    mst = mst[np.argsort(mst["distance"])]
    ref = _HDBSCAN()
    labels, _ = tree_to_labels(make_single_linkage(mst), ref.min_cluster_size, ref.cluster_selection_method,
                               ref.allow_single_cluster, ref.cluster_selection_epsilon, ref.max_cluster_size)
    return np.asarray(labels)


def _gpu_fit_supported(features, min_samples_list):
    """The device spanning trees reproduce sklearn's Euclidean Prim path edge for edge; anything else (the external ``hdbscan``
    package, non-finite features, sizes outside the kernels' limits) keeps the plain host fits."""
    if os.environ.get("LLCK_GPU_MST", "1") == "0" or not _HDBSCAN.__module__.startswith("sklearn."):
        return False
    n = len(features)
    if n < 2 or n > 131072 or not min_samples_list or max(min_samples_list) > min(128, n) or min(min_samples_list) < 1:
        return False
    try:
        from sklearn.cluster._hdbscan._linkage import MST_edge_dtype, make_single_linkage  # noqa: F401
        from sklearn.cluster._hdbscan._tree import tree_to_labels  # noqa: F401
    except Exception:  # noqa: BLE001
        return False
    return bool(np.isfinite(features).all())


def _fit_all(features, min_samples_list):
    """Labels of one HDBSCAN fit per min_samples value, in order (reference llckbdm.py:104-116 + :280-283).

    With sklearn's HDBSCAN as the clusterer the two O(n^2) stages of every fit -- k-nearest-neighbour core distances and Prim's
    spanning tree of the mutual-reachability graph -- run on the device for all min_samples values at once
    (``ensemble.hdbscan_msts_device``, edge lists identical to the host's), and only the O(n log n) tree condensation runs on the
    host, spread over worker processes.  Otherwise the fits run on the host (in parallel processes when there are enough of them;
    LLCK_CLUSTER_JOBS overrides the worker count, 1 = serial)."""
    n_fits = len(min_samples_list)
    jobs = int(os.environ.get("LLCK_CLUSTER_JOBS", "0")) or min(n_fits, os.cpu_count() or 1)
    parallel = jobs > 1 and n_fits >= 8 and len(features) >= 4000
    if _gpu_fit_supported(features, min_samples_list):
        src, dst, w = hdbscan_msts_device(features, min_samples_list)
        if not parallel:
            return [_labels_from_mst(src[f], dst[f], w[f]) for f in range(n_fits)]
        from joblib import Parallel, delayed
        return Parallel(n_jobs=jobs, prefer="processes")(delayed(_labels_from_mst)(src[f], dst[f], w[f]) for f in range(n_fits))
    if not parallel:
        return [_fit_one(features, ms) for ms in min_samples_list]
    from joblib import Parallel, delayed
    return Parallel(n_jobs=jobs, prefer="processes")(delayed(_fit_one)(features, ms) for ms in min_samples_list)


def _results_from_labelings(samples, features, labelings):
    """ClusteringResult per labeling with >= 1 cluster (reference llckbdm.py:285-321), silhouettes from one device launch.
    Clusters are grouped by one stable sort per labeling instead of one ``labels == k`` scan per cluster."""
    keep = [np.asarray(lab) for lab in labelings if np.asarray(lab).max(initial=-1) >= 0]
    if not keep:
        return []
    sil_all = silhouette_samples_device(features, keep)
    rates = samples.copy()
    rates[:, 1] = 1 / rates[:, 1]                      # T2 is averaged as a rate (reference llckbdm.py:345-349)
    results = []
    for labels, sil in zip(keep, sil_all):
        num_clusters = int(labels.max()) + 1             # HDBSCAN labels are 0..k-1 (+ -1 for noise)
        order = np.argsort(labels, kind="stable")        # ascending indices inside every cluster == np.nonzero(labels == k)
        bounds = np.searchsorted(labels[order], np.arange(num_clusters + 1))
        starts, counts = bounds[:-1], np.diff(bounds)
        clustered = [(order[bounds[k]:bounds[k + 1]],) for k in range(num_clusters)]
        if np.all(counts > 0):
            cluster_sil = np.add.reduceat(sil[order], starts) / counts
            summary = np.add.reduceat(rates[order], starts, axis=0) / counts[:, None]
            summary[:, 1] = 1 / summary[:, 1]
        else:                                            # not produced by HDBSCAN; keep the reference's per-cluster semantics
            cluster_sil = np.array([np.average(sil[c]) for c in clustered])
            summary = _summarize_clusters(samples=samples, clusters=clustered)
        results.append(ClusteringResult(num_clusters=num_clusters, labels=labels, clustered=clustered,
                                        non_clustered=np.nonzero(labels == -1),
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
