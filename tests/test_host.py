"""CPU tests of the host logic and of the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from llckbdm_b200 import _native
from llckbdm_b200.ensemble import flatten_signals, flops_per_solve, lpt_shards
from llckbdm_b200.kbdm import raise_for_status, resolve_m_l
from llckbdm_b200.sampling import filter_samples

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(native_lib):
    header = open(os.path.join(ROOT, "include", "llck.h")).read()
    declared = set(re.findall(r"\b(llck_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_native.SYMBOLS)
    for sym in declared:
        assert getattr(native_lib, sym) is not None
    assert native_lib.llck_version() == 100


def test_pure_abi_functions(native_lib):
    assert native_lib.llck_leading_dim(1) == 64
    assert native_lib.llck_leading_dim(64) == 64
    assert native_lib.llck_leading_dim(700) == 704
    assert native_lib.llck_leading_dim(1024) == 1024
    w1 = native_lib.llck_workspace_bytes(1, 1024, 0)
    w2 = native_lib.llck_workspace_bytes(2, 1024, 0)
    # 6 pipeline matrices + 5 for the divide-and-conquer SVD of the bidiagonal, plus panel/vector scratch
    assert 11 * 1024 * 1024 * 16 <= w1 < 11 * 1024 * 1024 * 16 + (1 << 23)
    assert w2 > w1 and native_lib.llck_workspace_bytes(0, 1024, 0) == 0
    assert native_lib.llck_workspace_bytes(1, 1024, _native.FLAG_DEBUG_KEEP) > w1 + 8 * 1024 * 1024 * 16 - (1 << 20)
    assert native_lib.llck_debug_offset(3, 128, 1) - native_lib.llck_debug_offset(3, 128, 0) == 3 * 128 * 128 * 16


def test_bad_arguments_are_rejected_without_a_gpu(native_lib):
    m = (ctypes.c_int32 * 1)(8)
    l = (ctypes.c_int32 * 1)(9)          # l > m
    off = (ctypes.c_int64 * 1)(0)
    dummy = ctypes.c_void_p(16)
    rc = native_lib.llck_kbdm_batched(dummy, off, m, l, 1, 0.0, 5e-4, 1, dummy, 64, None, None, 0, dummy, 16, dummy, dummy,
                                      dummy, 1 << 30, 0, None, None)
    assert rc == 1


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from llckbdm_b200.kbdm import kbdm
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kbdm(np.ones(64, dtype=complex), 5e-4, m=8)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "llckbdm_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn


def test_resolve_m_l_matches_reference_validation():
    assert resolve_m_l(2048, None, 30, 1) == (30, 30)
    assert resolve_m_l(2048, 30, None, 1) == (30, 30)
    assert resolve_m_l(2048, 1024, 1024, 1) == (1024, 1024)
    with pytest.raises(ValueError, match="l or m must be specified"):
        resolve_m_l(2048, None, None, 1)
    with pytest.raises(ValueError, match="l can't be greater than m"):
        resolve_m_l(2048, 20, 30, 1)
    with pytest.raises(ValueError, match=r"m or l can't be greater than \(n \+ 1 - p\)/2."):
        resolve_m_l(2048, 1025, None, 1)
    with pytest.raises(ValueError):
        resolve_m_l(2048, 1024, None, 2)


def test_non_finite_input_raises_like_scipy_check_finite():
    from llckbdm_b200.kbdm import check_finite
    c = np.ones(64, dtype=complex)
    check_finite(c, 8, 1)
    c[15] = np.nan                      # c[0 .. 2m+p-2] = c[0..15] is what the Hankel matrices use
    with pytest.raises(ValueError, match="array must not contain infs or NaNs"):
        check_finite(c, 8, 1)
    c[15], c[16] = 1.0, np.inf          # beyond the used slice: the reference does not look at it
    check_finite(c, 8, 1)


def test_status_mapping():
    raise_for_status(0)
    for st in (1, 2, 3, 4):
        with pytest.raises(np.linalg.LinAlgError):
            raise_for_status(st, 10)


def test_filter_samples():
    x = np.array([[1.0, 0.1, 5.0, 0.0], [1e-7, 0.1, 5.0, 0.0], [1.0, -0.1, 5.0, 0.0], [1.0, np.inf, 1.0, 0.0], [np.nan, 1.0, 1.0, 0.0]])
    out = filter_samples(x)
    assert out.shape == (2, 4) and out[0, 0] == 1.0 and np.isinf(out[1, 1])
    empty = np.array([])
    assert np.array_equal(filter_samples(empty), empty)


def test_lpt_shards_balance_and_cover():
    costs = [flops_per_solve(700 + round(k * 324 / 99), 700 + round(k * 324 / 99)) for k in range(100)]
    for world in (1, 2, 4, 8):
        shards = lpt_shards(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(100))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) / (sum(loads) / world) < 1.05


def test_flatten_signals():
    a = np.arange(4) + 0j
    flat, off = flatten_signals(a, 3)
    assert flat.shape == (4,) and list(off) == [0, 0, 0]
    flat, off = flatten_signals([a, a[:2], a], 3)
    assert flat.shape == (10,) and list(off) == [0, 4, 6]
    with pytest.raises(ValueError):
        flatten_signals([a], 2)


def test_sig_gen_and_metrics():
    from llckbdm_b200 import sig_gen
    from llckbdm_b200.metrics import calculate_freq_domain_rmse
    from oracle.kbdm_oracle import BRAIN_SIM_PARAMS, brain_sim
    t, f = sig_gen.gen_t_freq_arrays(2048, 5e-4)
    assert len(t) == 2048 and len(f) == 2048
    c = sig_gen.multi_fid(np.linspace(0, 5e-4 * 2048, 2048, endpoint=False), BRAIN_SIM_PARAMS)
    assert np.abs(c - brain_sim(2048, 0.0, 0)).max() < 1e-13
    assert calculate_freq_domain_rmse(c, BRAIN_SIM_PARAMS, 5e-4) < 1e-10
    with pytest.raises(ValueError, match="T2 must be positive"):
        sig_gen.fid(t, 1.0, 0.0, 1.0)
    with pytest.raises(ValueError, match="Amplitude can't be negative"):
        sig_gen.fid(t, -1.0, 1.0, 1.0)
    peak = sig_gen.lorentzian_peak(f, 1.0, 0.1, 100.0)
    assert abs(f[np.argmax(peak.real)] - 100.0) < 1.0


def test_labels_from_mst_reproduces_hdbscan_fit():
    """Host half of the device-accelerated HDBSCAN fits: feeding the clusterer's OWN spanning tree (sklearn's Prim) through
    llckbdm._labels_from_mst must reproduce HDBSCAN(min_samples=k).fit(X).labels_ exactly (the GPU test checks that the device
    spanning trees equal sklearn's edge for edge)."""
    from sklearn.cluster._hdbscan._linkage import mst_from_data_matrix
    from sklearn.metrics import DistanceMetric
    from sklearn.neighbors import NearestNeighbors
    from llckbdm_b200 import llckbdm as L
    rng = np.random.default_rng(1)
    cent = rng.uniform(-1, 1, (8, 3))
    X = np.concatenate([np.repeat(cent, 10, axis=0) + 1e-5 * rng.standard_normal((80, 3)), rng.uniform(-1, 1, (300, 3))])
    X = np.column_stack([X, np.zeros(len(X))])
    for k in (1, 4, 11):
        cd = np.ascontiguousarray(NearestNeighbors(n_neighbors=k, algorithm="kd_tree").fit(X).kneighbors(X, k)[0][:, -1])
        mst = mst_from_data_matrix(np.asarray(X, order="C"), cd, DistanceMetric.get_metric("euclidean"), 1.0)
        got = L._labels_from_mst(mst["current_node"], mst["next_node"], mst["distance"])
        assert np.array_equal(got, L._fit_one(X, k))


def test_cluster_grouping_matches_reference_semantics(monkeypatch):
    """_results_from_labelings (one stable sort per labeling) == the reference's per-cluster np.nonzero / np.average loops
    (llckbdm.py:297-313, 324-353); the device silhouettes are replaced by a fixed array here."""
    from llckbdm_b200 import llckbdm as L
    rng = np.random.default_rng(0)
    n = 3000
    samples = np.column_stack([rng.random(n) + 0.1, rng.random(n) * 0.1 + 0.01, rng.uniform(-500, 500, n), rng.uniform(-1, 1, n)])
    labels = rng.integers(-1, 25, n)
    sil = rng.uniform(-1, 1, n)
    monkeypatch.setattr(L, "silhouette_samples_device", lambda feats, labelings: np.array([sil for _ in labelings]))
    res = L._results_from_labelings(samples, samples, [labels, np.full(n, -1)])
    assert len(res) == 1 and res[0].num_clusters == 25
    clusters = [np.nonzero(labels == k) for k in range(25)]
    assert all(np.array_equal(a[0], b[0]) for a, b in zip(res[0].clustered, clusters))
    assert np.allclose(res[0].summarized_line_list, L._summarize_clusters(samples, clusters), rtol=1e-13, atol=0)
    assert np.allclose(res[0].clustered_silhouettes, [np.average(sil[c]) for c in clusters], rtol=1e-13, atol=1e-16)
