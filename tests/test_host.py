"""CPU tests of the host logic and of the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from llckbdm_b200 import _native
from llckbdm_b200.ensemble import flatten_signals, flops_per_solve, lpt_shards, plan_chunks
from llckbdm_b200.kbdm import raise_for_status, resolve_m_l
from llckbdm_b200.sampling import filter_samples

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(native_lib):
    header = open(os.path.join(ROOT, "include", "llck.h")).read()
    declared = set(re.findall(r"^(?:int|size_t)\s+(llck_[a-z_0-9]+)\s*\(", header, flags=re.M))
    assert declared == set(_native.SYMBOLS)
    for sym in declared:
        assert getattr(native_lib, sym) is not None
    assert native_lib.llck_version() == 200


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
    off = (ctypes.c_int64 * 1)(0)
    dummy = ctypes.c_void_p(16)

    def call(m, l, n, p=1, opts=None, ws=1 << 40):
        return native_lib.llck_kbdm_batched(dummy, off, (ctypes.c_int64 * 1)(n), (ctypes.c_int32 * 1)(m), (ctypes.c_int32 * 1)(l),
                                            p, 0.0, 5e-4, 1, dummy, 4 * l, None, None, 0, dummy, m, dummy, dummy,
                                            dummy, ws, 0, opts, None, None)

    assert call(8, 9, 64) == _native.E_BADARG                # l > m
    assert call(8, 8, 14) == _native.E_SHORT_SIGNAL          # needs 2m + p - 1 = 16 points
    assert call(8, 8, 16, p=2) == _native.E_SHORT_SIGNAL     # p = 2 needs 17
    assert call(2100, 2100, 5000) == _native.E_TOO_LARGE     # m above LLCK_M_MAX
    assert call(64, 64, 128, ws=1024) == _native.E_WORKSPACE
    bad = _native.Options(svd_mode=7)
    assert call(8, 8, 64, opts=ctypes.byref(bad)) == _native.E_BADARG
    bad = _native.Options(cluster_size=3)
    assert call(8, 8, 64, opts=ctypes.byref(bad)) == _native.E_BADARG
    with pytest.raises(ValueError, match="shorter than the 2m"):
        _native.check_rc(_native.E_SHORT_SIGNAL, "llck_kbdm_batched")


def test_library_reads_no_environment():
    """The C-ABI library has no hidden switches: no getenv anywhere in the CUDA sources (tuning knobs are llck_options fields)."""
    csrc = os.path.join(ROOT, "llckbdm_b200", "csrc")
    for f in os.listdir(csrc):
        assert "getenv" not in open(os.path.join(csrc, f)).read(), f


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


def test_plan_chunks_whole_waves():
    assert plan_chunks(100, 780) == [100]                    # fits: one chunk
    assert plan_chunks(1184, 780) == [592, 592]              # two equal chunks of 4 waves, not 780 + 404
    assert plan_chunks(148, 780) == [148] and plan_chunks(296, 780) == [296]
    assert plan_chunks(149, 780) == [148, 1]                 # a nearly empty second wave becomes a cluster-mode chunk
    assert plan_chunks(200, 780) == [148, 52]
    assert plan_chunks(240, 780) == [240]                    # the last wave is more than half full: leave it
    for M, cap in ((10000, 780), (65536, 3000), (1185, 780), (5000, 200)):
        sizes = plan_chunks(M, cap)
        assert sum(sizes) == M and max(sizes) <= cap and min(sizes) > 0
        assert all(s % 148 == 0 for s in sizes[:-2])
    assert plan_chunks(300, 100) == [100, 100, 100]          # memory cap below one wave
    assert plan_chunks(7, 3) == [3, 2, 2]
    assert plan_chunks(0, 10) == []


def test_two_stream_split_only_when_it_shortens_the_critical_path():
    from llckbdm_b200.ensemble import two_stream_split
    m = np.array([700 + round(k * 324 / 99) for k in range(100)])          # config C2: 100 ragged members on 148 SMs
    big, small = two_stream_split(m, m, 148)
    assert len(big) == 48 and len(small) == 52 and m[big].min() > m[small].max()
    assert sorted(list(big) + list(small)) == list(range(100))
    assert two_stream_split(np.full(100, 1024), np.full(100, 1024), 148) is None      # equal sizes: nothing to gain
    assert two_stream_split(m[:60], m[:60], 148) is None                             # <= half a wave: plain cluster launch
    assert two_stream_split(np.arange(100, 246), np.arange(100, 246), 148) is None   # 146 members: only 2 spare SMs
    assert two_stream_split(np.arange(100, 300), np.arange(100, 300), 148) is None   # more than one wave


def test_group_labels_equals_unique_plus_stable_argsort():
    from llckbdm_b200.ensemble import group_labels
    rng = np.random.default_rng(0)
    cases = [rng.integers(-1, 30, 1000), rng.integers(0, 5, 100), np.array([5, 5, 9, -1, 9, 100000]), np.arange(10) // 3 * 7 - 1,
             np.zeros(5, dtype=int), np.array([-1, -1, 2, 2, 0, 1])]
    for lab in cases:
        order, cluster_of, seg, values = group_labels(lab)
        uniq, inv = np.unique(lab, return_inverse=True)
        want = np.argsort(inv, kind="stable")
        assert np.array_equal(order, want) and np.array_equal(cluster_of, inv[want]) and np.array_equal(values, uniq)
        assert seg[0] == 0 and np.array_equal(seg[1:], np.cumsum(np.bincount(inv)))


def test_flatten_signals():
    a = np.arange(4) + 0j
    flat, off, lens = flatten_signals(a, 3)
    assert flat.shape == (4,) and list(off) == [0, 0, 0] and list(lens) == [4, 4, 4]
    flat, off, lens = flatten_signals([a, a[:2], a], 3)
    assert flat.shape == (10,) and list(off) == [0, 4, 6] and list(lens) == [4, 2, 4]
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


def _sklearn_mst(X, k_self):
    """sklearn's own exact spanning tree of the mutual-reachability graph (core distance = k_self-th neighbour incl. the point)."""
    from sklearn.cluster._hdbscan._linkage import mst_from_data_matrix
    from sklearn.metrics import DistanceMetric
    from sklearn.neighbors import NearestNeighbors
    cd = np.ascontiguousarray(NearestNeighbors(n_neighbors=k_self, algorithm="kd_tree").fit(X).kneighbors(X, k_self)[0][:, -1])
    return mst_from_data_matrix(np.asarray(X, order="C"), cd, DistanceMetric.get_metric("euclidean"), 1.0)


def test_native_labels_from_mst_reproduce_hdbscan_fit():
    """Host half of the device-accelerated HDBSCAN fits (llck_hdbscan_labels: dendrogram, condensed tree, EOM, labelling on native
    threads): feeding the clusterer's OWN spanning tree (sklearn's exact Prim) through it must reproduce
    sklearn.cluster.HDBSCAN(min_samples=k+1).fit(X).labels_ label for label -- k+1 because the reference's hdbscan package does
    not count the point itself (the GPU test checks that the device spanning trees equal sklearn's edge for edge).
    Inputs include tight clusters, duplicate points (zero distances -> infinite lambdas) and exact distance ties."""
    from sklearn.cluster import HDBSCAN
    from llckbdm_b200 import llckbdm as L
    rng = np.random.default_rng(1)
    cent = rng.uniform(-1, 1, (8, 3))
    X = np.concatenate([np.repeat(cent, 10, axis=0) + 1e-5 * rng.standard_normal((80, 3)), rng.uniform(-1, 1, (300, 3))])
    X = np.column_stack([X, np.zeros(len(X))])
    X2 = X.copy()
    X2[5] = X2[3]; X2[6] = X2[3]; X2[90:96] = X2[90]            # duplicates
    X2[100:140, :3] = np.round(X2[100:140, :3], 1)              # coarse grid: exact ties
    X3 = np.column_stack([rng.integers(0, 6, (500, 3)).astype(float), np.zeros(500)])   # integer lattice: massive ties + duplicates
    for Xc in (X, X2, X3):
        ks = (1, 2, 4, 11, 30)
        msts = [_sklearn_mst(Xc, k + 1) for k in ks]
        src = np.stack([t["current_node"] for t in msts]); dst = np.stack([t["next_node"] for t in msts])
        w = np.stack([t["distance"] for t in msts])
        got = L._labels_from_msts(src, dst, w)
        for k, g in zip(ks, got):
            want = HDBSCAN(min_samples=k + 1, copy=True).fit(Xc).labels_
            assert np.array_equal(g, want), k
            assert np.array_equal(g, L._fit_one(Xc, k)), k       # the library path uses the same convention
    # one fit through the single-tree helper; clipping of k to n-1
    t = _sklearn_mst(X, 5)
    assert np.array_equal(L._labels_from_mst(t["current_node"], t["next_node"], t["distance"]), HDBSCAN(min_samples=5, copy=True).fit(X).labels_)
    tiny = X[:7]
    assert np.array_equal(L._fit_one(tiny, 50), HDBSCAN(min_samples=7, copy=True).fit(tiny).labels_)
    assert L._library_min_samples(50, 7) == 7 and L._library_min_samples(1, 100) == 2


def test_hdbscan_labels_entry_validates_its_arguments(native_lib):
    """llck_hdbscan_labels is host code: callable without a GPU.  Out-of-range nodes / permutations are bad arguments, a 2-point
    tree gives all noise (min_cluster_size 5), and the thread count does not change the result."""
    import ctypes
    src = np.array([[0, 1, 2, 3, 4, 5, 6, 7, 8]], dtype=np.int64)
    dst = src + 1
    w = np.array([[1, 1, 1, 1, 9, 1, 1, 1, 1]], dtype=np.float64)
    order = np.argsort(w, axis=1).astype(np.int64)

    def run(src, dst, w, order, n, nthreads=0, mcs=5):
        labels = np.full((1, n), 77, dtype=np.int32)
        rc = native_lib.llck_hdbscan_labels(src.ctypes.data, dst.ctypes.data, w.ctypes.data, order.ctypes.data if order is not None else None,
                                            n, 1, mcs, nthreads, labels.ctypes.data)
        return rc, labels[0]

    rc, lab = run(src, dst, w, order, 10)
    assert rc == 0 and set(lab[:5]) == {lab[0]} and set(lab[5:]) == {lab[5]} and lab[0] != lab[5] and lab.min() == 0   # two chains of 5
    rc1, lab1 = run(src, dst, w, order, 10, nthreads=1)
    assert rc1 == 0 and np.array_equal(lab, lab1)
    bad = src.copy(); bad[0, 3] = 10
    assert run(bad, dst, w, order, 10)[0] == _native.E_BADARG
    badorder = order.copy(); badorder[0, 0] = 9
    assert run(src, dst, w, badorder, 10)[0] == _native.E_BADARG
    assert run(src, dst, w, order, 10, mcs=1)[0] == _native.E_BADARG
    rc, lab = run(src[:, :1], dst[:, :1], w[:, :1], None, 2)
    assert rc == 0 and list(lab) == [-1, -1]
