"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the public API and the C ABI,
against the CPU oracle on the same seeded inputs, against the golden vectors of the real reference, the
reference's own acceptance tests, and -- at full size -- size-independent properties.

Tolerances: per member, rows matched by nearest pole; |dmu|/|mu| <= 1e-8 for ALL rows on
noisy inputs; |dD|/|D| <= 1e-8 for rows with |D| > 1e-3 max|D|; singular values rel 1e-8 (abs 1e-12)."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DWELL = 5e-4
TOL = 1e-8


def _D(ll):
    return ll[:, 0] * np.exp(1j * ll[:, 3])


def _compare_ll(ll_gpu, ll_ref):
    from oracle.kbdm_oracle import compare_members, mu_from_line_list
    return compare_members(mu_from_line_list(ll_gpu, DWELL), _D(ll_gpu), mu_from_line_list(ll_ref, DWELL), _D(ll_ref))


@pytest.fixture(scope="module")
def cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from llckbdm_b200 import _native
    _native.load()        # fails loudly if the in-tree extension is missing
    return torch


@pytest.mark.parametrize("name", ["noisy_m16", "noisy_m64", "noisy_m128", "noisy_m200_l30", "noisy_m96_p2", "noisy_m80_q",
                                  "noisy_m100_l40_p2_q", "noisy_m256"])
def test_kbdm_matches_reference_golden(cuda, golden_dir, name):
    from llckbdm_b200.kbdm import kbdm
    g = np.load(os.path.join(golden_dir, f"kbdm_{name}.npz"))
    l = None if int(g["l"]) < 0 else int(g["l"])
    ll, info = kbdm(g["data"], float(g["dwell"]), m=int(g["m"]), p=int(g["p"]), l=l, q=float(g["q"]))
    assert ll.shape == g["line_list"].shape and ll.dtype == np.float64
    dmu, dD = _compare_ll(ll, g["line_list"])
    assert dmu < TOL and dD < TOL, (dmu, dD)
    assert info.singular_values.shape == (int(g["m"]),)
    assert np.allclose(info.singular_values, g["singular_values"], rtol=1e-8, atol=1e-12)
    assert info.m == int(g["m"]) and info.p == int(g["p"]) and info.q == float(g["q"])


def test_known_answer_16_components_reference_test_kbdm_svd(cuda):
    """Same assertions as reference llckbdm/_tests/test_kbdm.py:8-42, GPU backend."""
    from llckbdm_b200.kbdm import kbdm
    from oracle.kbdm_oracle import BRAIN_SIM_PARAMS, brain_sim
    ll, info = kbdm(brain_sim(2048, 0.0, 0), DWELL, m=300)
    assert ll.shape == (300, 4) and info.m == 300 and info.l == 300 and info.p == 1 and info.q == 0
    est = ll[ll[:, 0] > 1e-4]
    est = est[np.argsort(est[:, 2])]
    assert len(est) == 16
    assert np.allclose(est[:, 0], BRAIN_SIM_PARAMS[:, 0], rtol=1e-6)
    assert np.allclose(est[:, 1], BRAIN_SIM_PARAMS[:, 1], rtol=1e-3)
    assert np.allclose(est[:, 2], BRAIN_SIM_PARAMS[:, 2], atol=0.3)
    assert np.allclose(est[:, 3], 0.0, atol=1e-10)


def test_m_l_defaults_and_validation_on_gpu_api(cuda, caplog):
    """Reference _tests/test_kbdm.py:62-122."""
    from llckbdm_b200.kbdm import kbdm
    from oracle.kbdm_oracle import brain_sim
    c = brain_sim(2048, 0.0, 0)
    with pytest.raises(ValueError, match="l or m must be specified"):
        kbdm(data=c, dwell=DWELL)
    with pytest.raises(ValueError, match="l can't be greater than m"):
        kbdm(data=c, dwell=DWELL, l=30, m=20)
    with pytest.raises(ValueError, match=r"m or l can't be greater than \(n \+ 1 - p\)/2."):
        kbdm(data=c, dwell=DWELL, m=1025)
    ll, info = kbdm(data=c, dwell=DWELL, l=30)
    assert ll.shape == (30, 4) and info.l == 30 and info.m == 30
    caplog.set_level('DEBUG')
    ll, info = kbdm(c, DWELL, m=10, q=1e-3)
    assert 'Using Tikhonov Regularization' in caplog.text
    assert ll.shape == (10, 4) and info.q == pytest.approx(1e-3)


def test_sample_kbdm_batched_equals_oracle_members(cuda):
    """One batched launch over ragged sizes (incl. m not a multiple of 32, l < m) == per-member oracle."""
    from llckbdm_b200.sampling import sample_kbdm
    from oracle.kbdm_oracle import brain_sim, kbdm_oracle
    c = brain_sim(2048, 1e-3, 5)
    m_range = [33, 100, 64, 257, 150, 7, 96]
    lls, infos = sample_kbdm(c, DWELL, m_range, p=1, l=None, q=0, filter_invalid_features=False)
    assert len(lls) == len(m_range) == len(infos)
    for m, ll, info in zip(m_range, lls, infos):
        ll_o, info_o = kbdm_oracle(c, DWELL, m=m)
        assert ll.shape == (m, 4) and info.m == m
        dmu, dD = _compare_ll(ll, ll_o)
        assert dmu < TOL and dD < TOL, (m, dmu, dD)
        assert np.allclose(info.singular_values, info_o.singular_values, rtol=1e-8, atol=1e-12)


def test_sample_kbdm_reference_contract(cuda):
    """Reference _tests/test_sampling.py:19-62: equality with direct kbdm, filter leaves the 16 true lines."""
    from llckbdm_b200.kbdm import kbdm
    from llckbdm_b200.sampling import filter_samples, sample_kbdm
    from oracle.kbdm_oracle import BRAIN_SIM_PARAMS, brain_sim
    c = brain_sim(2048, 0.0, 0)
    lls, infos = sample_kbdm(c, DWELL, range(100, 103), p=1, l=None, q=0, filter_invalid_features=False)
    assert len(lls) == 3
    ll0, info0 = kbdm(c, DWELL, m=100, p=1, l=None)
    big = ll0[:, 0] > 1e-4
    assert big.sum() == 16
    a = lls[0][lls[0][:, 0] > 1e-4]
    assert np.allclose(np.sort(a[:, 2]), np.sort(ll0[big, 2]), rtol=1e-9)
    assert infos[0].m == info0.m and infos[0].l == info0.l
    assert np.allclose(infos[0].singular_values[:16], info0.singular_values[:16], rtol=1e-9)
    f = filter_samples(kbdm(c, DWELL, m=150)[0])
    assert len(f) == 16
    f = f[np.argsort(f[:, 0])]
    truth = BRAIN_SIM_PARAMS[np.argsort(BRAIN_SIM_PARAMS[:, 0])]
    assert np.allclose(f[0], truth[0], atol=0.01) and np.allclose(f[-1], truth[-1], atol=0.01)


def test_min_rmse_kbdm_reference_test(cuda):
    """Reference _tests/test_min_rmse_kbdm.py:6-23: min_index == 2, rmse ~ 0."""
    from llckbdm_b200.min_rmse_kbdm import min_rmse_kbdm
    from oracle.kbdm_oracle import brain_sim
    r = min_rmse_kbdm(data=brain_sim(2048, 0.0, 0), dwell=DWELL, m_range=[30, 31, 180, 32, 33, 34], l=30)
    assert len(r.samples) == 6
    assert r.min_rmse == pytest.approx(0, abs=1e-6)
    assert r.min_index == 2


def test_llc_kbdm_reference_test(cuda):
    """Reference _tests/test_llckbdm.py:37-57: 16 lines after clustering, residual std < 1e-3."""
    from llckbdm_b200.llckbdm import llc_kbdm
    from llckbdm_b200.sig_gen import multi_fid
    from oracle.kbdm_oracle import brain_sim
    c = brain_sim(2048, 0.0, 0)
    res = llc_kbdm(c, DWELL, m_range=range(250, 260), l=30)
    ll = res.line_list[res.line_list[:, 0] > 1e-3]
    assert len(ll) == 16
    t = np.linspace(0, DWELL * 2048, 2048, endpoint=False)
    resid = c - multi_fid(t, ll)
    assert np.std(resid.real) < 1e-3 and np.std(resid.imag) < 1e-3
    with pytest.raises(ValueError, match="size of 'm_range' must be greater than 2."):
        llc_kbdm(c, DWELL, m_range=[100])


def test_llc_kbdm_cluster_parity_noisy(cuda):
    """Clustered estimates from GPU line lists == clustered estimates from oracle line lists (same CPU clusterer),
    rel 1e-6, identical cluster count."""
    from llckbdm_b200 import llckbdm as L
    from llckbdm_b200.sampling import filter_samples, sample_kbdm
    from oracle.kbdm_oracle import brain_sim, sample_kbdm_oracle
    c = brain_sim(2048, 1e-3, 11)
    m_range = list(range(120, 128))
    gpu, _ = sample_kbdm(c, DWELL, m_range, p=1, l=None)
    ref, _ = sample_kbdm_oracle(c, DWELL, m_range, p=1, l=None)
    assert [len(a) for a in gpu] == [len(a) for a in ref]
    out = []
    for lls in (gpu, ref):
        s = np.concatenate([a[np.lexsort((a[:, 0], a[:, 2]))] for a in lls])     # canonical row order
        s = filter_samples(s)
        r = L._cluster_line_lists(s, L._transform_line_lists(s, DWELL), min_samples=4)
        out.append((r.num_clusters, np.asarray(r.labels), np.asarray(r.summarized_line_list)))
    assert out[0][0] == out[1][0] and out[0][0] > 0
    assert np.array_equal(out[0][1], out[1][1])
    a, b = out[0][2], out[1][2]
    assert np.allclose(a[:, [0, 2]], b[:, [0, 2]], rtol=1e-6, atol=1e-9)
    assert np.allclose(a[:, 1], b[:, 1], rtol=1e-6)


def test_silhouette_kernel_matches_sklearn(cuda):
    """llck_silhouette_batched vs sklearn.metrics.silhouette_samples (the call at reference llckbdm.py:291): several labelings of
    the same points in one launch, noise label -1 as a cluster, singleton clusters, many tiny clusters, one big cluster.
    sklearn's Euclidean distances use the |x|^2+|y|^2-2xy expansion (absolute error ~1e-8 on near-coincident points), the
    kernel takes differences directly, hence the 1e-6 absolute tolerance on coefficients in [-1, 1]."""
    from sklearn.metrics import silhouette_samples
    from llckbdm_b200.ensemble import silhouette_samples_device
    rng = np.random.default_rng(2)
    cent = rng.uniform(-1, 1, (40, 3))
    X = np.concatenate([np.repeat(cent, 12, axis=0) + 1e-3 * rng.standard_normal((480, 3)), rng.uniform(-1, 1, (777, 3))])
    X = np.column_stack([X, np.zeros(len(X))])
    n = len(X)
    lab_a = np.concatenate([np.repeat(np.arange(40), 12), -np.ones(777, dtype=int)])
    lab_b = rng.integers(-1, 5, n)
    lab_c = np.arange(n) // 3                      # tiny clusters
    lab_c[-1] = 10 ** 6                            # a singleton
    lab_d = (X[:, 0] > 0).astype(int)
    labelings = [lab_a, lab_b, lab_c, lab_d]
    got = silhouette_samples_device(X, labelings)
    for lab, g in zip(labelings, got):
        want = silhouette_samples(X, lab)
        assert np.abs(g - want).max() < 1e-6, np.abs(g - want).max()
    assert got[2][-1] == 0.0


def test_gpu_hdbscan_spanning_trees_give_identical_labels(cuda):
    """Device core distances + Prim spanning trees (llck_hdbscan_core_distances / llck_hdbscan_mst) followed by the clusterer's own tree
    condensation == sklearn.cluster.HDBSCAN(min_samples=k).fit(X).labels_, label for label, for every k; the edge lists equal
    sklearn's mst_from_data_matrix bit for bit (duplicates and exact ties included)."""
    from sklearn.cluster._hdbscan._linkage import mst_from_data_matrix
    from sklearn.metrics import DistanceMetric
    from sklearn.neighbors import NearestNeighbors
    from llckbdm_b200 import llckbdm as L
    from llckbdm_b200.ensemble import hdbscan_msts_device
    rng = np.random.default_rng(4)
    cent = rng.uniform(-1, 1, (20, 3))
    X = np.concatenate([np.repeat(cent, 15, axis=0) + 1e-6 * rng.standard_normal((300, 3)), rng.uniform(-1, 1, (2100, 3))])
    X = np.column_stack([X, np.zeros(len(X))])
    X[7] = X[3]                                    # duplicate points: zero distances
    X[100:110, :3] = np.round(X[100:110, :3], 1)   # coarse grid: exact distance ties
    from sklearn.cluster import HDBSCAN
    ks = [1, 2, 5, 9, 33]                          # neighbour counts INCLUDING the point itself (sklearn's min_samples)
    src, dst, w = hdbscan_msts_device(X, ks)                      # thread-block cluster per fit, points resident on chip
    src1, dst1, w1 = hdbscan_msts_device(X, ks, single_cta=True)  # one CTA per fit streaming from L2 (large point sets)
    assert np.array_equal(src, src1) and np.array_equal(dst, dst1) and np.array_equal(w, w1)
    for f, k in enumerate(ks):
        cd = np.ascontiguousarray(NearestNeighbors(n_neighbors=k, algorithm="kd_tree").fit(X).kneighbors(X, k)[0][:, -1])
        ref = mst_from_data_matrix(np.asarray(X, order="C"), cd, DistanceMetric.get_metric("euclidean"), 1.0)
        assert np.array_equal(src[f], ref["current_node"]) and np.array_equal(dst[f], ref["next_node"])
        assert np.array_equal(w[f], ref["distance"])
        assert np.array_equal(L._labels_from_mst(src[f], dst[f], w[f]), HDBSCAN(min_samples=k, copy=True).fit(X).labels_)
    # the reference's hdbscan.HDBSCAN(min_samples=k) does not count the point itself: k+1 in sklearn's convention
    assert L._gpu_fit_supported(X, list(range(1, 12)))
    got = L._fit_all(X, list(range(1, 12)))
    for k, lab in zip(range(1, 12), got):
        assert np.array_equal(lab, HDBSCAN(min_samples=k + 1, copy=True).fit(X).labels_), k
        assert np.array_equal(lab, L._fit_one(X, k)), k
    L.CLUSTER_BACKEND = "host"                     # explicit switch (no environment variable): plain library fits
    try:
        assert not L._gpu_fit_supported(X, [1, 2])
        assert np.array_equal(L._fit_all(X, [3])[0], got[2])
    finally:
        L.CLUSTER_BACKEND = "device"


def test_llc_kbdm_matches_host_clustering_stage(cuda):
    """llc_kbdm end to end (GPU solves, parallel fits, device silhouettes, device RMSE selection) == the same line lists pushed
    through a literal host restatement of reference llckbdm.py:93-141 (same clusterer, sklearn silhouettes, oracle RMSE)."""
    from sklearn.metrics import silhouette_samples
    from llckbdm_b200 import llckbdm as L
    from llckbdm_b200.sampling import filter_samples, sample_kbdm
    from oracle.kbdm_oracle import brain_sim, min_rmse_oracle
    c = brain_sim(1024, 1e-3, 5)
    m_range = list(range(100, 112))
    res = L.llc_kbdm(c, DWELL, m_range)
    lls, _ = sample_kbdm(c, DWELL, m_range, p=1, l=None)
    samples = filter_samples(np.concatenate(lls))
    feats = L._transform_line_lists(samples, DWELL)
    cands, sils = [], []
    for ms in range(1, len(m_range)):
        labels = L._fit_one(feats, ms)
        nc = len(set(labels.tolist()) - {-1})
        if nc == 0:
            continue
        sv = silhouette_samples(feats, labels)
        clusters = [np.nonzero(labels == k) for k in range(nc)]
        cands.append(L._summarize_clusters(samples, clusters))
        sils.append(np.array([np.average(sv[cl]) for cl in clusters]))
    k, rmses = min_rmse_oracle(c, DWELL, cands)
    assert np.allclose(res.line_list, cands[k], rtol=1e-12, atol=0)
    assert abs(res.rmse - rmses[k]) < 1e-9 * rmses[k]
    assert np.abs(res.silhouette - sils[k]).max() < 1e-6


def test_pooled_samples_and_features_match_host_sequence(cuda):
    """llck_pool_features (device concatenate + filter_samples + _transform_line_lists) == the host sequence of reference
    llckbdm.py:94-98 on the same solve: kept rows bit-identical and in the same order, features to 1e-14."""
    from llckbdm_b200 import llckbdm as L
    from llckbdm_b200.sampling import filter_samples, sample_kbdm, sample_kbdm_pooled
    from oracle.kbdm_oracle import brain_sim
    c = brain_sim(1024, 1e-3, 9)
    m_range = [40, 300, 64, 129, 17]                      # ragged, unsorted: pooled order must follow m_range
    samples, feats = sample_kbdm_pooled(c, DWELL, m_range, p=1, l=None)
    lls, _ = sample_kbdm(c, DWELL, m_range, p=1, l=None)
    want = filter_samples(np.concatenate(lls))
    assert samples.shape == want.shape and np.array_equal(samples, want)
    wf = L._transform_line_lists(want, DWELL)
    assert np.abs(feats - wf).max() < 1e-14
    assert np.all(feats[:, 3] == 0.0)


def test_multi_fid_batched_device_matches_sig_gen(cuda):
    """llck_multi_fid_batched vs the host multi_fid (reference sig_gen.py:57-71) on ragged parameter sets."""
    from llckbdm_b200 import sig_gen
    from oracle.kbdm_oracle import BRAIN_SIM_PARAMS
    rng = np.random.default_rng(6)
    sets = [BRAIN_SIM_PARAMS, BRAIN_SIM_PARAMS[:3] * [1.5, 0.9, 1.0, 1.0] + [0, 0, 2.0, 0.3], np.array([[0.7, np.inf, -333.0, -1.0]])]
    N = 1000
    got = sig_gen.multi_fid_batched_device(sets, N, DWELL).cpu().numpy()
    t, _ = sig_gen.gen_t_freq_arrays(N, DWELL)
    for ps, g in zip(sets, got):
        want = sig_gen.multi_fid(t, ps)
        assert np.abs(g - want).max() < 1e-13 * np.abs(want).max()
    with pytest.raises(ValueError, match="T2 must be positive"):
        sig_gen.multi_fid_batched_device([np.array([[1.0, 0.0, 1.0, 0.0]])], 16, DWELL)


def test_iterative_llc_kbdm_runs_and_reduces_residual(cuda, capsys):
    """Residual-iteration driver (reference llckbdm.py:144-199): two iterations on a small noisy FID; the fitted model must explain
    the signal (frequency-domain RMSE of the pooled estimate far below the signal level)."""
    from llckbdm_b200.llckbdm import iterative_llc_kbdm
    from llckbdm_b200.metrics import calculate_freq_domain_rmse
    from oracle.kbdm_oracle import brain_sim
    c = brain_sim(1024, 1e-3, 2)
    res = iterative_llc_kbdm(c, DWELL, m_range=range(100, 108), max_iterations=2)
    assert "Iteration #0" in capsys.readouterr().out
    assert len(res.line_list) >= 10 and len(res.line_lists) >= 1 and np.isfinite(res.rmse)
    assert calculate_freq_domain_rmse(c, res.line_list, DWELL) < 0.05 * np.sqrt(np.mean(np.abs(np.fft.fft(c) / np.sqrt(len(c))) ** 2))


def test_singular_member_raises_linalgerror(cuda):
    """Exact zero singular value among the kept ones -> LinAlgError (np.linalg.inv behaviour at kbdm.py:186)."""
    from llckbdm_b200.kbdm import kbdm
    with pytest.raises(np.linalg.LinAlgError):
        kbdm(np.zeros(64, dtype=complex), DWELL, m=8)


def test_real_input_is_accepted(cuda):
    from llckbdm_b200.kbdm import kbdm
    from oracle.kbdm_oracle import brain_sim, kbdm_oracle
    c = brain_sim(256, 1e-3, 2).real.copy()
    ll, info = kbdm(c, DWELL, m=40)
    ll_o, _ = kbdm_oracle(c.astype(complex), DWELL, m=40)
    dmu, dD = _compare_ll(ll, ll_o)
    assert dmu < TOL and dD < TOL


def test_zgemm_stage_entry_hankel_and_conjt(cuda):
    """The production DMMA GEMM kernel through llck_zgemm: implicit Hankel operand, conj-transpose and plain."""
    torch = cuda
    from llckbdm_b200 import _native
    lib = _native.load()
    rng = np.random.default_rng(0)
    M, N, K = 150, 70, 131
    sig = rng.standard_normal(M + K + 5) + 1j * rng.standard_normal(M + K + 5)
    B = rng.standard_normal((K, N)) + 1j * rng.standard_normal((K, N))
    A = rng.standard_normal((M, K)) + 1j * rng.standard_normal((M, K))
    dev = torch.device("cuda:0")

    def to_dev(x):   # column-major device copy
        return torch.from_numpy(np.asfortranarray(x).T.copy().view(np.float64)).to(dev)

    def run(amode, Ad, lda, shift=0):
        Cd = torch.zeros((N, M, 2), dtype=torch.float64, device=dev)
        rc = lib.llck_zgemm(amode, Ad.data_ptr() if Ad is not None else None, lda, Bd.data_ptr(), K, Cd.data_ptr(), M,
                            M, N, K, sigd.data_ptr(), shift, None)
        assert rc == 0
        return Cd.cpu().numpy().view(np.complex128)[..., 0].T

    Bd, sigd = to_dev(B), torch.from_numpy(sig.view(np.float64)).to(dev)
    H = sig[np.arange(M)[:, None] + np.arange(K)[None, :] + 3]
    assert np.abs(run(2, None, 0, shift=3) - H @ B).max() < 1e-11
    assert np.abs(run(0, to_dev(A), M) - A @ B).max() < 1e-11
    At = rng.standard_normal((K, M)) + 1j * rng.standard_normal((K, M))
    assert np.abs(run(1, to_dev(At), K) - At.conj().T @ B).max() < 1e-11


def test_full_size_m1024_properties(cuda):
    """BASELINE size (N=2048, m=l=1024): size-independent properties (the pole-by-pole oracle comparison at this size is
    tests/test_gpu_configs.py::test_headline_m1024_all_poles_against_oracle): singular values vs LAPACK, generalized-eigen
    residual of sampled poles, amplitude identity, signal reconstruction."""
    from llckbdm_b200.ensemble import solve_ensemble
    from oracle.kbdm_oracle import brain_sim, hankel_matrices
    c = brain_sim(2048, 1e-3, 0)
    m = 1024
    res = solve_ensemble(c, [m], [m], 1, 0.0, DWELL)
    assert res.status[0] == 0
    U0, _, U1 = hankel_matrices(c, m, 1)
    s_ref = np.linalg.svd(U0, compute_uv=False)
    assert np.allclose(res.sing_vals[0], s_ref, rtol=1e-8, atol=1e-12)
    mu, D = res.mu[0], res.D[0]
    # every pole is a generalized eigenvalue of (U1, U0): sigma_min(U1 - mu U0) ~ 0, checked on a sample of poles
    smax = s_ref[0]
    for k in np.linspace(0, m - 1, 6).astype(int):
        smin = np.linalg.svd(U1 - mu[k] * U0, compute_uv=False)[-1]
        assert smin < 1e-9 * smax
    # sum_k D_k mu_k^n reproduces the signal (harmonic inversion is exact for l = m); checked on the first 64 points
    n = np.arange(64)
    recon = (D[None, :] * mu[None, :] ** n[:, None]).sum(axis=1)
    assert np.abs(recon - c[:64]).max() < 1e-7 * np.abs(c).max()
    ll = res.line_lists[0]
    assert np.allclose(ll[:, 0], np.abs(D)) and np.allclose(ll[:, 3], np.angle(D))


def test_bidiag_stage_entry(cuda):
    """Blocked Householder bidiagonalisation through llck_bidiag_test: A = Q B P^H with B REAL upper bidiagonal."""
    torch = cuda
    from llckbdm_b200 import _native
    from oracle.kbdm_oracle import brain_sim, hankel_matrices
    lib = _native.load()
    dev = torch.device("cuda:0")
    for m in (1, 2, 31, 33, 100, 257):
        ld = lib.llck_leading_dim(m)
        if m <= 33:
            rng = np.random.default_rng(m)
            A = rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m))
        else:
            A, _, _ = hankel_matrices(brain_sim(2 * m + 8, 1e-3, 0), m, 1)
        Ap = np.zeros((ld, ld), dtype=complex)
        Ap[:m, :m] = A
        Ad = torch.from_numpy(np.ascontiguousarray(Ap.T).view(np.float64)).to(dev)
        Qd = torch.zeros((ld, ld, 2), dtype=torch.float64, device=dev)
        Pd = torch.zeros_like(Qd)
        dd = torch.zeros(ld, dtype=torch.float64, device=dev)
        ed = torch.zeros(ld, dtype=torch.float64, device=dev)
        assert lib.llck_bidiag_test(Ad.data_ptr(), m, ld, dd.data_ptr(), ed.data_ptr(), Qd.data_ptr(), Pd.data_ptr(), None) == 0
        Q = Qd.cpu().numpy().view(np.complex128)[..., 0].T[:m, :m]
        P = Pd.cpu().numpy().view(np.complex128)[..., 0].T[:m, :m]
        B = np.diag(dd.cpu().numpy()[:m]) + np.diag(ed.cpu().numpy()[:m - 1], 1)
        scale = np.abs(A).max()
        assert np.abs(Q @ B @ P.conj().T - A).max() < 1e-12 * scale
        assert np.abs(Q.conj().T @ Q - np.eye(m)).max() < 1e-12 and np.abs(P.conj().T @ P - np.eye(m)).max() < 1e-12
        s_ref = np.linalg.svd(A, compute_uv=False)
        assert np.allclose(np.linalg.svd(B, compute_uv=False), s_ref, rtol=1e-9, atol=1e-12 * s_ref[0])


@pytest.mark.parametrize("svd_mode", ["dc", "jacobi"])
def test_both_svd_paths_tiny_and_ragged(cuda, svd_mode):
    """Tiny and ragged members (m = 1..65, l < m, p > 1, q > 0) through both SVD back ends of the bidiagonal, selected by the
    explicit llck_options struct: divide and conquer (default) and the real block Jacobi (what rank-deficient members fall back to)."""
    from llckbdm_b200 import _native
    from llckbdm_b200.ensemble import solve_ensemble
    from oracle.kbdm_oracle import brain_sim, compare_members, kbdm_oracle
    opts = _native.Options(svd_mode=_native.SVD_JACOBI if svd_mode == "jacobi" else _native.SVD_DC)
    c = brain_sim(2048, 1e-3, 7)
    ms = [1, 2, 3, 5, 31, 32, 34, 65]
    res = solve_ensemble(c, ms, ms, 1, 0.0, DWELL, options=opts)
    assert (res.status == 0).all()
    for k, m in enumerate(ms):
        _, info, mu, D = kbdm_oracle(c, DWELL, m=m, return_mu=True)
        dmu, dD = compare_members(res.mu[k, :m], res.D[k, :m], mu, D)
        assert dmu < TOL and dD < TOL, (m, dmu, dD)
        assert np.allclose(res.sing_vals[k, :m], info.singular_values, rtol=1e-8, atol=1e-12)
    for (m, l, p, q) in [(64, 1, 1, 0.0), (40, 2, 3, 0.0), (100, 7, 1, 1e-2), (33, 33, 4, 0.0)]:
        r = solve_ensemble(c, [m], [l], p, q, DWELL, options=opts)
        assert r.status[0] == 0
        _, _, mu, D = kbdm_oracle(c, DWELL, m=m, l=l, p=p, q=q, return_mu=True)
        dmu, dD = compare_members(r.mu[0, :l], r.D[0, :l], mu, D)
        assert dmu < TOL and dD < TOL, (m, l, p, q, dmu, dD)


def test_bdc_stage_entry(cuda):
    """Divide-and-conquer SVD of real bidiagonals through llck_bdc_test: ragged batch (one leaf ... five merge levels),
    gaussian / graded / clustered / split / diagonal inputs, against numpy.linalg.svd."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from bdc_check import run_bdc
    rng = np.random.default_rng(3)
    ds, es = [], []
    for m in (1, 2, 17, 32, 33, 64, 65, 130, 257):
        ds.append(rng.standard_normal(m) + 3.0); es.append(rng.standard_normal(m - 1))      # well conditioned
    for m in (96, 300, 512):
        g = np.logspace(0, -4, m)
        ds.append(g * (2.0 + rng.random(m))); es.append(0.3 * g[:-1] * rng.standard_normal(m - 1))   # graded
        ds.append(np.ones(m)); es.append(np.full(m - 1, 1e-3))                                       # clustered
        d = rng.standard_normal(m) + 3.0; e = rng.standard_normal(m - 1); e[m // 3] = 0.0
        ds.append(d); es.append(e)                                                                   # exact split
        ds.append(np.arange(1, m + 1, dtype=float)); es.append(np.zeros(m - 1))                      # diagonal
    solved = 0
    for (s, Us, V, fb), d, e in zip(run_bdc(ds, es), ds, es):
        m = len(d)
        B = np.diag(d) + np.diag(e, 1)
        s_ref = np.linalg.svd(B, compute_uv=False)
        if fb:      # only numerically rank-deficient members may be handed to the Jacobi path
            assert s_ref[-1] < 1e-7 * s_ref[0]
            continue
        solved += 1
        cond = s_ref[0] / s_ref[-1]
        assert np.abs(s - s_ref).max() < 1e-13 * s_ref[0]
        U = Us / s
        assert np.abs(U.T @ U - np.eye(m)).max() < 1e-14 * max(cond, 100.0)
        assert np.abs(V.T @ V - np.eye(m)).max() < 1e-14 * max(cond, 100.0)
        assert np.abs(B - Us @ V.T).max() < 1e-13 * s_ref[0]
    assert solved >= 18


def test_mixed_batch_rank_deficient_member_falls_back_to_jacobi(cuda):
    """One batch holding a noiseless (numerically rank-16) FID next to noisy ones: the divide-and-conquer SVD flags the
    rank-deficient member, the Jacobi path solves it, and the noisy members keep full parity."""
    from llckbdm_b200.ensemble import solve_ensemble
    from oracle.kbdm_oracle import BRAIN_SIM_PARAMS, brain_sim, compare_members, kbdm_oracle
    sigs = [brain_sim(512, 1e-3, 1), brain_sim(512, 0.0, 0), brain_sim(512, 1e-3, 2)]
    ms = [200, 256, 130]
    res = solve_ensemble(sigs, ms, ms, 1, 0.0, DWELL)
    assert (res.status == 0).all()
    for k in (0, 2):
        _, info, mu, D = kbdm_oracle(sigs[k], DWELL, m=ms[k], return_mu=True)
        dmu, dD = compare_members(res.mu[k, :ms[k]], res.D[k, :ms[k]], mu, D)
        assert dmu < TOL and dD < TOL, (k, dmu, dD)
        assert np.allclose(res.sing_vals[k, :ms[k]], info.singular_values, rtol=1e-8, atol=1e-12)
    ll = res.line_lists[1, :ms[1]]
    est = ll[(ll[:, 0] > 1e-4) & (ll[:, 1] > 0)]
    est = est[np.argsort(est[:, 2])]
    assert len(est) == 16
    assert np.allclose(est[:, 0], BRAIN_SIM_PARAMS[:, 0], rtol=1e-6)
    assert np.allclose(est[:, 2], BRAIN_SIM_PARAMS[:, 2], atol=1e-6)


def test_rmse_scoring_kernel_matches_oracle_and_reference_golden(cuda, golden_dir):
    """llck_rmse_batched (Parseval form, no FFT) vs the oracle's fft-based restatement of metrics.py:7-17 and the values the real
    reference produced (tests/golden/rmse_noisy.npz): ragged candidates, an empty one, on-the-fly row filter."""
    from llckbdm_b200.ensemble import score_candidates
    from oracle.kbdm_oracle import BRAIN_SIM_PARAMS, filter_samples_oracle, freq_domain_rmse_oracle
    fid = np.load(os.path.join(golden_dir, "brain_sim_fid.npz"))
    g = np.load(os.path.join(golden_dir, "sample_kbdm_noisy.npz"))
    r = np.load(os.path.join(golden_dir, "rmse_noisy.npz"))
    lls = [g[f"ll{i}"] for i in range(int(g["n"]))]
    got = score_candidates(fid["noisy"], DWELL, lls + [BRAIN_SIM_PARAMS, np.zeros((0, 4))])
    assert np.allclose(got[:len(lls)], r["rmses"], rtol=1e-9, atol=0)
    assert abs(got[len(lls)] - float(r["rmse_truth"])) < 1e-9 * float(r["rmse_truth"])
    assert got[-1] == np.inf
    # unfiltered solver output + filter on the fly == oracle on the filtered list; odd N and N not a multiple of 256
    rng = np.random.default_rng(5)
    for N in (2048, 1000, 777):
        data = fid["noisy"][:N]
        cand = np.column_stack([rng.random(40) - 0.2, rng.random(40) * 0.2 - 0.02, rng.uniform(-900, 900, 40), rng.uniform(-3, 3, 40)])
        cand[3, 1] = np.inf
        want = freq_domain_rmse_oracle(data, filter_samples_oracle(cand), DWELL)
        got = score_candidates(data, DWELL, [cand], filter_rows=True)[0]
        assert abs(got - want) < 1e-10 * want, (N, got, want)


def test_min_rmse_kbdm_fused_scoring_matches_oracle(cuda):
    """min_rmse_kbdm(samples=None): members solved and scored on the device in one pass == oracle solve + oracle scoring."""
    from llckbdm_b200.min_rmse_kbdm import min_rmse_kbdm
    from oracle.kbdm_oracle import brain_sim, min_rmse_oracle, sample_kbdm_oracle
    c = brain_sim(1024, 1e-3, 3)
    m_range = [60, 100, 128, 200]
    res = min_rmse_kbdm(c, DWELL, m_range=m_range, l=None)
    lls, _ = sample_kbdm_oracle(c, DWELL, m_range, 1, None)
    k, rmses = min_rmse_oracle(c, DWELL, lls)
    assert res.min_index == k and len(res.rmses_list) == len(rmses)
    assert np.allclose(res.rmses_list, rmses, rtol=1e-6)
    assert np.allclose(res.min_rmse, rmses[k], rtol=1e-6)
    # candidates given by the caller: same scores through the packed launch
    res2 = min_rmse_kbdm(c, DWELL, samples=res.samples)
    assert res2.min_index == k and np.allclose(res2.rmses_list, res.rmses_list, rtol=1e-12)
    with pytest.raises(ValueError, match="T2 must be positive"):
        min_rmse_kbdm(c, DWELL, samples=[np.array([[1.0, -1.0, 10.0, 0.0]])])


def test_noiseless_rank_deficient_input_true_components(cuda):
    """Noiseless brain_sim (numerical rank 16, cond ~1e18): kernels must not NaN/hang; the 16 true components match the
    oracle (spurious poles are not reproducible between any two implementations)."""
    from llckbdm_b200.kbdm import kbdm
    from oracle.kbdm_oracle import BRAIN_SIM_PARAMS, brain_sim
    ll, info = kbdm(brain_sim(2048, 0.0, 0), DWELL, m=512)
    assert np.isfinite(ll[:, [0, 2, 3]]).all()
    est = ll[(ll[:, 0] > 1e-4) & (ll[:, 1] > 0)]
    est = est[np.argsort(est[:, 2])]
    assert len(est) == 16
    assert np.allclose(est[:, 0], BRAIN_SIM_PARAMS[:, 0], rtol=1e-6)
    assert np.allclose(est[:, 2], BRAIN_SIM_PARAMS[:, 2], atol=1e-6)
