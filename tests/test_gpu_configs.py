"""GPU parity tests at the BASELINE.json configuration shapes (run with -m gpu on a B200): the headline size m = l = 1024
and the C2 / C3 / C4 / C5 shapes against the CPU oracle on the same inputs, wave-boundary batches, the LLC-KBDM clustering
end to end, the sharded multi-GPU path, and the C-ABI contract (asynchronous call, bounds errors, explicit options).

Tolerances: per member, rows matched by nearest pole; |dmu|/|mu| <= 1e-8 for ALL rows on noisy inputs;
|dD|/|D| <= 1e-8 for rows with |D| > 1e-3 max|D|; singular values rel 1e-8 (abs 1e-12)."""
import ctypes
import functools
import os
import socket
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DWELL = 5e-4
TOL = 1e-8


@pytest.fixture(scope="module")
def cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from llckbdm_b200 import _native
    _native.load()        # fails loudly if the in-tree extension is missing
    return torch


@functools.lru_cache(maxsize=None)
def _oracle_c2(m):
    """Oracle solve of the C1/C2 FID at Hankel size m (cached: m = 1024 costs a few seconds)."""
    from oracle.kbdm_oracle import brain_sim, kbdm_oracle
    return kbdm_oracle(brain_sim(2048, 1e-3, 0), DWELL, m=m, return_mu=True)


def _check_member(mu, D, sv, oracle, tag):
    from oracle.kbdm_oracle import compare_members
    _, info_o, mu_o, D_o = oracle
    dmu, dD = compare_members(mu, D, mu_o, D_o)
    assert dmu < TOL and dD < TOL, (tag, dmu, dD)
    assert np.allclose(sv, info_o.singular_values, rtol=1e-8, atol=1e-12), tag
    return dmu, dD


def test_headline_m1024_all_poles_against_oracle(cuda):
    """BASELINE headline (config C1): N = 2048, m = l = 1024 -- ALL 1024 poles, the well-conditioned amplitudes and all
    singular values against the oracle's scipy svd/eig solve of the same FID."""
    from llckbdm_b200.ensemble import solve_ensemble
    from llckbdm_b200.kbdm import kbdm
    from oracle.kbdm_oracle import brain_sim, mu_from_line_list
    c = brain_sim(2048, 1e-3, 0)
    res = solve_ensemble(c, [1024], [1024], 1, 0.0, DWELL)
    assert res.status[0] == 0
    _check_member(res.mu[0], res.D[0], res.sing_vals[0], _oracle_c2(1024), "m=1024 (one member, thread-block clusters)")
    # the public call returns the same solve
    ll, info = kbdm(c, DWELL, m=1024)
    assert ll.shape == (1024, 4) and info.singular_values.shape == (1024,)
    _check_member(mu_from_line_list(ll, DWELL), ll[:, 0] * np.exp(1j * ll[:, 3]), info.singular_values, _oracle_c2(1024), "kbdm(m=1024)")


def test_c2_slice_every_member_against_oracle(cuda):
    """A ragged slice of config C2 (m in [700, 1024]) in one batch: every member, including the largest, against the oracle."""
    from llckbdm_b200.ensemble import solve_ensemble
    from oracle.kbdm_oracle import brain_sim
    c = brain_sim(2048, 1e-3, 0)
    ms = [700, 703, 1024, 857]
    res = solve_ensemble(c, ms, ms, 1, 0.0, DWELL)
    assert (res.status == 0).all()
    for k, m in enumerate(ms):
        _check_member(res.mu[k, :m], res.D[k, :m], res.sing_vals[k, :m], _oracle_c2(m), f"C2 m={m}")


def test_c3_min_rmse_sweep_n4096_matches_oracle(cuda):
    """Config C3 shape: min_rmse_kbdm over m on a 4096-point FID with a pseudo-noise draw; members m = 256, 640, 1024:
    every member against the oracle, identical ``min_index``, RMSE values rel 1e-6."""
    from llckbdm_b200.ensemble import solve_ensemble
    from llckbdm_b200.min_rmse_kbdm import min_rmse_kbdm
    from oracle.kbdm_oracle import brain_sim, kbdm_oracle, min_rmse_oracle, sample_kbdm_oracle
    c = brain_sim(4096, 1e-3, 0)
    rng = np.random.default_rng(1)
    c3 = c + 1e-6 * (rng.standard_normal(4096) + 1j * rng.standard_normal(4096))
    m_range = [256, 640, 1024]
    res = solve_ensemble(c3, m_range, m_range, 1, 0.0, DWELL)
    assert (res.status == 0).all()
    for k, m in enumerate(m_range):
        _check_member(res.mu[k, :m], res.D[k, :m], res.sing_vals[k, :m], kbdm_oracle(c3, DWELL, m=m, return_mu=True), f"C3 m={m}")
    r = min_rmse_kbdm(c3, DWELL, m_range=m_range, l=None)
    lls, _ = sample_kbdm_oracle(c3, DWELL, m_range, 1, None)
    k, rmses = min_rmse_oracle(c3, DWELL, lls)
    assert r.min_index == k and len(r.samples) == len(lls)
    assert np.allclose(r.rmses_list, rmses, rtol=1e-6)


def test_c4_mrsi_voxels_m512_against_oracle(cuda):
    """Config C4 shape: independent 1024-point voxel FIDs (perturbed brain_sim parameters + noise), m = l = 512, one FID per
    member; 8 voxels of a 160-voxel batch (more than one wave of the one-CTA-per-member kernels) against the oracle."""
    from llckbdm_b200 import workloads
    from llckbdm_b200.ensemble import solve_ensemble
    from oracle.kbdm_oracle import kbdm_oracle
    nv = 160
    sig = workloads.c4_voxels_device(0, nv).cpu().numpy()
    res = solve_ensemble([sig[v] for v in range(nv)], [512] * nv, [512] * nv, 1, 0.0, DWELL)
    assert (res.status == 0).all()
    for v in (0, 1, 37, 80, 147, 148, 149, 159):
        _check_member(res.mu[v], res.D[v], res.sing_vals[v], kbdm_oracle(sig[v], DWELL, m=512, return_mu=True), f"C4 voxel {v}")


def test_c5_ragged_members_own_fids_against_oracle(cuda):
    """Config C5 shape: ragged m in [512, 1024], every member its own pseudo-noise draw of a 4096-point FID; sampled members
    (smallest, largest, two in between) against the oracle."""
    import torch
    from llckbdm_b200 import workloads
    from llckbdm_b200.ensemble import solve_ensemble, to_device_complex
    from oracle.kbdm_oracle import brain_sim, kbdm_oracle
    base = brain_sim(4096, 1e-3, 0)
    n = 24
    ms = workloads.c5_member_sizes(0, n, stride=37)
    sig = workloads.c5_members_device(to_device_complex(base, torch.device("cuda", 0)), 0, n).cpu().numpy()
    res = solve_ensemble([sig[k] for k in range(n)], ms, ms, 1, 0.0, DWELL)
    assert (res.status == 0).all()
    order = np.argsort(ms)
    for k in (order[0], order[n // 3], order[2 * n // 3], order[-1]):
        m = ms[k]
        _check_member(res.mu[k, :m], res.D[k, :m], res.sing_vals[k, :m], kbdm_oracle(sig[k], DWELL, m=m, return_mu=True), f"C5 member {k} m={m}")


@pytest.mark.parametrize("members", [149, 200])
def test_wave_boundary_ragged_batches(cuda, members):
    """149 and 200 members (one more than the 148 SMs; 1.35 waves) of ragged sizes in ONE launch sequence: every member against
    the oracle -- the second wave of the one-CTA-per-member kernels must be as correct as the first."""
    from llckbdm_b200.ensemble import solve_ensemble
    from oracle.kbdm_oracle import brain_sim, compare_members, kbdm_oracle
    c = brain_sim(1024, 1e-3, 17)
    rng = np.random.default_rng(members)
    ms = [int(x) for x in rng.integers(24, 129, members)]
    ms[0], ms[-1] = 128, 24
    ls = [m if k % 5 else max(1, m // 2) for k, m in enumerate(ms)]          # every fifth member truncated (l < m)
    res = solve_ensemble(c, ms, ls, 1, 0.0, DWELL, chunk=members)
    assert (res.status == 0).all()
    cache = {}
    for k, (m, l) in enumerate(zip(ms, ls)):
        if (m, l) not in cache:
            cache[(m, l)] = kbdm_oracle(c, DWELL, m=m, l=l, return_mu=True)
        _, info, mu, D = cache[(m, l)]
        dmu, dD = compare_members(res.mu[k, :l], res.D[k, :l], mu, D)
        assert dmu < TOL and dD < TOL, (k, m, l, dmu, dD)
        assert np.allclose(res.sing_vals[k, :m], info.singular_values, rtol=1e-8, atol=1e-12)


def _canonical_pool(lls):
    """Pooled samples with a canonical row order inside every member (eig order is arbitrary in both implementations)."""
    return np.concatenate([a[np.lexsort((a[:, 0], a[:, 2]))] for a in lls])


def test_llc_kbdm_full_clustering_loop_gpu_vs_oracle(cuda):
    """LLC-KBDM parity end to end (north_star: clustered estimates within 1e-6, identical cluster membership): 12 noisy members,
    the reference's whole min_samples = 1..M-1 loop.
    (1) GPU line lists and oracle line lists, pooled in the same canonical row order, get IDENTICAL labels from every fit, and
        the device clustering stage (spanning trees + native labelling) gives the same labels as the library clusterer;
    (2) hence identical cluster averages (rel 1e-6, every row of every candidate) and the same min-RMSE choice;
    (3) llc_kbdm itself pools in the solver's own row order (like the reference pools in LAPACK's eig order, which no two
        implementations share); HDBSCAN's tie handling depends on the point order, so the weak noise clusters move with it --
        permuting the rows inside the ORACLE's own members changes them just the same (checked here).  The well-conditioned
        lines (A > 1e-2 max A) of llc_kbdm's result agree with the oracle-derived result to 1e-6."""
    from llckbdm_b200 import llckbdm as L
    from llckbdm_b200.sampling import filter_samples, sample_kbdm
    from oracle.kbdm_oracle import brain_sim, min_rmse_oracle, sample_kbdm_oracle
    c = brain_sim(1024, 1e-3, 5)
    m_range = list(range(100, 112))
    gpu, _ = sample_kbdm(c, DWELL, m_range, p=1, l=None)
    ref, _ = sample_kbdm_oracle(c, DWELL, m_range, p=1, l=None)
    assert [len(a) for a in gpu] == [len(a) for a in ref]
    pools = [filter_samples(_canonical_pool(lls)) for lls in (gpu, ref)]
    feats = [L._transform_line_lists(s, DWELL) for s in pools]
    ks = list(range(1, len(m_range)))
    lab_gpu_lines = [L._fit_one(feats[0], k) for k in ks]          # library clusterer on GPU line lists
    lab_ref_lines = [L._fit_one(feats[1], k) for k in ks]          # library clusterer on oracle line lists
    lab_device = L._fit_all(feats[0], ks)                          # device spanning trees + native labelling on GPU line lists
    assert L._gpu_fit_supported(feats[0], ks)
    for k, a, b, d in zip(ks, lab_gpu_lines, lab_ref_lines, lab_device):
        assert np.array_equal(a, b), f"min_samples={k}: labels differ between GPU and oracle line lists"
        assert np.array_equal(a, d), f"min_samples={k}: device clustering stage differs from the library fit"

    def candidates(samples, labelings):
        out = []
        for labels in labelings:
            nc = len(set(labels.tolist()) - {-1})
            if nc:
                out.append(L._summarize_clusters(samples, [np.nonzero(labels == j) for j in range(nc)]))
        return out

    cand_gpu, cand_ref = candidates(pools[0], lab_device), candidates(pools[1], lab_ref_lines)
    assert len(cand_gpu) == len(cand_ref) > 0
    for a, b in zip(cand_gpu, cand_ref):
        assert a.shape == b.shape
        assert np.allclose(a[:, [0, 2]], b[:, [0, 2]], rtol=1e-6, atol=1e-9) and np.allclose(a[:, 1], b[:, 1], rtol=1e-6)
    kg, rg = min_rmse_oracle(c, DWELL, cand_gpu)
    kr, rr = min_rmse_oracle(c, DWELL, cand_ref)
    assert kg == kr and np.allclose(rg, rr, rtol=1e-6)

    # (3) the public call against the reference's pipeline restated on the oracle's line lists, both in their own row order
    def strong(ll):
        ll = ll[ll[:, 0] > 1e-2 * ll[:, 0].max()]
        return ll[np.argsort(ll[:, 2])]

    samples = filter_samples(np.concatenate(ref))
    f = L._transform_line_lists(samples, DWELL)
    cands = candidates(samples, [L._fit_one(f, k) for k in ks])
    kbest, rmses = min_rmse_oracle(c, DWELL, cands)
    res = L.llc_kbdm(c, DWELL, m_range)
    got, want = strong(res.line_list), strong(cands[kbest])
    assert got.shape == want.shape and len(got) >= 12
    assert np.allclose(got[:, [0, 2]], want[:, [0, 2]], rtol=1e-6, atol=1e-9)
    assert np.allclose(got[:, 1], want[:, 1], rtol=1e-6)
    assert abs(res.rmse - rmses[kbest]) < 1e-3 * rmses[kbest]
    # the same sensitivity inside the oracle itself: a row permutation within its members moves only weak rows
    rng = np.random.default_rng(1)
    shuffled = filter_samples(np.concatenate([x[rng.permutation(len(x))] for x in ref]))
    fs = L._transform_line_lists(shuffled, DWELL)
    cs = candidates(shuffled, [L._fit_one(fs, k) for k in ks])
    ksh, _ = min_rmse_oracle(c, DWELL, cs)
    assert np.allclose(strong(cs[ksh])[:, [0, 2]], want[:, [0, 2]], rtol=1e-6, atol=1e-9)


def test_iterative_llc_kbdm_device_residual_matches_host_restatement(cuda, capsys):
    """iterative_llc_kbdm keeps data, estimate and residual on the device; a literal host restatement of reference
    llckbdm.py:144-199 (residual and multi_fid in numpy, the product's llc_kbdm per iteration) must give the same lines."""
    from llckbdm_b200 import llckbdm as L
    from llckbdm_b200 import sig_gen
    from llckbdm_b200.metrics import calculate_freq_domain_rmse
    from oracle.kbdm_oracle import brain_sim
    c = brain_sim(1024, 1e-3, 2)
    m_range = range(100, 108)
    res = L.iterative_llc_kbdm(c, DWELL, m_range=m_range, max_iterations=3)
    assert "Iteration #0" in capsys.readouterr().out
    est = np.zeros_like(c)
    t, _ = sig_gen.gen_t_freq_arrays(len(c), DWELL)
    thresholds = np.linspace(0.6, 0, 3)
    lists = []
    for it in range(3):
        r = L.llc_kbdm(c - est, DWELL, m_range)
        if len(r.line_list) == 0:
            break
        keep = np.nonzero(r.silhouette > np.percentile(r.silhouette, thresholds[it]))
        ll = r.line_list[keep]
        est = est + sig_gen.multi_fid(t, ll)
        lists.append(ll)
    assert len(res.line_lists) == len(lists) >= 1
    for a, b in zip(res.line_lists, lists):
        assert a.shape == b.shape
        assert np.allclose(a[:, [0, 2]], b[:, [0, 2]], rtol=1e-6, atol=1e-9) and np.allclose(a[:, 1], b[:, 1], rtol=1e-6)
    want_rmse = calculate_freq_domain_rmse(est, np.concatenate(lists), DWELL)
    assert abs(res.rmse - want_rmse) <= 1e-6 * max(want_rmse, 1e-12) + 1e-13
    assert len(res.line_list) == sum(len(x) for x in lists)


def test_distributed_path_world1_equals_single_gpu(cuda):
    """The sharded product path (LPT shard, chunked local solve with a forced small chunk, record packing, re-assembly) with one
    rank must reproduce ensemble.solve_ensemble exactly; per-member FIDs and a shared FID."""
    from llckbdm_b200.distributed import solve_ensemble_distributed
    from llckbdm_b200.ensemble import solve_ensemble
    from oracle.kbdm_oracle import brain_sim
    ms = [100, 257, 64, 300, 33, 129, 200]
    ls = [100, 200, 64, 300, 33, 30, 200]
    sigs = [brain_sim(700, 1e-3, 40 + k) for k in range(len(ms))]
    for signals in (sigs, brain_sim(2048, 1e-3, 3)):
        stats = {}
        d = solve_ensemble_distributed(signals, ms, ls, 1, 0.0, DWELL, chunk=3, stats=stats)
        s = solve_ensemble(signals, ms, ls, 1, 0.0, DWELL, chunk=3)
        assert np.array_equal(d["status"], s.status) and np.array_equal(d["n_valid"], s.n_valid) and (s.status == 0).all()
        assert np.array_equal(d["line_lists"], s.line_lists) and np.array_equal(d["sing_vals"], s.sing_vals)
        assert stats["shard_sizes"] == [len(ms)] and stats["allgather_bytes_per_rank"] == len(ms) * (8 * (4 * 300 + 300) + 8)


def _nccl_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from llckbdm_b200.distributed import sample_kbdm_distributed, solve_ensemble_distributed
        from oracle.kbdm_oracle import brain_sim
        c = brain_sim(2048, 1e-3, 3)
        m_range = [100, 257, 64, 300, 33, 129, 200, 150, 96]
        lls, infos = sample_kbdm_distributed(c, DWELL, m_range, p=1, l=None)
        sigs = [brain_sim(700, 1e-3, 40 + k) for k in range(len(m_range))]
        r = solve_ensemble_distributed(sigs, m_range, m_range, 1, 0.0, DWELL, chunk=2)
        np.savez(os.path.join(out, f"r{rank}.npz"), n=len(lls), sv=np.concatenate([i.singular_values for i in infos]),
                 ll=np.concatenate(lls), ll2=r["line_lists"], st2=r["status"], shard=np.array(r["shards"][rank]))
    finally:
        dist.destroy_process_group()


def test_two_rank_nccl_sample_kbdm_distributed_equals_single_gpu(cuda, tmp_path):
    """sample_kbdm_distributed on 2 GPUs (NCCL all_gather of the packed records) == sampling.sample_kbdm on one, on every rank."""
    torch = cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2); the one-rank path is covered by test_distributed_path_world1_equals_single_gpu")
    import torch.multiprocessing as mp
    from llckbdm_b200.ensemble import solve_ensemble
    from llckbdm_b200.sampling import sample_kbdm
    from oracle.kbdm_oracle import brain_sim
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    c = brain_sim(2048, 1e-3, 3)
    m_range = [100, 257, 64, 300, 33, 129, 200, 150, 96]
    lls, infos = sample_kbdm(c, DWELL, m_range, p=1, l=None)
    sigs = [brain_sim(700, 1e-3, 40 + k) for k in range(len(m_range))]
    single = solve_ensemble(sigs, m_range, m_range, 1, 0.0, DWELL)
    def close(a, b):
        """Shards are other batches than the single-GPU run (other thread-block cluster sizes, other summation splits):
        compare the well-conditioned lines (sorted by frequency) instead of bits."""
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape
        big = b[:, 0] > 1e-3 * b[:, 0].max()
        fa, fb = a[big], b[big]
        fa, fb = fa[np.argsort(fa[:, 2])], fb[np.argsort(fb[:, 2])]
        return np.allclose(fa[:, [0, 2]], fb[:, [0, 2]], rtol=1e-8, atol=1e-10) and np.allclose(fa[:, 1], fb[:, 1], rtol=1e-7)

    shards = []
    for rank in range(2):
        r = np.load(tmp_path / f"r{rank}.npz")
        assert int(r["n"]) == len(lls)
        cuts = np.cumsum([len(a) for a in lls])[:-1]
        for a, b in zip(np.split(r["ll"], cuts), lls):
            assert close(a, b)
        assert np.allclose(r["sv"], np.concatenate([i.singular_values for i in infos]), rtol=1e-9, atol=1e-13)
        assert np.array_equal(r["st2"], single.status)
        for k, m in enumerate(m_range):
            assert close(r["ll2"][k, :m], single.line_lists[k, :m])
        shards.append(list(r["shard"]))
    assert sorted(shards[0] + shards[1]) == list(range(len(m_range))) and shards[0] and shards[1]


def test_solve_call_is_asynchronous_and_stream_ordered(cuda):
    """llck_kbdm_batched returns once the launch sequence is enqueued: with the stream held back by a long sleep kernel the call
    must come back long before the stream drains, and the results are valid after a synchronize."""
    torch = cuda
    from llckbdm_b200 import ensemble
    from oracle.kbdm_oracle import brain_sim, compare_members, kbdm_oracle
    dev = torch.device("cuda", 0)
    c = brain_sim(1024, 1e-3, 4)
    sig = ensemble.to_device_complex(c, dev)
    nb = 148
    ms = [384 + (k % 5) for k in range(nb)]                                      # ~0.3 s of device work
    zeros = [0] * nb
    r = ensemble.solve_device(sig, zeros, ms, ms, 1, 0.0, DWELL)                 # warm-up: module load, attribute calls, allocations
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = ensemble.solve_device(sig, zeros, ms, ms, 1, 0.0, DWELL)
    t_call = time.perf_counter() - t0
    t0 = time.perf_counter()
    torch.cuda.synchronize()
    t_wait = time.perf_counter() - t0
    assert t_wait > 0.05, (t_call, t_wait)                                      # the stream was still busy when the call returned ...
    assert t_call < 0.5 * (t_call + t_wait), (t_call, t_wait)                   # ... and for longer than the call itself took
    assert int(r["status"].abs().sum().item()) == 0
    for k in (0, 77, 147):
        m = ms[k]
        _, _, mu, D = kbdm_oracle(c, DWELL, m=m, return_mu=True)
        dmu, dD = compare_members(r["mu"][k, :m].cpu().numpy(), r["D"][k, :m].cpu().numpy(), mu, D)
        assert dmu < TOL and dD < TOL
    assert r["info"][13] > 0 and r["info"][2] == 448 and r["info"][14] == 1
    # the cached workspace and the parked graphs can be released and come back on demand
    ensemble.release_workspace()
    r = ensemble.solve_device(sig, zeros[:3], ms[:3], ms[:3], 1, 0.0, DWELL)
    torch.cuda.synchronize()
    assert int(r["status"].abs().sum().item()) == 0


def test_short_signal_is_an_error_not_an_out_of_bounds_read(cuda):
    """sig_len is validated per member: a list input whose FID is shorter than 2m + p - 1 raises, nothing is launched."""
    from llckbdm_b200.ensemble import solve_ensemble
    from oracle.kbdm_oracle import brain_sim
    sigs = [brain_sim(256, 1e-3, 1), brain_sim(100, 1e-3, 2)]
    with pytest.raises(ValueError, match="shorter than the 2m"):
        solve_ensemble(sigs, [64, 64], [64, 64], 1, 0.0, DWELL)
    res = solve_ensemble(sigs, [64, 50], [64, 50], 1, 0.0, DWELL)
    assert (res.status == 0).all()


def test_rmse_scoring_long_fid_tiled_kernel(cuda):
    """N above the shared-memory limit of the one-tile scoring kernel (12800 points): the tiled kernel against the oracle's
    fft-based restatement of metrics.py:7-17; even and odd N."""
    from llckbdm_b200.ensemble import score_candidates
    from oracle.kbdm_oracle import BRAIN_SIM_PARAMS, brain_sim, filter_samples_oracle, freq_domain_rmse_oracle
    rng = np.random.default_rng(8)
    for N in (16384, 13001, 20000):
        data = brain_sim(N, 1e-3, 3)
        cand = np.column_stack([rng.random(30) - 0.1, rng.random(30) * 0.2 - 0.01, rng.uniform(-900, 900, 30), rng.uniform(-3, 3, 30)])
        want = freq_domain_rmse_oracle(data, filter_samples_oracle(cand), DWELL)
        got = score_candidates(data, DWELL, [cand, BRAIN_SIM_PARAMS, np.zeros((0, 4))], filter_rows=True)
        assert abs(got[0] - want) < 1e-10 * want, (N, got[0], want)
        want_truth = freq_domain_rmse_oracle(data, BRAIN_SIM_PARAMS, DWELL)
        assert abs(got[1] - want_truth) < 1e-9 * want_truth
        assert got[2] == np.inf


def test_m_above_supported_maximum_raises_descriptive_error(cuda):
    from llckbdm_b200.kbdm import kbdm
    with pytest.raises(ValueError, match="largest Hankel dimension the CUDA solver supports"):
        kbdm(np.ones(5000, dtype=complex), DWELL, m=2100)


def test_options_struct_selects_the_svd_back_end(cuda):
    """llck_options replaces the environment switches: Jacobi SVD for every member vs the default divide and conquer."""
    from llckbdm_b200 import _native
    from llckbdm_b200.ensemble import solve_ensemble
    from oracle.kbdm_oracle import brain_sim, compare_members, kbdm_oracle
    c = brain_sim(512, 1e-3, 9)
    ms = [70, 33, 128]
    a = solve_ensemble(c, ms, ms, 1, 0.0, DWELL, options=_native.Options(svd_mode=_native.SVD_JACOBI))
    b = solve_ensemble(c, ms, ms, 1, 0.0, DWELL, options=_native.Options(cluster_size=1, aed_window=16))
    for res in (a, b):
        assert (res.status == 0).all()
        for k, m in enumerate(ms):
            _, info, mu, D = kbdm_oracle(c, DWELL, m=m, return_mu=True)
            dmu, dD = compare_members(res.mu[k, :m], res.D[k, :m], mu, D)
            assert dmu < TOL and dD < TOL, (k, dmu, dD)
            assert np.allclose(res.sing_vals[k, :m], info.singular_values, rtol=1e-8, atol=1e-12)
    assert ctypes.sizeof(_native.Options) == 40


def test_jacobi_fallback_runs_as_device_side_while_graph(cuda):
    """Members the divide-and-conquer SVD flags as rank deficient (noiseless FIDs) are finished by Jacobi sweeps that a CUDA-graph
    WHILE node repeats on the device until convergence -- no host read-back.  Same results as with every sweep enqueued
    (LLCK_FLAG_NO_GRAPH); noisy members in the same batch keep oracle parity; the 16 true components of the noiseless member are found."""
    from llckbdm_b200 import _native
    from llckbdm_b200.ensemble import solve_ensemble
    from oracle.kbdm_oracle import BRAIN_SIM_PARAMS, brain_sim, compare_members, kbdm_oracle
    sigs = [brain_sim(512, 1e-3, 1), brain_sim(512, 0.0, 0), brain_sim(512, 1e-3, 2), brain_sim(512, 0.0, 0)]
    ms = [200, 256, 130, 97]
    a = solve_ensemble(sigs, ms, ms, 1, 0.0, DWELL)
    b = solve_ensemble(sigs, ms, ms, 1, 0.0, DWELL, flags=_native.FLAG_NO_GRAPH)
    assert a.info["chunks"][0][14] == 1 and b.info["chunks"][0][14] == 0
    assert a.info["chunks"][0][13] + 500 < b.info["chunks"][0][13]           # kernels enqueued: the graph launch counts once
    for res in (a, b):
        assert (res.status == 0).all()
        for k in (0, 2):
            _, info, mu, D = kbdm_oracle(sigs[k], DWELL, m=ms[k], return_mu=True)
            dmu, dD = compare_members(res.mu[k, :ms[k]], res.D[k, :ms[k]], mu, D)
            assert dmu < TOL and dD < TOL, (k, dmu, dD)
        for k in (1, 3):
            ll = res.line_lists[k, :ms[k]]
            est = ll[(ll[:, 0] > 1e-4) & (ll[:, 1] > 0)]
            est = est[np.argsort(est[:, 2])]
            assert len(est) == 16
            assert np.allclose(est[:, 0], BRAIN_SIM_PARAMS[:, 0], rtol=1e-6) and np.allclose(est[:, 2], BRAIN_SIM_PARAMS[:, 2], atol=1e-6)
    assert np.allclose(a.sing_vals[1, :16], b.sing_vals[1, :16], rtol=1e-10)


def test_pooled_path_in_several_chunks_keeps_member_order(cuda):
    """solve_pooled with the ensemble split into cost-sorted chunks (as the scheduler does for ensembles larger than the HBM or
    with a wave remainder): the pooled rows must still come in m_range order and equal the one-chunk result bit for bit per member."""
    from llckbdm_b200.ensemble import solve_pooled
    from oracle.kbdm_oracle import brain_sim
    c = brain_sim(1024, 1e-3, 9)
    ms = [40, 300, 64, 129, 17, 200, 33]
    s1, f1, st1 = solve_pooled(c, ms, ms, 1, 0.0, DWELL)
    s3, f3, st3 = solve_pooled(c, ms, ms, 1, 0.0, DWELL, chunk=3)
    assert (st1 == 0).all() and (st3 == 0).all()
    assert s1.shape == s3.shape and f1.shape == f3.shape
    # chunks are other batches (other thread-block cluster sizes): compare per member, well-conditioned rows, to parity tolerance
    big1, big3 = s1[:, 0] > 1e-3 * s1[:, 0].max(), s3[:, 0] > 1e-3 * s3[:, 0].max()
    assert np.array_equal(big1, big3)
    assert np.allclose(s1[big1][:, [0, 2]], s3[big3][:, [0, 2]], rtol=1e-8, atol=1e-10)
    assert np.allclose(f1[big1], f3[big3], rtol=1e-8, atol=1e-10)


def test_kbdm_p2_q_truncated_at_c2_size(cuda):
    """Shift p = 2, Tikhonov q > 0 and l < m at a config-C2 size (m = 700, l = 200) against the oracle: the options of
    reference kbdm.py:19 are not only exercised at toy sizes."""
    from llckbdm_b200.ensemble import solve_ensemble
    from oracle.kbdm_oracle import brain_sim, compare_members, kbdm_oracle
    c = brain_sim(2048, 1e-3, 0)
    res = solve_ensemble(c, [700, 700], [200, 700], 2, 1e-3, DWELL)
    assert (res.status == 0).all()
    for k, l in enumerate((200, 700)):
        _, info, mu, D = kbdm_oracle(c, DWELL, m=700, l=l, p=2, q=1e-3, return_mu=True)
        dmu, dD = compare_members(res.mu[k, :l], res.D[k, :l], mu, D)
        assert dmu < TOL and dD < TOL, (l, dmu, dD)
        assert np.allclose(res.sing_vals[k, :700], info.singular_values, rtol=1e-8, atol=1e-12)


def test_two_stream_schedule_for_a_ragged_partial_wave(cuda):
    """100 ragged members (more than half a wave, less than one): the scheduler runs the 48 largest with a 2-CTA cluster each on a
    second stream next to the 52 others on the caller's stream.  Same results as the plain one-stream launch sequence (parity
    tolerance: other cluster sizes, other summation splits) and oracle parity on sampled members of both groups."""
    from llckbdm_b200 import _native
    from llckbdm_b200.ensemble import solve_ensemble, two_stream_split
    from oracle.kbdm_oracle import brain_sim, compare_members, kbdm_oracle
    c = brain_sim(1024, 1e-3, 21)
    ms = [60 + 2 * k for k in range(100)]
    assert two_stream_split(np.array(ms), np.array(ms), 148) is not None
    a = solve_ensemble(c, ms, ms, 1, 0.0, DWELL)                                              # two streams
    b = solve_ensemble(c, ms, ms, 1, 0.0, DWELL, options=_native.Options(cluster_size=1))     # explicit cluster size: one stream
    assert (a.status == 0).all() and (b.status == 0).all() and np.array_equal(a.n_valid, b.n_valid)
    for k in (0, 30, 51, 52, 77, 99):
        m = ms[k]
        _, info, mu, D = kbdm_oracle(c, DWELL, m=m, return_mu=True)
        for res in (a, b):
            dmu, dD = compare_members(res.mu[k, :m], res.D[k, :m], mu, D)
            assert dmu < TOL and dD < TOL, (k, dmu, dD)
        assert np.allclose(a.sing_vals[k, :m], info.singular_values, rtol=1e-8, atol=1e-12)


def test_maximum_size_m2048_properties(cuda):
    """The largest supported Hankel dimension (LLCK_M_MAX = 2048, N = 4096): singular values against LAPACK, generalized-eigen residual
    of sampled poles, signal reconstruction -- the oracle's eig at this size would take a minute, the properties do not need it."""
    from llckbdm_b200.ensemble import solve_ensemble
    from oracle.kbdm_oracle import brain_sim, hankel_matrices
    c = brain_sim(4096, 1e-3, 0)
    m = 2048
    res = solve_ensemble(c, [m], [m], 1, 0.0, DWELL)
    assert res.status[0] == 0
    U0, _, U1 = hankel_matrices(c, m, 1)
    s_ref = np.linalg.svd(U0, compute_uv=False)
    assert np.allclose(res.sing_vals[0], s_ref, rtol=1e-8, atol=1e-12)
    mu, D = res.mu[0], res.D[0]
    for k in (0, 1000, 2047):
        smin = np.linalg.svd(U1 - mu[k] * U0, compute_uv=False)[-1]
        assert smin < 1e-9 * s_ref[0]
    n = np.arange(64)
    recon = (D[None, :] * mu[None, :] ** n[:, None]).sum(axis=1)
    assert np.abs(recon - c[:64]).max() < 1e-6 * np.abs(c).max()


def test_empty_and_degenerate_ensembles(cuda):
    """Empty m_range, a one-member ensemble, and members that all filter to nothing are handled like the reference does
    (sampling.py:52-72: empty lists; empty members are dropped)."""
    from llckbdm_b200.sampling import sample_kbdm, sample_kbdm_pooled
    from llckbdm_b200.llckbdm import llc_kbdm
    from oracle.kbdm_oracle import brain_sim
    c = brain_sim(512, 1e-3, 3)
    assert sample_kbdm(c, DWELL, [], p=1, l=None) == ([], [])
    s, f = sample_kbdm_pooled(c, DWELL, [], p=1, l=None)
    assert s.shape == (0, 4) and f.shape == (0, 4)
    lls, infos = sample_kbdm(c, DWELL, [17], p=1, l=None)
    assert len(lls) == 1 and infos[0].m == 17
    tiny = 1e-9 * c                                         # every amplitude below the 1e-6 filter threshold: all members dropped
    lls, infos = sample_kbdm(tiny, DWELL, [20, 30], p=1, l=None)
    assert lls == [] and infos == []
    r = llc_kbdm(tiny, DWELL, [20, 30, 40])
    assert len(r.line_list) == 0 and r.rmse is None


@pytest.mark.parametrize("name", ["llc_clean_m250_l30", "llc_clean_m100_l40"])
def test_llc_kbdm_matches_the_real_reference_driver_golden(cuda, name):
    """llc_kbdm against the output of the REAL reference's llckbdm.llc_kbdm (tests/golden/llc_*.npz, generated by
    oracle/gen_golden_llc.py with the documented hdbscan stub = sklearn HDBSCAN(min_samples=k+1)); the first case is the
    reference's own test_llc_kbdm input (_tests/test_llckbdm.py:37-57).  Lines with A > 1e-3 (the reference test's own filter):
    same count, amplitudes / T2 / frequencies rel 1e-6, phases abs 1e-8."""
    from llckbdm_b200.llckbdm import llc_kbdm
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    res = llc_kbdm(g["data"], float(g["dwell"]), [int(x) for x in g["m_range"]], l=int(g["l"]))

    def strong(ll):
        ll = ll[ll[:, 0] > 1e-3]
        return ll[np.argsort(ll[:, 2])]

    got, want = strong(res.line_list), strong(g["line_list"])
    assert got.shape == want.shape == (16, 4)
    assert np.allclose(got[:, :3], want[:, :3], rtol=1e-6, atol=0)
    assert np.allclose(got[:, 3], want[:, 3], atol=1e-8)
    assert res.rmse < 1e-9


def test_truncated_rank_at_headline_size_m1024_l30(cuda):
    """m = 1024 with only l = 30 singular triplets kept (the SVD-bound shape of BASELINE.md §2): poles, amplitudes and ALL 1024
    singular values against the oracle."""
    from llckbdm_b200.ensemble import solve_ensemble
    from oracle.kbdm_oracle import brain_sim, compare_members, kbdm_oracle
    c = brain_sim(2048, 1e-3, 0)
    res = solve_ensemble(c, [1024], [30], 1, 0.0, DWELL)
    assert res.status[0] == 0
    _, info, mu, D = kbdm_oracle(c, DWELL, m=1024, l=30, return_mu=True)
    dmu, dD = compare_members(res.mu[0, :30], res.D[0, :30], mu, D)
    assert dmu < TOL and dD < TOL, (dmu, dD)
    assert info.singular_values.shape == (1024,)
    assert np.allclose(res.sing_vals[0], info.singular_values, rtol=1e-8, atol=1e-12)
