"""CPU tests: the numpy oracle against the golden vectors produced by the REAL reference
(oracle/gen_golden.py) and against the reference's own known-answer tests."""
import glob
import os

import numpy as np
import pytest

from oracle.kbdm_oracle import (BRAIN_SIM_PARAMS, brain_sim, compare_members, filter_samples_oracle, hankel_matrices,
                                kbdm_oracle, mu_from_line_list, sample_kbdm_oracle, flops_per_solve)

DWELL = 5e-4


def _cases(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "kbdm_*.npz")))


def _D(ll):
    return ll[:, 0] * np.exp(1j * ll[:, 3])


def test_brain_sim_matches_reference_fid(golden_dir):
    g = np.load(os.path.join(golden_dir, "brain_sim_fid.npz"))
    assert np.array_equal(g["params"], BRAIN_SIM_PARAMS)
    assert np.abs(brain_sim(2048, 0.0, 0) - g["clean"]).max() < 1e-14
    assert np.abs(brain_sim(2048, 1e-3, 0) - g["noisy"]).max() < 1e-14


@pytest.mark.parametrize("name", ["noisy_m16", "noisy_m64", "noisy_m128", "noisy_m200_l30", "noisy_m96_p2", "noisy_m80_q",
                                  "noisy_m100_l40_p2_q", "noisy_m256"])
def test_oracle_matches_reference_golden_noisy(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"kbdm_{name}.npz"))
    l = None if int(g["l"]) < 0 else int(g["l"])
    ll, info = kbdm_oracle(g["data"], float(g["dwell"]), m=int(g["m"]), p=int(g["p"]), l=l, q=float(g["q"]))
    assert ll.shape == g["line_list"].shape
    dmu, dD = compare_members(mu_from_line_list(ll, DWELL), _D(ll), mu_from_line_list(g["line_list"], DWELL), _D(g["line_list"]))
    assert dmu < 1e-10 and dD < 1e-9            # tolerance: relative, all poles; D for |D| > 1e-3 max|D|
    assert np.allclose(info.singular_values, g["singular_values"], rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("name", ["clean_m150", "clean_m180_l30"])
def test_oracle_matches_reference_golden_clean_true_components(golden_dir, name):
    """Noiseless input has numerical rank 16: only the 16 true components are well defined."""
    g = np.load(os.path.join(golden_dir, f"kbdm_{name}.npz"))
    l = None if int(g["l"]) < 0 else int(g["l"])
    ll, _ = kbdm_oracle(g["data"], DWELL, m=int(g["m"]), p=1, l=l, q=0.0)
    a = filter_samples_oracle(ll)
    a = a[a[:, 0] > 1e-4]
    b = g["line_list"][(g["line_list"][:, 0] > 1e-4) & (g["line_list"][:, 1] > 0)]
    a = a[np.argsort(a[:, 2])]
    b = b[np.argsort(b[:, 2])]
    assert len(a) == len(b) == 16
    assert np.allclose(a[:, [0, 2]], b[:, [0, 2]], rtol=1e-7)
    assert np.allclose(a[:, 1], b[:, 1], rtol=1e-6)


def test_known_answer_16_components():
    """Reference llckbdm/_tests/test_kbdm.py:8-42."""
    c = brain_sim(2048, 0.0, 0)
    ll, info = kbdm_oracle(c, DWELL, m=300)
    assert ll.shape == (300, 4) and info.m == 300 and info.l == 300 and info.p == 1
    est = ll[ll[:, 0] > 1e-4]
    est = est[np.argsort(est[:, 2])]
    assert len(est) == 16
    assert np.allclose(est[:, 0], BRAIN_SIM_PARAMS[:, 0], rtol=1e-6)
    assert np.allclose(est[:, 1], BRAIN_SIM_PARAMS[:, 1], rtol=1e-3)
    assert np.allclose(est[:, 2], BRAIN_SIM_PARAMS[:, 2], atol=0.3)
    assert np.allclose(est[:, 3], 0.0, atol=1e-10)


def test_hankel_rows_reference_test_compute_U_matrices():
    """Reference llckbdm/_tests/test_kbdm.py:45-59."""
    c = brain_sim(2048, 0.0, 0)
    m, p = 300, 2
    U0, Up1, Up = hankel_matrices(c, m, p)
    assert np.array_equal(U0[0], c[:m]) and np.array_equal(Up1[0], c[p - 1:m + p - 1]) and np.array_equal(Up[0], c[p:m + p])
    assert np.array_equal(U0[m - 1], c[m - 1:2 * m - 1]) and np.array_equal(Up[m - 1], c[m + p - 1:2 * m + p - 1])


def test_sample_kbdm_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "sample_kbdm_minrmse.npz"))
    c = brain_sim(2048, 0.0, 0)
    lls, infos = sample_kbdm_oracle(c, DWELL, [30, 31, 180, 32, 33, 34], p=1, l=30)
    assert len(lls) == int(g["n"]) == 6
    for i in range(6):
        assert np.allclose(infos[i].singular_values, g[f"sv{i}"], rtol=1e-9, atol=1e-12)
    # spurious rows of this rank-deficient (noiseless) input are not reproducible between LAPACK call orders;
    # the well-conditioned member (m=180) must agree on its true components
    a = lls[2][lls[2][:, 0] > 1e-4]
    b = g["ll2"][g["ll2"][:, 0] > 1e-4]
    assert len(a) == len(b) == 16
    assert np.allclose(np.sort(a[:, 2]), np.sort(b[:, 2]), rtol=1e-8)


def test_validation_messages():
    c = brain_sim(256, 0.0, 0)
    with pytest.raises(ValueError, match="l or m must be specified"):
        kbdm_oracle(c, DWELL)
    with pytest.raises(ValueError, match="l can't be greater than m"):
        kbdm_oracle(c, DWELL, l=30, m=20)
    with pytest.raises(ValueError, match=r"m or l can't be greater than \(n \+ 1 - p\)/2."):
        kbdm_oracle(c, DWELL, m=129)


def test_einsum_and_gemm_normalisation_agree():
    c = brain_sim(256, 1e-3, 0)
    a, _ = kbdm_oracle(c, DWELL, m=24, how="einsum")
    b, _ = kbdm_oracle(c, DWELL, m=24, how="gemm")
    assert np.allclose(a[:, [0, 2, 3]], b[:, [0, 2, 3]], rtol=1e-9, atol=1e-12)


def test_flop_model():
    assert flops_per_solve(1024, 1024) == pytest.approx(193 * 1024 ** 3)


def test_rmse_oracle_matches_reference_golden(golden_dir):
    """freq_domain_rmse_oracle / min_rmse_oracle vs values produced by the real reference (metrics.py:7-17, min_rmse_kbdm.py:21-56)."""
    from oracle.kbdm_oracle import BRAIN_SIM_PARAMS, freq_domain_rmse_oracle, min_rmse_oracle
    fid = np.load(os.path.join(golden_dir, "brain_sim_fid.npz"))
    g = np.load(os.path.join(golden_dir, "sample_kbdm_noisy.npz"))
    r = np.load(os.path.join(golden_dir, "rmse_noisy.npz"))
    lls = [g[f"ll{i}"] for i in range(int(g["n"]))]
    k, rmses = min_rmse_oracle(fid["noisy"], 5e-4, lls)
    assert np.allclose(rmses, r["rmses"], rtol=1e-10, atol=0)
    assert k == int(r["min_index"]) and abs(rmses[k] - float(r["min_rmse"])) <= 1e-10 * float(r["min_rmse"])
    assert abs(freq_domain_rmse_oracle(fid["noisy"], BRAIN_SIM_PARAMS, 5e-4) - float(r["rmse_truth"])) <= 1e-10 * float(r["rmse_truth"])
    assert min_rmse_oracle(fid["noisy"], 5e-4, [np.zeros((0, 4)), lls[0]])[1][0] == np.inf
