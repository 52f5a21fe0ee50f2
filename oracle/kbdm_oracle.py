"""CPU oracle for the KBDM per-member solve -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy/scipy restatement of the hot path of danilomendesdias/llckbdm (reference file:line cited per
function).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; ``llckbdm_b200`` never does (its product path is
the CUDA extension and fails loudly without it).

Parity status: PINNED.  ``oracle/gen_golden.py`` ran the *real* reference (imported from
/root/reference with the two shims of SURVEY.md §8c) in the build container and committed its
outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks this restatement against those
vectors and against the reference's own known-answer test (16 brain_sim components,
``llckbdm/_tests/test_kbdm.py:8-42``).

The arithmetic of the reference lives in LAPACK (``scipy.linalg.svd`` -> zgesdd, ``kbdm.py:166``;
``scipy.linalg.eig`` -> zgeev, ``kbdm.py:192``); the reference pins neither scipy nor numpy
(``requirements.txt:1-2`` lower bounds only).  This oracle calls the same two scipy entry points.
"""
import numpy as np
from scipy.linalg import svd, eig


class OracleInfo:
    """Mirror of ``KbdmInfo`` (reference ``llckbdm/kbdm.py:10-16``)."""

    def __init__(self, m, l, p, q, singular_values):
        self.m, self.l, self.p, self.q, self.singular_values = m, l, p, q, singular_values


def hankel_matrices(data, m, p):
    """U0[i,j]=c[i+j], U^{p-1}[i,j]=c[i+j+p-1], U^p[i,j]=c[i+j+p]  (reference ``kbdm.py:95-130``)."""
    idx = np.arange(m)[:, None] + np.arange(m)[None, :]
    data = np.asarray(data, dtype=complex)
    return data[idx], data[idx + p - 1], data[idx + p]


def reduce_gep(Up_1, Up, l, q=0.0):
    """SVD of U^{p-1}, truncation to l, Dsqi, reduced operator (reference ``kbdm.py:166-189``).

    Returns (U_red [l,l], R_l [m,l], dsqi [l], s [m])."""
    L, s, Rh = svd(Up_1)
    L_ = L[:, :l]
    R_ = Rh[:l, :].conj().T
    s_ = s[:l]
    if q > 0:
        g = s_ + q * q / s_          # kbdm.py:179-184 (inv of a diagonal matrix == reciprocal)
    else:
        g = s_
    if np.any(g == 0):
        raise np.linalg.LinAlgError("Singular matrix")   # np.linalg.inv behaviour, kbdm.py:186
    dsqi = 1.0 / np.sqrt(g)
    U_red = (dsqi[:, None] * (L_.conj().T @ Up @ R_)) * dsqi[None, :]
    return U_red, R_, dsqi, s


def normalise(B, U0, how="gemm"):
    """N_k = B_k^T U0 B_k (bilinear, no conjugate), B_hat = B / sqrt(N)  (reference ``kbdm.py:215-240``).

    ``how='einsum'`` reproduces the reference's 3-operand einsum literally (slow, used for the timed
    CPU baseline); ``how='gemm'`` is the same contraction as a GEMM + column dot (used in tests)."""
    if how == "einsum":
        n = np.einsum('jk,ij,ik->k', B, U0, B)
    else:
        n = ((U0 @ B) * B).sum(axis=0)
    return B * np.sqrt(1.0 / n)


def kbdm_oracle(data, dwell, m=None, p=1, l=None, q=0, how="gemm", return_mu=False):
    """One KBDM solve (reference ``kbdm.py:19-92``): returns (line_list float64[l,4], OracleInfo[, mu, D])."""
    if m is None and l is None:
        raise ValueError("l or m must be specified")
    elif m is None:
        m = l
    elif l is None:
        l = m
    elif l > m:
        raise ValueError("l can't be greater than m")
    m_max = (data.size + 1 - p) / 2
    if m > m_max or l > m_max:
        raise ValueError("m or l can't be greater than (n + 1 - p)/2.")

    U0, Up_1, Up = hankel_matrices(data, m, p)
    U_red, R_, dsqi, s = reduce_gep(Up_1, Up, l, q)
    mu, P = eig(U_red)                                   # kbdm.py:192
    B = R_ @ (dsqi[:, None] * P)                         # kbdm.py:198
    B_norm = normalise(B, U0, how)                       # kbdm.py:202
    D_sqrt = np.asarray(data[:m], dtype=complex) @ B_norm   # kbdm.py:71
    D = D_sqrt * D_sqrt
    A = np.abs(D)
    PH = np.angle(D)
    with np.errstate(divide='ignore', invalid='ignore'):
        omega = -1j * np.log(mu) / dwell                 # kbdm.py:82
        F = omega.real / (2 * np.pi)
        T2 = 1.0 / omega.imag
    ll = np.column_stack((A, T2, F, PH))
    info = OracleInfo(m=m, l=l, p=p, q=q, singular_values=s)
    if return_mu:
        return ll, info, mu, D
    return ll, info


def filter_samples_oracle(samples, amplitude_tol=1e-6):
    """Reference ``sampling.py:75-97``."""
    if len(samples) == 0:
        return samples
    return samples[(samples[:, 0] > amplitude_tol) & (samples[:, 1] > 0)]


def sample_kbdm_oracle(data, dwell, m_range, p, l, q=0, filter_invalid_features=True, how="gemm"):
    """Reference ``sampling.py:8-72`` (serial loop, empty members dropped)."""
    lls, infos = [], []
    for m in m_range:
        ll, info = kbdm_oracle(data, dwell, m=m, p=p, l=l, q=q, how=how)
        if filter_invalid_features:
            ll = filter_samples_oracle(ll)
        if len(ll) > 0:
            lls.append(ll)
            infos.append(info)
    return lls, infos


# ----------------------------------------------------------------------------------------------
# inputs and the parity comparator
# ----------------------------------------------------------------------------------------------

# data/params_brain_sim_1_5T.csv of the reference, sorted by frequency as _tests/fixtures.py:23-35 does.
BRAIN_SIM_PARAMS = np.array([
    [1.0, 0.002712968, 75.31704, 0.0],
    [0.11611, 0.0138504155, 160.06464, 0.0],
    [0.291727, 0.0199203187, 246.46896, 0.0],
    [0.428882, 0.0735294118, 255.5172, 0.0],
    [0.0290276, 0.0066489362, 268.8984, 0.0],
    [0.0184325, 0.0909090909, 269.5356, 0.0],
    [0.0450798, 0.0833333333, 290.43576, 0.0],
    [0.0427286, 0.1162790698, 299.99376, 0.0],
    [0.202612, 0.0925925926, 386.7804, 0.0],
    [0.0777794, 0.1136363636, 410.22936, 0.0],
    [0.0201887, 0.1052631579, 414.94464, 0.0],
    [0.0411176, 0.1470588235, 455.08824, 0.0],
    [0.0150218, 0.2222222222, 464.5188, 0.0],
    [0.105428, 0.0456621005, 482.3604, 0.0],
    [0.299129, 0.04, 503.388, 0.0],
    [0.824383, 0.0087950748, 525.30768, 0.0],
])


def multi_fid_oracle(t, params):
    """Reference ``sig_gen.py:27-71``: sum_k a_k exp(-t/T2_k) exp(i(2 pi f_k t + ph_k))."""
    out = np.zeros(len(t), dtype=complex)
    for a, t2, f, ph in params:
        out = out + a * np.exp(-t / t2) * np.exp(1j * (2 * np.pi * f * t + ph))
    return out


def freq_domain_rmse_oracle(data, params_est, dwell):
    """Reference ``metrics.py:7-17``: RMSE of the REAL parts of fft(data)/sqrt(N) and fft(model)/sqrt(N), the model being
    ``multi_fid`` of the candidate line list on ``t = np.arange(0, N*dwell, dwell)`` (``sig_gen.py:8-24``)."""
    N = len(data)
    t = np.arange(0, N * dwell, dwell)
    est = np.fft.fft(multi_fid_oracle(t, params_est)) / np.sqrt(N)
    ref = np.fft.fft(np.asarray(data, dtype=complex)) / np.sqrt(N)
    return float(np.sqrt(np.mean((ref.real - est.real) ** 2)))


def min_rmse_oracle(data, dwell, samples):
    """Reference ``min_rmse_kbdm.py:33-55``: (argmin index, rmse list); empty candidates score inf."""
    rmses = [freq_domain_rmse_oracle(data, ll, dwell) if len(ll) > 0 else np.inf for ll in samples]
    return int(np.argmin(rmses)), rmses


def brain_sim(N=2048, sigma=1e-3, seed=0, dwell=5e-4, params=BRAIN_SIM_PARAMS):
    """brain_sim(N, sigma, seed) of SURVEY.md §8(d): the 16-component FID (+ seeded complex noise)."""
    t = np.linspace(0, dwell * N, N, endpoint=False)     # _tests/fixtures.py:18-20
    c = multi_fid_oracle(t, params)
    if sigma > 0:
        rng = np.random.default_rng(seed)
        c = c + sigma * (rng.standard_normal(N) + 1j * rng.standard_normal(N))
    return c


def mu_from_line_list(ll, dwell):
    """Invert kbdm.py:82-85: mu = exp(i*dwell*(2 pi F + i/T2))."""
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.exp(1j * dwell * (2 * np.pi * ll[:, 2] + 1j / ll[:, 1]))


def match_rows(mu_a, mu_b):
    """Assignment of rows of a to rows of b by nearest pole (Hungarian on |mu_a - mu_b|)."""
    from scipy.optimize import linear_sum_assignment
    cost = np.abs(mu_a[:, None] - mu_b[None, :])
    r, c = linear_sum_assignment(cost)
    out = np.empty(len(mu_a), dtype=int)
    out[r] = c
    return out


def compare_members(mu_a, D_a, mu_b, D_b, amp_floor=1e-3):
    """Returns (max |dmu|/|mu| over all rows, max |dD|/|D| over rows with |D| > amp_floor*max|D|)."""
    perm = match_rows(mu_a, mu_b)
    mu_b = mu_b[perm]
    D_b = D_b[perm]
    dmu = np.max(np.abs(mu_a - mu_b) / np.abs(mu_b))
    big = np.abs(D_b) > amp_floor * np.max(np.abs(D_b))
    dD = np.max(np.abs(D_a[big] - D_b[big]) / np.abs(D_b[big])) if big.any() else 0.0
    return dmu, dD


def flops_per_solve(m, l):
    """Algorithmic real-FP64 flop model F(m,l) of SURVEY.md §8(d)."""
    return 53.0 * m ** 3 + 8.0 * l * m * m + 16.0 * l * l * m + 108.0 * l ** 3 + 8.0 * m * m * l
