"""Generate tests/golden/*.npz by running the REAL reference (read-only import from /root/reference).

Runs only in the build container (the GPU box has no /root/reference).  Shims:
``numpy.complex = complex`` (removed in numpy >= 1.24, used at reference kbdm.py:111-113).
Only kbdm.py / sampling.py / sig_gen.py are imported -- they do not need hdbscan.

    python oracle/gen_golden.py
"""
import os
import sys

import numpy as np

np.complex = complex  # shim 1
sys.path.insert(0, '/root/reference')
from llckbdm.kbdm import kbdm  # noqa: E402
from llckbdm.sampling import sample_kbdm  # noqa: E402
from llckbdm import sig_gen  # noqa: E402
from llckbdm.metrics import calculate_freq_domain_rmse  # noqa: E402
from llckbdm.min_rmse_kbdm import min_rmse_kbdm  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '..', 'tests', 'golden')


def brain_sim_ref(N, sigma, seed, dwell=5e-4):
    import pandas as pd
    df = pd.read_csv('/root/reference/data/params_brain_sim_1_5T.csv',
                     names=['amplitude', 't2', 'frequency', 'phase']).sort_values(['frequency'])
    t = np.linspace(0, dwell * N, N, endpoint=False)
    c = sig_gen.multi_fid(t, df.values)
    if sigma > 0:
        rng = np.random.default_rng(seed)
        c = c + sigma * (rng.standard_normal(N) + 1j * rng.standard_normal(N))
    return c, df.values


# (name, N, sigma, seed, m, p, l, q)
CASES = [
    ("noisy_m16", 256, 1e-3, 0, 16, 1, None, 0.0),
    ("noisy_m64", 512, 1e-3, 0, 64, 1, None, 0.0),
    ("noisy_m128", 2048, 1e-3, 0, 128, 1, None, 0.0),
    ("noisy_m200_l30", 2048, 1e-3, 1, 200, 1, 30, 0.0),
    ("noisy_m96_p2", 2048, 1e-3, 2, 96, 2, None, 0.0),
    ("noisy_m80_q", 2048, 1e-3, 3, 80, 1, None, 1e-3),
    ("noisy_m100_l40_p2_q", 2048, 1e-2, 4, 100, 2, 40, 1e-2),
    ("noisy_m256", 2048, 1e-3, 0, 256, 1, None, 0.0),
    ("clean_m150", 2048, 0.0, 0, 150, 1, None, 0.0),
    ("clean_m180_l30", 2048, 0.0, 0, 180, 1, 30, 0.0),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, N, sigma, seed, m, p, l, q in CASES:
        c, params = brain_sim_ref(N, sigma, seed)
        ll, info = kbdm(c, 5e-4, m=m, p=p, l=l, q=q)
        np.savez_compressed(os.path.join(OUT, f"kbdm_{name}.npz"),
                            data=c, dwell=5e-4, m=m, p=p, l=-1 if l is None else l, q=q,
                            N=N, sigma=sigma, seed=seed,
                            line_list=ll, singular_values=info.singular_values)
        print(name, ll.shape)
    # the brain_sim FID itself (pins oracle.brain_sim against reference sig_gen + CSV)
    c, params = brain_sim_ref(2048, 0.0, 0)
    cn, _ = brain_sim_ref(2048, 1e-3, 0)
    np.savez_compressed(os.path.join(OUT, "brain_sim_fid.npz"), clean=c, noisy=cn, params=params)
    # sample_kbdm: reference test shape (_tests/test_min_rmse_kbdm.py:6-23) incl. filtering
    lls, infos = sample_kbdm(c, 5e-4, [30, 31, 180, 32, 33, 34], p=1, l=30, q=0, filter_invalid_features=True)
    np.savez_compressed(os.path.join(OUT, "sample_kbdm_minrmse.npz"),
                        n=len(lls), **{f"ll{i}": x for i, x in enumerate(lls)},
                        **{f"sv{i}": inf.singular_values for i, inf in enumerate(infos)})
    lls, infos = sample_kbdm(cn, 5e-4, range(100, 104), p=1, l=None, q=0, filter_invalid_features=True)
    np.savez_compressed(os.path.join(OUT, "sample_kbdm_noisy.npz"),
                        n=len(lls), **{f"ll{i}": x for i, x in enumerate(lls)},
                        **{f"sv{i}": inf.singular_values for i, inf in enumerate(infos)})
    # frequency-domain RMSE scoring (metrics.py:7-17) and the min_rmse_kbdm selection (min_rmse_kbdm.py:21-56) on the noisy members
    rm = [calculate_freq_domain_rmse(cn, ll, 5e-4) for ll in lls]
    res = min_rmse_kbdm(cn, 5e-4, samples=lls)
    rm_truth = calculate_freq_domain_rmse(cn, params, 5e-4)
    np.savez_compressed(os.path.join(OUT, "rmse_noisy.npz"), rmses=np.array(rm), min_index=res.min_index, min_rmse=res.min_rmse,
                        rmse_truth=rm_truth)
    print("done")


if __name__ == '__main__':
    main()
