"""Frequency-domain RMSE (drop-in for reference llckbdm/metrics.py:7-17; cheap CPU post-processing)."""
import numpy as np

from .sig_gen import gen_t_freq_arrays, multi_fid


def calculate_freq_domain_rmse(data, params_est, dwell):
    N = len(data)
    t_array, _ = gen_t_freq_arrays(N=N, dwell=dwell)
    est = np.fft.fft(multi_fid(t_array=t_array, params=params_est)) / np.sqrt(N)
    ref = np.fft.fft(data) / np.sqrt(N)
    return np.sqrt(np.mean((ref.real - est.real) ** 2))      # RMSE of the REAL parts only (metrics.py:17)
