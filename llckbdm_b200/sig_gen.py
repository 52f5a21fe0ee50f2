"""Synthetic FID / Lorentzian generators (drop-in for reference llckbdm/sig_gen.py; host-side input
generators, not part of the GPU hot path)."""
import logging

import numpy as np

logger = logging.getLogger(__name__)


def gen_t_freq_arrays(N, dwell):
    """Time axis 0, dwell, ... (N points) and fft-shifted frequency axis (reference sig_gen.py:8-24)."""
    return np.arange(0, N * dwell, dwell), np.fft.fftshift(np.fft.fftfreq(N, dwell))


def _validate_parameters(a, t2, f, phase):
    """Reference sig_gen.py:140-169: T2 > 0, a >= 0, warn when |phase| > 2 pi."""
    if t2 <= 0:
        raise ValueError("T2 must be positive.")
    if a < 0:
        raise ValueError("Amplitude can't be negative.")
    if np.abs(phase) > 2 * np.pi:
        logger.warning('Phase is greater than 2 * pi and phase must be given in rad/s. '
                       'Check whether the correct unit is being used.')


def fid(t_array, a, t2, f, phase=0.):
    """a * exp(-t/T2) * exp(i (2 pi f t + phase))  (reference sig_gen.py:27-54)."""
    _validate_parameters(a, t2, f, phase)
    return a * np.exp(-t_array / t2) * np.exp(1j * (2 * np.pi * f * t_array + phase))


def multi_fid(t_array, params):
    """Sum of FIDs; params rows are (amplitude, t2, frequency, phase) (reference sig_gen.py:57-71)."""
    return np.sum([fid(t_array, *param) for param in params], axis=0)


def fft(data):
    """Shifted FFT normalised by sqrt(N) (reference sig_gen.py:74-88)."""
    return np.fft.fftshift(np.fft.fft(data)) / np.sqrt(len(data))


def lorentzian_peak(freq_array, a, t2, f, phase=0):
    """a e^{i phase} / (1/T2 + 2 pi i (nu - f))  (reference sig_gen.py:91-121)."""
    _validate_parameters(a, t2, f, phase)
    return a * np.exp(1j * phase) / ((1. / t2) + 2j * np.pi * (freq_array - f))


def spec(freq_array, params):
    """Sum of Lorentzian peaks (reference sig_gen.py:124-137)."""
    return np.sum([lorentzian_peak(freq_array, *param) for param in params], axis=0)


def multi_fid_batched_device(params, N, dwell, device=None):
    """``multi_fid`` for many parameter sets at once on the GPU (llck_multi_fid_batched): ``params`` is a list of [K_b, 4] arrays
    (or one [B, K, 4] array) of (amplitude, t2, frequency, phase) rows; returns a complex128 CUDA tensor [B, N] on
    ``t = n * dwell`` (the grid of ``gen_t_freq_arrays``).  Rows are validated like ``fid`` does (sig_gen.py:140-169)."""
    import torch
    from . import _native
    if not torch.cuda.is_available():
        raise RuntimeError("llckbdm_b200 requires a CUDA device (B200, sm_100a); there is no CPU fallback.")
    sets = [np.asarray(p, dtype=np.float64).reshape(-1, 4) for p in params]
    for ps in sets:
        if (ps[:, 1] <= 0).any():
            raise ValueError("T2 must be positive.")
        if (ps[:, 0] < 0).any():
            raise ValueError("Amplitude can't be negative.")
    B = len(sets)
    rows = np.array([len(ps) for ps in sets], dtype=np.int32)
    kmax = max(1, int(rows.max())) if B else 1
    packed = np.zeros((B, kmax, 4))
    for i, ps in enumerate(sets):
        packed[i, :len(ps)] = ps
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    lib = _native.load()
    with torch.cuda.device(dev):
        out = torch.empty((B, int(N)), dtype=torch.complex128, device=dev)
        for b0 in range(0, B, 65535):
            b1 = min(B, b0 + 65535)
            pd = torch.from_numpy(packed[b0:b1]).to(dev)
            rd = torch.from_numpy(rows[b0:b1]).to(dev)
            rc = lib.llck_multi_fid_batched(pd.data_ptr(), kmax * 4, rd.data_ptr(), b1 - b0, int(N), float(dwell),
                                            out[b0:b1].data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
            _native.check_rc(rc, "llck_multi_fid_batched")
    return out
