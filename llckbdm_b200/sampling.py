"""Drop-in for ``llckbdm.sampling`` (reference llckbdm/sampling.py).  ``sample_kbdm`` keeps the
reference's list-in / list-out contract, but all members are solved by ONE batched GPU launch
sequence instead of the serial ``for m in m_range`` loop (sampling.py:52-70)."""
import logging

import numpy as np

from .ensemble import solve_ensemble, solve_pooled
from .kbdm import KbdmInfo, check_finite, raise_for_status, resolve_m_l

logger = logging.getLogger(__name__)


def filter_samples(samples, amplitude_tol=1e-6):
    """Keep rows with amplitude > amplitude_tol and T2 > 0 (reference sampling.py:75-97)."""
    if len(samples) == 0:
        return samples
    keep = (samples[:, 0] > amplitude_tol) & (samples[:, 1] > 0)
    return samples[keep]


def sample_kbdm(data, dwell, m_range, p, l, q=0, filter_invalid_features=True):
    """KBDM for every m in m_range.

    Returns (line_lists, infos) in m_range order; members whose (filtered) line list is empty are
    omitted, exactly as the reference does (sampling.py:67-70).
    """
    line_lists, infos, _ = sample_kbdm_scored(data, dwell, m_range, p, l, q, filter_invalid_features, score_rmse=False)
    return line_lists, infos


def sample_kbdm_scored(data, dwell, m_range, p, l, q=0, filter_invalid_features=True, score_rmse=True):
    """``sample_kbdm`` plus, per kept member, the frequency-domain RMSE of its filtered line list against ``data`` computed on
    the device from the solver's output (what reference min_rmse_kbdm.py:33-41 computes in a CPU loop afterwards).
    Returns (line_lists, infos, rmses)."""
    ms, ls = [], []
    for m in m_range:
        logger.info(f'Computing KBDM with m = {m}')
        mm, ll_ = resolve_m_l(data.size, m, l, p)
        check_finite(data, mm, p)
        ms.append(mm)
        ls.append(ll_)
    if not ms:
        return [], [], []
    if q > 0:
        logger.debug('Using Tikhonov Regularization with q=%f', q)
    res = solve_ensemble(np.asarray(data).ravel(), ms, ls, p, q, dwell, score_rmse=score_rmse and filter_invalid_features)
    line_lists, infos, rmses = [], [], []
    for k, (mm, ll_) in enumerate(zip(ms, ls)):
        raise_for_status(int(res.status[k]), mm)
        line_list = np.ascontiguousarray(res.line_lists[k, :ll_, :])
        if filter_invalid_features:
            line_list = filter_samples(line_list)
        if len(line_list) > 0:
            line_lists.append(line_list)
            infos.append(KbdmInfo(m=mm, l=ll_, p=p, q=q, singular_values=res.sing_vals[k, :mm].copy()))
            if res.rmse is not None:
                rmses.append(float(res.rmse[k]))
    return line_lists, infos, rmses


def sample_kbdm_pooled(data, dwell, m_range, p, l, q=0):
    """What reference llckbdm.py:76-98 builds on the host -- ``sample_kbdm`` over m_range, np.concatenate, ``filter_samples`` and
    ``_transform_line_lists`` -- in one device pass: returns (samples [n, 4], features [n, 4])."""
    ms, ls = [], []
    for m in m_range:
        logger.info(f'Computing KBDM with m = {m}')
        mm, ll_ = resolve_m_l(data.size, m, l, p)
        check_finite(data, mm, p)
        ms.append(mm)
        ls.append(ll_)
    if not ms:
        return np.zeros((0, 4)), np.zeros((0, 4))
    if q > 0:
        logger.debug('Using Tikhonov Regularization with q=%f', q)
    samples, features, status = solve_pooled(np.asarray(data).ravel(), ms, ls, p, q, dwell)
    for k, mm in enumerate(ms):
        raise_for_status(int(status[k]), mm)
    return samples, features
