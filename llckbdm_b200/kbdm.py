"""Drop-in for ``llckbdm.kbdm`` (reference llckbdm/kbdm.py): same signature, return types and errors;
the solve itself is one call of the CUDA batched solver (batch of one)."""
import logging

import attr
import numpy as np

from . import _native
from .ensemble import solve_ensemble

logger = logging.getLogger(__name__)


@attr.s
class KbdmInfo:
    """Record returned next to the line list (reference kbdm.py:10-16): the same attrs class, singular_values has length m."""
    m = attr.ib()
    l = attr.ib()
    p = attr.ib()
    q = attr.ib()
    singular_values = attr.ib()


def resolve_m_l(data_size, m, l, p):
    """Argument defaults and validation, behaviour of reference kbdm.py:50-62 (same messages)."""
    if m is None and l is None:
        raise ValueError("l or m must be specified")
    if m is None:
        m = l
    elif l is None:
        l = m
    elif l > m:
        raise ValueError("l can't be greater than m")
    m_max = (data_size + 1 - p) / 2
    if m > m_max or l > m_max:
        raise ValueError("m or l can't be greater than (n + 1 - p)/2.")
    if m > _native.M_MAX:          # limit of this implementation (LLCK_M_MAX, include/llck.h); the reference has none
        raise ValueError(f"m = {int(m)} is above the largest Hankel dimension the CUDA solver supports ({_native.M_MAX})")
    return int(m), int(l)


def check_finite(data, m, p):
    """The reference's scipy.linalg.svd / eig calls run with check_finite=True (kbdm.py:166,192): a NaN or Inf among the points the
    Hankel matrices use, c[0 .. 2m+p-2], raises this ValueError there; same here, before anything is launched."""
    used = np.asarray(data).ravel()[:2 * int(m) + int(p) - 1]
    if not np.isfinite(used).all():
        raise ValueError("array must not contain infs or NaNs")


def raise_for_status(status, m=None):
    """Per-member numerical failure -> numpy.linalg.LinAlgError, like np.linalg.inv / scipy.linalg.eig would raise
    inside the reference (kbdm.py:186,192)."""
    where = "" if m is None else f" (m={m})"
    if status == _native.STATUS_OK:
        return
    if status == _native.STATUS_SINGULAR:
        raise np.linalg.LinAlgError("Singular matrix" + where)
    if status == _native.STATUS_QR_NOCONV:
        raise np.linalg.LinAlgError("eig algorithm (multishift QR) did not converge" + where)
    if status == _native.STATUS_SVD_NOCONV:
        raise np.linalg.LinAlgError("SVD did not converge" + where)
    if status == _native.STATUS_NONFINITE:
        raise np.linalg.LinAlgError("non-finite pole or amplitude" + where)
    raise np.linalg.LinAlgError(f"solver status {status}" + where)


def kbdm(data, dwell, m=None, p=1, l=None, q=0):
    """Krylov Basis Diagonalization Method on one FID.

    :param numpy.ndarray data: complex (or real) time-domain signal.
    :param float dwell: sampling interval in seconds.
    :param int|None m: Hankel dimension (rows/columns of the U matrices).
    :param int p: pencil shift (U^p B = mu U^{p-1} B); default 1.
    :param int|None l: number of singular triplets kept (l <= m); default m.
    :param float q: Tikhonov parameter; 0 disables it.
    :return: (line_list float64[l,4] with columns (A, T2, F, PH), KbdmInfo)
    """
    m, l = resolve_m_l(data.size, m, l, p)
    check_finite(data, m, p)
    if q > 0:
        logger.debug('Using Tikhonov Regularization with q=%f', q)
    res = solve_ensemble(np.asarray(data).ravel(), [m], [l], p, q, dwell)
    raise_for_status(int(res.status[0]), m)
    info = KbdmInfo(m=m, l=l, p=p, q=q, singular_values=res.sing_vals[0, :m].copy())
    return np.ascontiguousarray(res.line_lists[0, :l, :]), info
