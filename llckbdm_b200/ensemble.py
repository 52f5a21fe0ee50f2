"""Ensemble scheduler: the batched GPU solve behind ``kbdm`` / ``sample_kbdm``.

``solve_ensemble`` runs every member of an ensemble (one FID + one Hankel size m each) through ONE
call of the C-ABI ``llck_kbdm_batched`` per chunk -- this replaces the serial
``for m in m_range: kbdm(...)`` loop of the reference (llckbdm/sampling.py:52-70).

Multi-GPU (one process per GPU, torch.distributed): members are independent, so they are sharded
longest-processing-time-first by the cost model F(m,l) and each rank solves its shard with no
data-path collective; the only exchange step is one all_gather of the fixed-stride result buffers
(line lists, singular values, status) so that every rank can rebuild the Python lists in
``m_range`` order for the CPU clustering stage (llckbdm/llckbdm.py:94-124).
"""
import ctypes

import numpy as np

from . import _native


def flops_per_solve(m, l):
    """Algorithmic real-FP64 flop model of one solve."""
    m = float(m)
    l = float(l)
    return 53.0 * m ** 3 + 8.0 * l * m * m + 16.0 * l * l * m + 108.0 * l ** 3 + 8.0 * m * m * l


def lpt_shards(costs, world_size):
    """Longest-processing-time-first assignment of members to ranks. Returns list of index lists."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += costs[i]
    return [sorted(s) for s in shards]


class EnsembleResult:
    """Padded, fixed-stride results of a batch (host numpy arrays)."""

    def __init__(self, line_lists, mu, D, sing_vals, n_valid, status, m, l, info=None):
        self.line_lists = line_lists    # float64 [M, lmax, 4]
        self.mu = mu                    # complex128 [M, lmax]
        self.D = D                      # complex128 [M, lmax]
        self.sing_vals = sing_vals      # float64 [M, mmax]
        self.n_valid = n_valid          # int32 [M]
        self.status = status            # int32 [M]
        self.m = m
        self.l = l
        self.info = info or {}


def _require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("llckbdm_b200 requires a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch


def max_chunk(ld, device=None, reserve_frac=0.8, flags=0):
    """Largest number of members of leading dimension ld whose workspace fits in free HBM."""
    torch = _require_cuda()
    free, _total = torch.cuda.mem_get_info(device)
    lib = _native.load()
    per = lib.llck_workspace_bytes(1, ld, flags)
    return max(1, int(free * reserve_frac // per))


def solve_device(signals_dev, sig_offset, m, l, p, q, dwell, flags=0, workspace=None, stream=None, want_mu=True):
    """Run llck_kbdm_batched on device-resident signals.

    signals_dev: torch complex128 CUDA tensor (flat); sig_offset/m/l: host int sequences (one per member).
    Returns dict of torch CUDA tensors (line_lists, mu, D, sing_vals, n_valid, status) + info list + workspace.
    """
    torch = _require_cuda()
    lib = _native.load()
    batch = len(m)
    dev = signals_dev.device
    m_arr = np.ascontiguousarray(m, dtype=np.int32)
    l_arr = np.ascontiguousarray(l, dtype=np.int32)
    off_arr = np.ascontiguousarray(sig_offset, dtype=np.int64)
    mmax, lmax = int(m_arr.max()), int(l_arr.max())
    ld = lib.llck_leading_dim(mmax)
    need = lib.llck_workspace_bytes(batch, ld, flags)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    line_lists = torch.empty((batch, lmax, 4), dtype=torch.float64, device=dev)
    mu = torch.empty((batch, lmax), dtype=torch.complex128, device=dev) if want_mu else None
    D = torch.empty((batch, lmax), dtype=torch.complex128, device=dev) if want_mu else None
    sv = torch.empty((batch, mmax), dtype=torch.float64, device=dev)
    n_valid = torch.empty(batch, dtype=torch.int32, device=dev)
    status = torch.empty(batch, dtype=torch.int32, device=dev)
    info = (ctypes.c_int32 * 16)()
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    rc = lib.llck_kbdm_batched(
        signals_dev.data_ptr(), off_arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
        m_arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), l_arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
        int(p), float(q), float(dwell), batch,
        line_lists.data_ptr(), lmax * 4,
        mu.data_ptr() if want_mu else None, D.data_ptr() if want_mu else None, lmax,
        sv.data_ptr(), mmax,
        n_valid.data_ptr(), status.data_ptr(),
        workspace.data_ptr(), need, int(flags),
        st.cuda_stream, info)
    _native.check_rc(rc, "llck_kbdm_batched")
    return dict(line_lists=line_lists, mu=mu, D=D, sing_vals=sv, n_valid=n_valid, status=status,
                info=list(info), workspace=workspace, ld=ld)


def flatten_signals(signals, M):
    """One shared 1-D FID, or a list of M FIDs -> (flat complex128 array, int64 offsets)."""
    if isinstance(signals, np.ndarray) and signals.ndim == 1:
        return np.ascontiguousarray(signals, dtype=np.complex128), np.zeros(M, dtype=np.int64)
    sigs = [np.ascontiguousarray(s, dtype=np.complex128) for s in signals]
    if len(sigs) != M:
        raise ValueError("need one signal per member")
    lens = np.array([len(s) for s in sigs], dtype=np.int64)
    offsets = np.concatenate(([0], np.cumsum(lens)[:-1])).astype(np.int64)
    return np.concatenate(sigs), offsets


def solve_ensemble(signals, m, l, p, q, dwell, device=None, chunk=None):
    """Solve an ensemble given HOST inputs; returns an ``EnsembleResult`` of host arrays.

    signals: either one 1-D complex array shared by all members, or a list of 1-D complex arrays (one per member).
    The host->device copy of the FIDs and the device->host copy of the results are part of this call.
    """
    torch = _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    M = len(m)
    flat, offsets = flatten_signals(signals, M)
    m = np.asarray(m, dtype=np.int32)
    l = np.asarray(l, dtype=np.int32)
    mmax, lmax = int(m.max()), int(l.max())
    lib = _native.load()
    ld = lib.llck_leading_dim(mmax)
    with torch.cuda.device(dev):
        if chunk is None:
            chunk = min(M, max_chunk(ld, dev))
        sig_dev = torch.from_numpy(flat.view(np.float64)).to(dev).view(torch.complex128)
        out_ll = np.zeros((M, lmax, 4))
        out_mu = np.zeros((M, lmax), dtype=np.complex128)
        out_D = np.zeros((M, lmax), dtype=np.complex128)
        out_sv = np.zeros((M, mmax))
        out_nv = np.zeros(M, dtype=np.int32)
        out_st = np.zeros(M, dtype=np.int32)
        ws = None
        infos = []
        # cost-sorted chunks keep similar sizes together (less padding work inside a launch)
        order = np.argsort(-(m.astype(np.int64) * 4096 + l), kind="stable")
        for c0 in range(0, M, chunk):
            idx = order[c0:c0 + chunk]
            r = solve_device(sig_dev, offsets[idx], m[idx], l[idx], p, q, dwell, workspace=ws)
            ws = r["workspace"]
            lm, mm = r["line_lists"].shape[1], r["sing_vals"].shape[1]
            out_ll[idx, :lm] = r["line_lists"].cpu().numpy()
            out_mu[idx, :lm] = r["mu"].cpu().numpy()
            out_D[idx, :lm] = r["D"].cpu().numpy()
            out_sv[idx, :mm] = r["sing_vals"].cpu().numpy()
            out_nv[idx] = r["n_valid"].cpu().numpy()
            out_st[idx] = r["status"].cpu().numpy()
            infos.append(r["info"])
    return EnsembleResult(out_ll, out_mu, out_D, out_sv, out_nv, out_st, m, l, info={"chunks": infos, "ld": ld})
