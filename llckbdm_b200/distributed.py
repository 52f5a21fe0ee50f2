"""Multi-GPU ensemble solve: one process per GPU (torch.distributed), members sharded by cost.

Ensemble members are independent (reference llckbdm/sampling.py:52 has no cross-iteration state), so
the data path needs NO collective: each rank solves its LPT shard.  The single exchange step is one
``all_gather`` of fixed-stride result buffers so that every rank holds the complete, ``m_range``-ordered
line lists for the host-side clustering (reference llckbdm/llckbdm.py:94-124).  With the NCCL backend
the gather runs over NVLink 5 / NVSwitch on device tensors; the gloo backend (CPU tensors) is used
by the CPU tests, which inject a solver stub.
"""
import numpy as np

from .ensemble import flops_per_solve, lpt_shards, solve_device, flatten_signals


def _pack(line_lists, sing_vals, n_valid, status, lmax, mmax, count, torch, device):
    """Fixed-stride per-rank buffer: [count, lmax*4 + mmax + 2] float64 (status and n_valid stored as doubles)."""
    width = lmax * 4 + mmax + 2
    buf = torch.zeros((count, width), dtype=torch.float64, device=device)
    k = line_lists.shape[0]
    if k:
        buf[:k, :line_lists.shape[1] * 4] = line_lists.reshape(k, -1)
        buf[:k, lmax * 4:lmax * 4 + sing_vals.shape[1]] = sing_vals
        buf[:k, lmax * 4 + mmax] = n_valid.to(torch.float64)
        buf[:k, lmax * 4 + mmax + 1] = status.to(torch.float64)
    return buf


def solve_ensemble_distributed(signals, m, l, p, q, dwell, group=None, local_solver=None, device=None):
    """Sharded ensemble solve.  Every rank passes the SAME arguments; every rank returns the full result
    (line_lists float64[M,lmax,4], sing_vals float64[M,mmax], n_valid int[M], status int[M]) in member order.

    local_solver(signals_flat, offsets, m, l, p, q, dwell) -> dict of torch tensors is injectable for CPU tests.
    """
    import torch
    import torch.distributed as dist
    M = len(m)
    m = np.asarray(m, dtype=np.int32)
    l = np.asarray(l, dtype=np.int32)
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    shards = lpt_shards([flops_per_solve(mi, li) for mi, li in zip(m, l)], world)
    mine = np.asarray(shards[rank], dtype=np.int64)
    count = max(len(s) for s in shards)
    mmax, lmax = int(m.max()), int(l.max())
    flat, offsets = flatten_signals(signals, M)
    if local_solver is None:
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        sig_dev = torch.from_numpy(flat.view(np.float64)).to(dev).view(torch.complex128)

        def local_solver(idx):
            return solve_device(sig_dev, offsets[idx], m[idx], l[idx], p, q, dwell, want_mu=False)
    else:
        dev = torch.device("cpu") if device is None else torch.device(device)
        user_solver = local_solver

        def local_solver(idx):
            return user_solver(flat, offsets[idx], m[idx], l[idx], p, q, dwell)
    if len(mine):
        r = local_solver(mine)
        buf = _pack(r["line_lists"], r["sing_vals"], r["n_valid"], r["status"], lmax, mmax, count, torch, dev)
    else:
        buf = torch.zeros((count, lmax * 4 + mmax + 2), dtype=torch.float64, device=dev)
    if world > 1:
        gathered = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(gathered, buf, group=group)          # the ONE exchange step
    else:
        gathered = [buf]
    out_ll = np.zeros((M, lmax, 4))
    out_sv = np.zeros((M, mmax))
    out_nv = np.zeros(M, dtype=np.int32)
    out_st = np.zeros(M, dtype=np.int32)
    for rk, idx in enumerate(shards):
        if not idx:
            continue
        g = gathered[rk][:len(idx)].cpu().numpy()
        out_ll[idx] = g[:, :lmax * 4].reshape(len(idx), lmax, 4)
        out_sv[idx] = g[:, lmax * 4:lmax * 4 + mmax]
        out_nv[idx] = g[:, lmax * 4 + mmax].astype(np.int32)
        out_st[idx] = g[:, lmax * 4 + mmax + 1].astype(np.int32)
    return dict(line_lists=out_ll, sing_vals=out_sv, n_valid=out_nv, status=out_st, shards=shards)


def sample_kbdm_distributed(data, dwell, m_range, p, l, q=0, filter_invalid_features=True, group=None):
    """``sampling.sample_kbdm`` with the members sharded over the ranks of ``group`` (same return contract)."""
    from .kbdm import KbdmInfo, raise_for_status, resolve_m_l
    from .sampling import filter_samples
    ms, ls = [], []
    for mm in m_range:
        a, b = resolve_m_l(data.size, mm, l, p)
        ms.append(a)
        ls.append(b)
    if not ms:
        return [], []
    res = solve_ensemble_distributed(np.asarray(data).ravel(), ms, ls, p, q, dwell, group=group)
    line_lists, infos = [], []
    for k, (mm, ll_) in enumerate(zip(ms, ls)):
        raise_for_status(int(res["status"][k]), mm)
        line_list = np.ascontiguousarray(res["line_lists"][k, :ll_, :])
        if filter_invalid_features:
            line_list = filter_samples(line_list)
        if len(line_list) > 0:
            line_lists.append(line_list)
            infos.append(KbdmInfo(m=mm, l=ll_, p=p, q=q, singular_values=res["sing_vals"][k, :mm].copy()))
    return line_lists, infos
