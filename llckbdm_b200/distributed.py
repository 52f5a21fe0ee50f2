"""Multi-GPU ensemble solve: one process per GPU (torch.distributed), members sharded by cost.

Ensemble members are independent (reference llckbdm/sampling.py:52 has no cross-iteration state), so
the data path needs NO collective: each rank uploads only the FIDs of its own longest-processing-time-first
shard and solves it in memory-sized chunks (``ensemble.solve_chunks``).  The single exchange step is one
``all_gather`` of the fixed-stride result buffer (line lists + singular values + int32 n_valid/status packed
into one byte buffer per rank) so that every rank holds the complete, ``m_range``-ordered line lists for the
clustering stage (reference llckbdm/llckbdm.py:94-124).  With the NCCL backend the gather runs over
NVLink 5 / NVSwitch on device tensors and the gathered buffer comes back to the host in ONE copy; the gloo
backend (CPU tensors) is used by the CPU tests, which inject a solver stub.
"""
import numpy as np

from .ensemble import flops_per_solve, lpt_shards, solve_chunks, to_device_complex


def record_bytes(lmax, mmax):
    """Bytes of one member's fixed-stride record: line list [lmax, 4] f64 | singular values [mmax] f64 | n_valid, status int32."""
    return 8 * (4 * lmax + mmax) + 8


def shard_signals(signals, M, mine):
    """The FIDs a rank has to upload -> (flat complex128 array, offsets, lengths of ITS members): the one shared FID, or only its own
    members' FIDs packed back to back (the other members' signals are never touched)."""
    if isinstance(signals, np.ndarray) and signals.ndim == 1:
        flat = np.ascontiguousarray(signals, dtype=np.complex128)
        return flat, np.zeros(len(mine), dtype=np.int64), np.full(len(mine), flat.size, dtype=np.int64)
    if len(signals) != M:
        raise ValueError("need one signal per member")
    parts = [np.ascontiguousarray(signals[i], dtype=np.complex128).ravel() for i in mine]
    lens = np.array([len(x) for x in parts], dtype=np.int64)
    offs = (np.concatenate(([0], np.cumsum(lens)[:-1])) if len(parts) else np.zeros(0)).astype(np.int64)
    flat = np.concatenate(parts) if parts else np.zeros(0, dtype=np.complex128)
    return flat, offs, lens


def _pack_into(buf, rows, r, lmax, mmax, torch):
    """Write one chunk's results into rows ``rows`` of the rank's record buffer (uint8 [count, record_bytes])."""
    k = len(rows)
    f64 = buf.view(torch.float64).view(buf.shape[0], -1)          # [count, 4 lmax + mmax + 1]
    i32 = buf.view(torch.int32).view(buf.shape[0], -1)            # [count, 2 (4 lmax + mmax) + 2]
    rows_t = torch.as_tensor(rows, dtype=torch.int64, device=buf.device)
    ll = r["line_lists"].reshape(k, -1)
    f64[rows_t, :ll.shape[1]] = ll
    f64[rows_t, 4 * lmax:4 * lmax + r["sing_vals"].shape[1]] = r["sing_vals"]
    i32[rows_t, 2 * (4 * lmax + mmax)] = r["n_valid"].to(torch.int32)
    i32[rows_t, 2 * (4 * lmax + mmax) + 1] = r["status"].to(torch.int32)


class ShardPlan:
    """Longest-processing-time-first assignment of the members to the ranks + the geometry of the gathered record buffer."""

    def __init__(self, m, l, world, rank):
        self.m = np.asarray(m, dtype=np.int32)
        self.l = np.asarray(l, dtype=np.int32)
        self.M = len(self.m)
        self.world, self.rank = world, rank
        self.shards = lpt_shards([flops_per_solve(mi, li) for mi, li in zip(self.m, self.l)], world)
        self.mine = np.asarray(self.shards[rank], dtype=np.int64)
        self.count = max(len(s) for s in self.shards)
        self.mmax, self.lmax = int(self.m.max()), int(self.l.max())
        self.rec = record_bytes(self.lmax, self.mmax)


def plan_shards(m, l, group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return ShardPlan(m, l, dist.get_world_size(group), dist.get_rank(group))
    return ShardPlan(m, l, 1, 0)


def solve_shard_device(plan, sig_dev, my_off, my_len, p, q, dwell, chunk=None, buf=None, flags=0, infos=None):
    """This rank's shard, FIDs already on the device: chunked batched solves whose results are packed into the rank's record
    buffer (uint8 [count, rec], device).  Enqueue only -- nothing here waits for the stream."""
    import torch
    dev = sig_dev.device
    if buf is None:
        buf = torch.zeros((plan.count, plan.rec), dtype=torch.uint8, device=dev)
    if len(plan.mine):
        for rows, r in solve_chunks(sig_dev, my_off, my_len, plan.m[plan.mine], plan.l[plan.mine], p, q, dwell, chunk=chunk,
                                    want_mu=False, flags=flags):
            _pack_into(buf, np.asarray(rows), r, plan.lmax, plan.mmax, torch)
            if infos is not None:
                infos.append((len(rows), r["info"]))
    return buf


def gather_records(plan, buf, group=None):
    """The ONE exchange step: all_gather of every rank's record buffer -> [world, count, rec] on every rank."""
    import torch
    import torch.distributed as dist
    if plan.world == 1:
        return buf.unsqueeze(0)
    gathered = torch.empty((plan.world, plan.count, plan.rec), dtype=torch.uint8, device=buf.device)
    dist.all_gather(list(gathered.unbind(0)), buf, group=group)
    return gathered


_PINNED = {}


def records_to_host(gathered):
    """ONE device-to-host copy of the gathered records, through a cached pinned staging buffer when they live on a GPU
    (a pageable copy of ~50 MB costs more than the all_gather itself)."""
    import torch
    if gathered.device.type != "cuda":
        return gathered.numpy()
    n = gathered.numel()
    stage = _PINNED.get("records")
    if stage is None or stage.numel() < n:
        stage = torch.empty(n, dtype=torch.uint8).pin_memory()
        _PINNED["records"] = stage
    view = stage[:n].view(gathered.shape)
    view.copy_(gathered, non_blocking=True)
    torch.cuda.current_stream(gathered.device).synchronize()
    return view.numpy()


def unpack_records(plan, gathered_host):
    """Host copy of the gathered records -> member-ordered arrays (every member belongs to exactly one shard, so the outputs need
    no zero fill)."""
    lmax, mmax = plan.lmax, plan.mmax
    f64 = gathered_host.view(np.float64).reshape(plan.world, plan.count, -1)
    i32 = gathered_host.view(np.int32).reshape(plan.world, plan.count, -1)
    out_ll = np.empty((plan.M, lmax, 4))
    out_sv = np.empty((plan.M, mmax))
    out_nv = np.empty(plan.M, dtype=np.int32)
    out_st = np.empty(plan.M, dtype=np.int32)
    for rk, idx in enumerate(plan.shards):
        if not idx:
            continue
        k = len(idx)
        out_ll[idx] = f64[rk, :k, :4 * lmax].reshape(k, lmax, 4)
        out_sv[idx] = f64[rk, :k, 4 * lmax:4 * lmax + mmax]
        out_nv[idx] = i32[rk, :k, 2 * (4 * lmax + mmax)]
        out_st[idx] = i32[rk, :k, 2 * (4 * lmax + mmax) + 1]
    return dict(line_lists=out_ll, sing_vals=out_sv, n_valid=out_nv, status=out_st, shards=plan.shards)


def solve_ensemble_distributed(signals, m, l, p, q, dwell, group=None, local_solver=None, device=None, chunk=None, stats=None):
    """Sharded ensemble solve.  Every rank passes the SAME arguments; every rank returns the full result
    (line_lists float64[M,lmax,4], sing_vals float64[M,mmax], n_valid int32[M], status int32[M]) in member order.

    local_solver(signals_flat, offsets, lens, m, l, p, q, dwell) -> iterator of (local indices, dict of torch tensors) is
    injectable for CPU tests (it replaces the chunked device solve).  ``stats`` (dict, optional) receives the shard sizes, the
    bytes each rank contributes to the all_gather and the bytes uploaded / downloaded by this rank.
    """
    import torch
    plan = plan_shards(m, l, group)
    my_flat, my_off, my_len = shard_signals(signals, plan.M, plan.mine)
    if local_solver is None:
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(dev):
            sig_dev = to_device_complex(my_flat, dev) if len(plan.mine) else torch.zeros(1, dtype=torch.complex128, device=dev)
            buf = solve_shard_device(plan, sig_dev, my_off, my_len, p, q, dwell, chunk=chunk)
    else:
        dev = torch.device("cpu") if device is None else torch.device(device)
        buf = torch.zeros((plan.count, plan.rec), dtype=torch.uint8, device=dev)
        if len(plan.mine):
            for rows, r in local_solver(my_flat, my_off, my_len, plan.m[plan.mine], plan.l[plan.mine], p, q, dwell):
                _pack_into(buf, np.asarray(rows), r, plan.lmax, plan.mmax, torch)
    gathered = gather_records(plan, buf, group)
    out = unpack_records(plan, records_to_host(gathered))                    # ONE device-to-host copy
    if stats is not None:
        stats.update(world=plan.world, shard_sizes=[len(s) for s in plan.shards], allgather_bytes_per_rank=int(plan.count * plan.rec),
                     h2d_bytes=int(my_flat.size * 16), d2h_bytes=int(plan.world * plan.count * plan.rec))
    return out


def sample_kbdm_distributed(data, dwell, m_range, p, l, q=0, filter_invalid_features=True, group=None):
    """``sampling.sample_kbdm`` with the members sharded over the ranks of ``group`` (same return contract, same errors)."""
    from .kbdm import KbdmInfo, check_finite, raise_for_status, resolve_m_l
    from .sampling import filter_samples
    ms, ls = [], []
    for mm in m_range:
        a, b = resolve_m_l(data.size, mm, l, p)
        check_finite(data, a, p)
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
