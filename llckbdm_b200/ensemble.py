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

    def __init__(self, line_lists, mu, D, sing_vals, n_valid, status, m, l, info=None, rmse=None):
        self.line_lists = line_lists    # float64 [M, lmax, 4]
        self.mu = mu                    # complex128 [M, lmax]
        self.D = D                      # complex128 [M, lmax]
        self.sing_vals = sing_vals      # float64 [M, mmax]
        self.n_valid = n_valid          # int32 [M]
        self.status = status            # int32 [M]
        self.m = m
        self.l = l
        self.info = info or {}
        self.rmse = rmse                # float64 [M] frequency-domain RMSE of each member's filtered line list (if requested)


def _require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("llckbdm_b200 requires a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch


def max_chunk(ld, device=None, reserve_frac=0.8, flags=0):
    """Largest number of members of leading dimension ld whose workspace fits in free HBM."""
    torch = _require_cuda()
    free, _total = torch.cuda.mem_get_info(device)
    # blocks torch's caching allocator holds but has not handed out (e.g. the workspace of the previous call) are reusable
    free += max(0, torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device))
    dev_index = torch.cuda.current_device() if device is None or torch.device(device).index is None else torch.device(device).index
    cached = _WORKSPACES.get(("cuda", dev_index))
    if cached is not None:
        free += cached.numel()                       # the kept workspace is what the next call runs in
    lib = _native.load()
    per = lib.llck_workspace_bytes(1, ld, flags)
    return max(1, int(free * reserve_frac // per))


def plan_chunks(M, cap, wave=148):
    """Chunk sizes (a list summing to M) for M members when at most ``cap`` fit in memory.  The one-CTA-per-member kernels run
    in waves of ``wave`` (= SM count) members: chunks are as few and as equal as possible and whole numbers of waves, and a
    remainder of at most half a wave becomes its own small chunk, which the library runs with a thread-block CLUSTER per member
    (batches <= 74) instead of leaving a nearly empty last wave of single CTAs."""
    cap = max(1, int(cap))
    M = int(M)
    if M <= 0:
        return []
    if wave <= 0 or M <= wave or cap < wave:
        sizes, left = [], M
        n = -(-M // cap)
        for i in range(n):
            s = -(-left // (n - i))
            sizes.append(s)
            left -= s
        return sizes
    tail = M % wave
    split_tail = 0 < tail <= wave // 2
    body = M - tail if split_tail else M
    cap_w = (cap // wave) * wave                                    # whole waves that fit
    n = -(-body // cap_w)
    sizes, left = [], body
    for i in range(n):
        s = -(-left // (n - i))
        s = min(-(-s // wave) * wave, cap_w, left)
        sizes.append(s)
        left -= s
        if left <= 0:
            break
    if split_tail:
        sizes.append(tail)
    return sizes


def sm_count(device=None):
    torch = _require_cuda()
    return torch.cuda.get_device_properties(torch.cuda.current_device() if device is None else device).multi_processor_count


def solve_chunks(sig_dev, offsets, lens, m, l, p, q, dwell, chunk=None, want_mu=True, flags=0, options=None, order=None):
    """Generator over cost-sorted chunks of an ensemble whose FIDs are already on the device: yields (idx, result) with ``idx`` the
    member indices of the chunk and ``result`` the dict of ``solve_device`` (device tensors; all chunks run in the cached workspace).  ``chunk=None`` sizes the chunks to the free HBM, rounded to whole waves (``plan_chunks``)."""
    torch = _require_cuda()
    lib = _native.load()
    m = np.asarray(m, dtype=np.int32)
    l = np.asarray(l, dtype=np.int32)
    M = len(m)
    if M == 0:
        return
    dev = sig_dev.device
    if chunk is None:
        ld = lib.llck_leading_dim(int(m.max()))
        sizes = plan_chunks(M, max_chunk(ld, dev, flags=flags), sm_count(dev))
    else:
        sizes = [min(int(chunk), M - c0) for c0 in range(0, M, int(chunk))]
    if order is None:
        # cost-sorted chunks keep similar sizes together (less padding work inside a launch); one chunk keeps the caller's order
        order = np.arange(M) if len(sizes) == 1 else np.argsort(-(m.astype(np.int64) * 4096 + l), kind="stable")
    c0 = 0
    for size in sizes:
        idx = order[c0:c0 + size]
        c0 += size
        split = None
        if options is None or options.cluster_size == 0:
            split = two_stream_split(m[idx], l[idx], sm_count(dev)) if not (flags & _native.FLAG_TIMING) else None
        if split is None:
            yield idx, solve_device(sig_dev, offsets[idx], m[idx], l[idx], p, q, dwell, flags=flags, want_mu=want_mu,
                                    sig_len=lens[idx], options=options)
            continue
        # more than half a wave but less than one, ragged sizes: the largest members get a 2-CTA cluster each on a second stream,
        # the others one CTA each on the caller's stream -- together one CTA per SM, and the slowest member finishes ~1.6x sooner
        big, small = idx[split[0]], idx[split[1]]
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            need = lib.llck_workspace_bytes(len(big), lib.llck_leading_dim(int(m[big].max())), flags)
            r_big = solve_device(sig_dev, offsets[big], m[big], l[big], p, q, dwell, flags=flags, want_mu=want_mu, sig_len=lens[big],
                                 options=_options_with_cluster(options, 2), stream=side, workspace=get_workspace(need, dev, slot=1))
        r_small = solve_device(sig_dev, offsets[small], m[small], l[small], p, q, dwell, flags=flags, want_mu=want_mu,
                               sig_len=lens[small], options=_options_with_cluster(options, 1))
        main.wait_stream(side)
        for t in r_big.values():
            if isinstance(t, torch.Tensor):
                t.record_stream(main)
        yield big, r_big
        yield small, r_small


def two_stream_split(m, l, sms):
    """For one launch sequence of M members with sms/2 < M < sms and ragged sizes: positions (into m) of the k = sms - M largest
    members, which can have two CTAs each, and of the rest -- or None when that would not shorten the critical path (the
    one-CTA-per-member kernels take as long as their slowest member; a 2-CTA cluster is ~1.6x faster)."""
    M = len(m)
    k = sms - M
    if M <= sms // 2 or M >= sms or k < 4:
        return None
    cost = np.array([flops_per_solve(a, b) for a, b in zip(m, l)])
    order = np.argsort(-cost, kind="stable")
    if max(cost[order[0]] / 1.6, cost[order[k]]) > 0.9 * cost[order[0]]:
        return None
    return np.sort(order[:k]), np.sort(order[k:])


def _options_with_cluster(options, cluster_size):
    opts = _native.Options() if options is None else _native.Options.from_buffer_copy(options)
    opts.struct_size = ctypes.sizeof(_native.Options)
    opts.cluster_size = cluster_size
    return opts


_SIDE_STREAMS = {}


def _side_stream(dev):
    torch = _require_cuda()
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


_WORKSPACES = {}


def get_workspace(nbytes, dev, slot=0):
    """The per-device solver workspace, grown on demand and kept between calls (the workspace of a full chunk is most of the HBM:
    handing it back to the caching allocator after every call fragments it).  ``release_workspace`` frees it.  ``slot`` 1 is the
    second workspace of the two-stream schedule."""
    torch = _require_cuda()
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device()) + ((slot,) if slot else ())
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        _WORKSPACES.pop(key, None)
        del ws
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        _WORKSPACES[key] = ws
    return ws


def release_workspace(dev=None):
    """Drop the cached solver workspace of ``dev`` (all devices if None) and release the library's parked CUDA graphs."""
    try:
        _native.load().llck_release_resources()
    except Exception:  # noqa: BLE001
        pass
    if dev is None:
        _WORKSPACES.clear()
        return
    torch = _require_cuda()
    dev = torch.device(dev)
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    for k in [k for k in _WORKSPACES if k[:2] == key]:
        _WORKSPACES.pop(k, None)


def solve_device(signals_dev, sig_offset, m, l, p, q, dwell, flags=0, workspace=None, stream=None, want_mu=True,
                 sig_len=None, options=None):
    """Enqueue llck_kbdm_batched on device-resident signals; ASYNCHRONOUS (returns before the stream drains).

    signals_dev: torch complex128 CUDA tensor (flat); sig_offset/m/l: host int sequences (one per member);
    sig_len: points of each member's FID (default: everything from its offset to the end of ``signals_dev``);
    options: ``_native.Options`` (explicit tuning knobs) or None.
    workspace: caller-owned uint8 CUDA tensor of at least llck_workspace_bytes; default: the cached per-device workspace.
    Returns dict of torch CUDA tensors (line_lists, mu, D, sing_vals, n_valid, status) + info list.
    """
    torch = _require_cuda()
    lib = _native.load()
    batch = len(m)
    dev = signals_dev.device
    m_arr = np.ascontiguousarray(m, dtype=np.int32)
    l_arr = np.ascontiguousarray(l, dtype=np.int32)
    off_arr = np.ascontiguousarray(sig_offset, dtype=np.int64)
    if sig_len is None:
        len_arr = int(signals_dev.numel()) - off_arr
    else:
        len_arr = np.ascontiguousarray(sig_len, dtype=np.int64)
    mmax, lmax = int(m_arr.max()), int(l_arr.max())
    if mmax > _native.M_MAX:
        _native.check_rc(_native.E_TOO_LARGE, "llck_kbdm_batched")
    ld = lib.llck_leading_dim(mmax)
    need = lib.llck_workspace_bytes(batch, ld, flags)
    if workspace is None or workspace.numel() < need:
        workspace = get_workspace(need, dev)
    # zero-filled: rows beyond a member's own l (m) are padding of the fixed-stride buffers
    line_lists = torch.zeros((batch, lmax, 4), dtype=torch.float64, device=dev)
    mu = torch.zeros((batch, lmax), dtype=torch.complex128, device=dev) if want_mu else None
    D = torch.zeros((batch, lmax), dtype=torch.complex128, device=dev) if want_mu else None
    sv = torch.zeros((batch, mmax), dtype=torch.float64, device=dev)
    n_valid = torch.empty(batch, dtype=torch.int32, device=dev)
    status = torch.empty(batch, dtype=torch.int32, device=dev)
    info = (ctypes.c_int32 * 16)()
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    rc = lib.llck_kbdm_batched(
        signals_dev.data_ptr(), off_arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
        len_arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
        m_arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), l_arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
        int(p), float(q), float(dwell), batch,
        line_lists.data_ptr(), lmax * 4,
        mu.data_ptr() if want_mu else None, D.data_ptr() if want_mu else None, lmax,
        sv.data_ptr(), mmax,
        n_valid.data_ptr(), status.data_ptr(),
        workspace.data_ptr(), need, int(flags), ctypes.byref(options) if options is not None else None,
        st.cuda_stream, info)
    _native.check_rc(rc, "llck_kbdm_batched")
    return dict(line_lists=line_lists, mu=mu, D=D, sing_vals=sv, n_valid=n_valid, status=status, info=list(info), ld=ld)


def score_rmse_device(data_dev, dwell, line_lists_dev, n_rows_dev, filter_rows=True, amplitude_tol=1e-6, stream=None):
    """Frequency-domain RMSE of every candidate line list against one FID, on the device (llck_rmse_batched; replaces the
    loop of reference min_rmse_kbdm.py:33-41 over metrics.py:7-17).

    data_dev: complex128 CUDA tensor [N]; line_lists_dev: float64 CUDA tensor [batch, rows_max, 4];
    n_rows_dev: int32 CUDA tensor [batch].  Returns a float64 CUDA tensor [batch] (+inf where no row is valid)."""
    torch = _require_cuda()
    lib = _native.load()
    batch = line_lists_dev.shape[0]
    dev = data_dev.device
    out = torch.empty(batch, dtype=torch.float64, device=dev)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    rc = lib.llck_rmse_batched(data_dev.data_ptr(), int(data_dev.numel()), float(dwell), line_lists_dev.data_ptr(),
                               int(line_lists_dev.shape[1]) * 4, n_rows_dev.data_ptr(), batch, 1 if filter_rows else 0,
                               float(amplitude_tol), out.data_ptr(), st.cuda_stream)
    _native.check_rc(rc, "llck_rmse_batched")
    return out


def score_candidates(data, dwell, candidates, filter_rows=False, device=None):
    """Host lists in, host RMSE list out: packs the candidate line lists, one H2D copy, one kernel, one D2H copy."""
    torch = _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    M = len(candidates)
    if M == 0:
        return []
    rows = np.array([len(c) for c in candidates], dtype=np.int32)
    rmax = max(1, int(rows.max()))
    packed = np.zeros((M, rmax, 4))
    for i, c in enumerate(candidates):
        if rows[i] > 0:
            packed[i, :rows[i]] = np.asarray(c, dtype=np.float64).reshape(-1, 4)
    with torch.cuda.device(dev):
        d_dev = to_device_complex(np.asarray(data).ravel(), dev)
        out = score_rmse_device(d_dev, dwell, torch.from_numpy(packed).to(dev), torch.from_numpy(rows).to(dev), filter_rows=filter_rows)
        return [float(x) for x in out.cpu().numpy()]


def solve_pooled(signal, m, l, p, q, dwell, device=None, chunk=None, amplitude_tol=1e-6):
    """Solve an ensemble on ONE shared FID (host array or complex128 CUDA tensor) and return the pooled, filtered line lists and their clustering features, both
    produced on the device from the solver's output buffer (llck_pool_features; replaces the host concatenate / filter_samples /
    _transform_line_lists of reference llckbdm.py:94-98).

    Returns (samples float64 [n, 4], features float64 [n, 4], status int32 [M]); rows are in member order (the order of ``m``),
    then in the solver's row order -- the order np.concatenate(sample_kbdm(...)[0]) gives."""
    torch = _require_cuda()
    lib = _native.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    M = len(m)
    if M == 0:
        return np.zeros((0, 4)), np.zeros((0, 4)), np.zeros(0, dtype=np.int32)
    m = np.asarray(m, dtype=np.int32)
    l = np.asarray(l, dtype=np.int32)
    status = np.zeros(M, dtype=np.int32)
    per_member = [None] * M
    if isinstance(signal, torch.Tensor):          # FID already resident on the device (iterative_llc_kbdm keeps the residual there)
        dev = signal.device
    with torch.cuda.device(dev):
        sig_dev = signal.reshape(-1) if isinstance(signal, torch.Tensor) else to_device_complex(np.asarray(signal).ravel(), dev)
        offsets, lens = np.zeros(M, dtype=np.int64), np.full(M, int(sig_dev.numel()), dtype=np.int64)
        for idx, r in solve_chunks(sig_dev, offsets, lens, m, l, p, q, dwell, chunk=chunk, want_mu=False):
            counts = r["n_valid"].to(torch.int64)
            offs = torch.cumsum(counts, 0) - counts
            total = int(counts.sum().item())
            status[idx] = r["status"].cpu().numpy()
            samples = torch.empty((max(total, 1), 4), dtype=torch.float64, device=dev)
            feats = torch.empty((max(total, 1), 4), dtype=torch.float64, device=dev)
            rows = torch.from_numpy(np.ascontiguousarray(l[idx], dtype=np.int32)).to(dev)
            rc = lib.llck_pool_features(r["line_lists"].data_ptr(), int(r["line_lists"].shape[1]) * 4, rows.data_ptr(), offs.data_ptr(),
                                        len(idx), float(dwell), float(amplitude_tol), samples.data_ptr(), feats.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream)
            _native.check_rc(rc, "llck_pool_features")
            s_h, f_h = samples[:total].cpu().numpy(), feats[:total].cpu().numpy()
            cuts = np.cumsum(counts.cpu().numpy())[:-1]
            for k, (sp, fp) in zip(idx, zip(np.split(s_h, cuts), np.split(f_h, cuts))):
                per_member[k] = (sp, fp)
    return (np.concatenate([pm[0] for pm in per_member]), np.concatenate([pm[1] for pm in per_member]), status)


def hdbscan_msts_device(features, min_samples_list, device=None, single_cta=False):
    """Core distances (one brute-force pass for all k) and Prim spanning trees of the mutual-reachability graph for every
    min_samples value at once (llck_hdbscan_core_distances / llck_hdbscan_mst), edge-for-edge what
    sklearn.cluster._hdbscan._linkage.mst_from_data_matrix returns for the same points.

    Returns (src int64 [F, n-1], dst int64 [F, n-1], w float64 [F, n-1])."""
    torch = _require_cuda()
    lib = _native.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    X = np.ascontiguousarray(features, dtype=np.float64)
    n, F = X.shape[0], len(min_samples_list)
    kmax = int(max(min_samples_list))
    with torch.cuda.device(dev):
        Xd = torch.from_numpy(X).to(dev)
        core = torch.empty((kmax, n), dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        _native.check_rc(lib.llck_hdbscan_core_distances(Xd.data_ptr(), n, kmax, core.data_ptr(), st), "llck_hdbscan_core_distances")
        rows = torch.tensor([int(k) - 1 for k in min_samples_list], dtype=torch.int32, device=dev)
        mst_flags = _native.MST_SINGLE_CTA if single_cta else 0
        if n and bool((X[:, 3] == X[0, 3]).all()):
            mst_flags |= _native.MST_DIM3        # every point has the same 4th coordinate (the phase feature is zeroed): skip it, exactly
        mr = torch.empty((F, n), dtype=torch.float64, device=dev)
        cs = torch.empty((F, n), dtype=torch.int32, device=dev)
        src = torch.empty((F, n - 1), dtype=torch.int64, device=dev)
        dst = torch.empty((F, n - 1), dtype=torch.int64, device=dev)
        w = torch.empty((F, n - 1), dtype=torch.float64, device=dev)
        _native.check_rc(lib.llck_hdbscan_mst(Xd.data_ptr(), n, core.data_ptr(), rows.data_ptr(), F, mr.data_ptr(), cs.data_ptr(),
                                              src.data_ptr(), dst.data_ptr(), w.data_ptr(),
                                              mst_flags, st), "llck_hdbscan_mst")
        return src.cpu().numpy(), dst.cpu().numpy(), w.cpu().numpy()


def group_labels(labels):
    """Stable grouping of the points by label value: returns (order int32 [n] = point indices sorted by label, ascending inside
    every label; cluster_of int32 [n] = group index of the point at each sorted position; seg int32 [nseg + 1] = start offsets of the
    groups in ``order``; values = the label value of every group, ascending).  Small label ranges (any HDBSCAN output) are sorted
    as 16-bit keys -- numpy's stable sort is then a radix sort, ~7x faster than np.unique + argsort on 4 x 10^4 points."""
    labels = np.asarray(labels)
    n = labels.shape[0]
    lo, hi = (int(labels.min()), int(labels.max())) if n else (0, 0)
    if n and hi - lo < 32000:
        shifted = (labels - lo).astype(np.int16)
        counts = np.bincount(shifted, minlength=hi - lo + 1)
        present = counts > 0
        if not present.all():                      # label values with gaps: squeeze them out (still no comparison sort)
            remap = (np.cumsum(present) - 1).astype(np.int16)
            shifted = remap[shifted]
            values = np.nonzero(present)[0] + lo
            counts = counts[present]
        else:
            values = np.arange(lo, hi + 1)
        order = np.argsort(shifted, kind="stable")
        seg = np.concatenate(([0], np.cumsum(counts)))
        return order.astype(np.int32), shifted[order].astype(np.int32), seg.astype(np.int32), values
    values, inv = np.unique(labels, return_inverse=True)
    order = np.argsort(inv, kind="stable")
    seg = np.concatenate(([0], np.cumsum(np.bincount(inv, minlength=len(values)))))
    return order.astype(np.int32), inv[order].astype(np.int32), seg.astype(np.int32), values


def silhouette_samples_device(features, labelings, device=None, groups=None):
    """Silhouette coefficient of every point for each labeling of the same points, on the device (llck_silhouette_batched;
    replaces sklearn.metrics.silhouette_samples as called at reference llckbdm.py:291).

    features: float64 [n, 4]; labelings: list of int arrays [n] (every label value, including -1, is a cluster, as in sklearn);
    groups: optional list of ``group_labels(labels)`` results (computed here if not given).
    Returns float64 [len(labelings), n]."""
    torch = _require_cuda()
    lib = _native.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    X = np.ascontiguousarray(features, dtype=np.float64)
    n = X.shape[0]
    if X.ndim != 2 or X.shape[1] != 4:
        raise ValueError("features must be [n, 4]")
    C = len(labelings)
    if C == 0 or n == 0:
        return np.zeros((C, n))
    if groups is None:
        groups = [group_labels(labels) for labels in labelings]
    order = np.empty((C, n), dtype=np.int32)
    seg = np.zeros((C, n + 1), dtype=np.int32)
    nseg = np.empty(C, dtype=np.int32)
    cluster_of = np.empty((C, n), dtype=np.int32)
    for c, (o, cof, sg, values) in enumerate(groups):
        order[c] = o
        cluster_of[c] = cof
        nseg[c] = len(values)
        seg[c, :len(sg)] = sg
    out = np.empty((C, n))
    with torch.cuda.device(dev):
        Xd = torch.from_numpy(X).to(dev)
        chunk = 256                                        # clusterings per launch (bounds the device buffers)
        for c0 in range(0, C, chunk):
            c1 = min(C, c0 + chunk)
            od, sd, nd, cd = (torch.from_numpy(a[c0:c1]).to(dev) for a in (order, seg, nseg, cluster_of))
            res = torch.empty((c1 - c0, n), dtype=torch.float64, device=dev)
            rc = lib.llck_silhouette_batched(Xd.data_ptr(), n, od.data_ptr(), sd.data_ptr(), nd.data_ptr(), cd.data_ptr(),
                                             c1 - c0, res.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
            _native.check_rc(rc, "llck_silhouette_batched")
            out[c0:c1] = res.cpu().numpy()
    return out


def flatten_signals(signals, M):
    """One shared 1-D FID, or a list of M FIDs -> (flat complex128 array, int64 offsets, int64 lengths)."""
    if isinstance(signals, np.ndarray) and signals.ndim == 1:
        return (np.ascontiguousarray(signals, dtype=np.complex128), np.zeros(M, dtype=np.int64),
                np.full(M, signals.size, dtype=np.int64))
    sigs = [np.ascontiguousarray(s, dtype=np.complex128).ravel() for s in signals]
    if len(sigs) != M:
        raise ValueError("need one signal per member")
    lens = np.array([len(s) for s in sigs], dtype=np.int64)
    offsets = np.concatenate(([0], np.cumsum(lens)[:-1])).astype(np.int64)
    return np.concatenate(sigs), offsets, lens


def to_device_complex(flat, dev):
    """Host complex128 array -> flat device tensor (one H2D copy)."""
    torch = _require_cuda()
    return torch.from_numpy(np.ascontiguousarray(flat, dtype=np.complex128).view(np.float64)).to(dev).view(torch.complex128)


def solve_ensemble(signals, m, l, p, q, dwell, device=None, chunk=None, score_rmse=False, options=None, flags=0):
    """Solve an ensemble given HOST inputs; returns an ``EnsembleResult`` of host arrays.

    signals: either one 1-D complex array shared by all members, or a list of 1-D complex arrays (one per member).
    The host->device copy of the FIDs and the device->host copy of the results are part of this call.
    score_rmse (one shared FID only): also score every member's FILTERED line list (A > 1e-6, T2 > 0) against the FID on the
    device, straight from the solver's output buffer -> ``result.rmse``.
    """
    torch = _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    M = len(m)
    flat, offsets, lens = flatten_signals(signals, M)
    m = np.asarray(m, dtype=np.int32)
    l = np.asarray(l, dtype=np.int32)
    mmax, lmax = int(m.max()), int(l.max())
    lib = _native.load()
    ld = lib.llck_leading_dim(mmax)
    if score_rmse and not (isinstance(signals, np.ndarray) and signals.ndim == 1):
        raise ValueError("score_rmse needs one shared FID")
    with torch.cuda.device(dev):
        sig_dev = to_device_complex(flat, dev)
        out_ll = np.zeros((M, lmax, 4))
        out_mu = np.zeros((M, lmax), dtype=np.complex128)
        out_D = np.zeros((M, lmax), dtype=np.complex128)
        out_sv = np.zeros((M, mmax))
        out_nv = np.zeros(M, dtype=np.int32)
        out_st = np.zeros(M, dtype=np.int32)
        out_rm = np.full(M, np.inf) if score_rmse else None
        infos = []
        for idx, r in solve_chunks(sig_dev, offsets, lens, m, l, p, q, dwell, chunk=chunk, options=options, flags=flags):
            lm, mm = r["line_lists"].shape[1], r["sing_vals"].shape[1]
            out_ll[idx, :lm] = r["line_lists"].cpu().numpy()
            out_mu[idx, :lm] = r["mu"].cpu().numpy()
            out_D[idx, :lm] = r["D"].cpu().numpy()
            out_sv[idx, :mm] = r["sing_vals"].cpu().numpy()
            out_nv[idx] = r["n_valid"].cpu().numpy()
            out_st[idx] = r["status"].cpu().numpy()
            if score_rmse:
                rows = torch.from_numpy(np.ascontiguousarray(l[idx], dtype=np.int32)).to(dev)
                out_rm[idx] = score_rmse_device(sig_dev, dwell, r["line_lists"], rows, filter_rows=True).cpu().numpy()
            infos.append(r["info"])
    return EnsembleResult(out_ll, out_mu, out_D, out_sv, out_nv, out_st, m, l, info={"chunks": infos, "ld": ld}, rmse=out_rm)
