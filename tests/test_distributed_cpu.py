"""world_size-2 gloo test of the sharding + all_gather plumbing (CPU, solver stub injected)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llckbdm_b200.distributed import solve_ensemble_distributed


def _stub_solver(flat, offsets, lens, m, l, p, q, dwell):
    """Deterministic fake 'solve' in two chunks: results depend only on the member's own FID and (m, l), so any rank must
    produce the same rows whatever part of the FID buffer it was handed."""
    k = len(m)
    for idx in (np.arange(0, k, 2), np.arange(1, k, 2)):          # two ragged 'chunks', like ensemble.solve_chunks yields
        if len(idx) == 0:
            continue
        lmax, mmax = int(max(l[idx])), int(max(m[idx]))
        ll = torch.zeros((len(idx), lmax, 4), dtype=torch.float64)
        sv = torch.zeros((len(idx), mmax), dtype=torch.float64)
        for j, i in enumerate(idx):
            first = float(np.real(flat[offsets[i]])) + 1000.0 * float(lens[i])
            ll[j, :l[i], :] = float(m[i]) + torch.arange(l[i] * 4, dtype=torch.float64).reshape(l[i], 4) / 1000.0 + first
            sv[j, :m[i]] = torch.arange(m[i], 0, -1, dtype=torch.float64) * float(m[i])
        yield idx, dict(line_lists=ll, sing_vals=sv, n_valid=torch.tensor([int(l[i]) for i in idx], dtype=torch.int32),
                        status=torch.tensor([int(m[i]) % 3 for i in idx], dtype=torch.int32))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ms = [5, 9, 7, 12, 6, 11, 8]
        ls = [5, 4, 7, 12, 3, 11, 8]
        sig = np.arange(64) + 0j
        stats = {}
        res = solve_ensemble_distributed(sig, ms, ls, 1, 0.0, 5e-4, local_solver=_stub_solver, stats=stats)
        # one FID per member: every rank must be handed only its own members' points
        sigs = [np.arange(40 + 3 * i) + 100.0 * i + 0j for i in range(len(ms))]
        stats2 = {}
        res2 = solve_ensemble_distributed(sigs, ms, ls, 1, 0.0, 5e-4, local_solver=_stub_solver, stats=stats2)
        np.savez(os.path.join(out, f"r{rank}.npz"), ll=res["line_lists"], sv=res["sing_vals"], nv=res["n_valid"], st=res["status"],
                 shard=np.array(res["shards"][rank]), ll2=res2["line_lists"], h2d2=stats2["h2d_bytes"],
                 ag=stats["allgather_bytes_per_rank"])
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allgather_reassembles_member_order(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "r0.npz")
    r1 = np.load(tmp_path / "r1.npz")
    ms = [5, 9, 7, 12, 6, 11, 8]
    ls = [5, 4, 7, 12, 3, 11, 8]
    single = solve_ensemble_distributed(np.arange(64) + 0j, ms, ls, 1, 0.0, 5e-4, local_solver=_stub_solver)
    sigs = [np.arange(40 + 3 * i) + 100.0 * i + 0j for i in range(len(ms))]
    single2 = solve_ensemble_distributed(sigs, ms, ls, 1, 0.0, 5e-4, local_solver=_stub_solver)
    for r in (r0, r1):
        assert np.array_equal(r["ll"], single["line_lists"])
        assert np.array_equal(r["sv"], single["sing_vals"])
        assert list(r["nv"]) == ls and r["nv"].dtype == np.int32
        assert list(r["st"]) == [x % 3 for x in ms] and r["st"].dtype == np.int32
        assert np.array_equal(r["ll2"], single2["line_lists"])
        assert int(r["ag"]) == 4 * (8 * (4 * 12 + 12) + 8)          # 4 records of (line list + singular values + 2 int32)
    # each rank uploaded only its own shard's FIDs
    total = sum(len(x) for x in sigs) * 16
    assert int(r0["h2d2"]) + int(r1["h2d2"]) == total and 0 < int(r0["h2d2"]) < total
    assert sorted(list(r0["shard"]) + list(r1["shard"])) == list(range(7))
    assert len(r0["shard"]) > 0 and len(r1["shard"]) > 0
