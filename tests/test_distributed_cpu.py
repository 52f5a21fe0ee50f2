"""world_size-2 gloo test of the sharding + all_gather plumbing (CPU, solver stub injected)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llckbdm_b200.distributed import solve_ensemble_distributed


def _stub_solver(flat, offsets, m, l, p, q, dwell):
    """Deterministic fake 'solve': results depend only on (offset, m, l) so any rank must produce the same rows."""
    k, lmax, mmax = len(m), int(max(l)), int(max(m))
    ll = torch.zeros((k, lmax, 4), dtype=torch.float64)
    sv = torch.zeros((k, mmax), dtype=torch.float64)
    for i in range(k):
        ll[i, :l[i], :] = float(m[i]) + torch.arange(l[i] * 4, dtype=torch.float64).reshape(l[i], 4) / 1000.0 + float(offsets[i])
        sv[i, :m[i]] = torch.arange(m[i], 0, -1, dtype=torch.float64) * float(m[i])
    return dict(line_lists=ll, sing_vals=sv, n_valid=torch.tensor([int(x) for x in l], dtype=torch.int32),
                status=torch.zeros(k, dtype=torch.int32))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ms = [5, 9, 7, 12, 6, 11, 8]
        ls = [5, 4, 7, 12, 3, 11, 8]
        sig = np.arange(64) + 0j
        res = solve_ensemble_distributed(sig, ms, ls, 1, 0.0, 5e-4, local_solver=_stub_solver)
        np.savez(os.path.join(out, f"r{rank}.npz"), ll=res["line_lists"], sv=res["sing_vals"], nv=res["n_valid"],
                 shard=np.array(res["shards"][rank]))
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
    for r in (r0, r1):
        assert np.array_equal(r["ll"], single["line_lists"])
        assert np.array_equal(r["sv"], single["sing_vals"])
        assert list(r["nv"]) == ls
    assert sorted(list(r0["shard"]) + list(r1["shard"])) == list(range(7))
    assert len(r0["shard"]) > 0 and len(r1["shard"]) > 0
