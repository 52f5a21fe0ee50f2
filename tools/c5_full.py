"""Config C5 at full size: 10^4 ensemble members, m_k = 512 + (k mod 513) (Hankel dimension up to 1024), every member its own
pseudo-noise draw (sigma 1e-6, seed 1000+k) of brain_sim(4096, 1e-3, 0), sharded over the ranks by LPT, chunked solves,
ONE NCCL all_gather of the packed records; sampled parity against the CPU oracle on rank 0.
    torchrun --nproc-per-node N tools/c5_full.py [members]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, torch.distributed as dist
from llckbdm_b200 import distributed, ensemble, workloads

M = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
base = workloads.brain_sim(4096, 1e-3, 0)
ms = workloads.c5_member_sizes(0, M)
plan = distributed.plan_shards(ms, ms)
mine = plan.mine
t0 = time.perf_counter()
sig = np.stack(workloads.pseudo_noise_members(base, [1000 + int(k) for k in mine]))          # only this rank's members
gen_s = time.perf_counter() - t0
my_off = np.arange(len(mine), dtype=np.int64) * 4096
my_len = np.full(len(mine), 4096, dtype=np.int64)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
sig_dev = ensemble.to_device_complex(sig.reshape(-1), dev)
buf = distributed.solve_shard_device(plan, sig_dev, my_off, my_len, 1, 0.0, 5e-4)
torch.cuda.synchronize()
t_solve = time.perf_counter() - t0
gathered = distributed.gather_records(plan, buf)
res = distributed.unpack_records(plan, distributed.records_to_host(gathered))
if world > 1:
    dist.barrier()
t_all = time.perf_counter() - t0
tt = torch.tensor([t_all, t_solve], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    from oracle.kbdm_oracle import compare_members, kbdm_oracle, mu_from_line_list
    worst = 0.0
    checked = []
    for k in (0, M // 3 + 1, 2 * M // 3 + 5, 512 if M > 512 else M - 1):
        m = ms[k]
        c = workloads.pseudo_noise_members(base, [1000 + k])[0]
        ll_o, info_o, mu, D = kbdm_oracle(c, 5e-4, m=m, return_mu=True)
        ll = res["line_lists"][k, :m]
        dmu, dD = compare_members(mu_from_line_list(ll, 5e-4), ll[:, 0] * np.exp(1j * ll[:, 3]), mu, D)
        dsv = float(np.max(np.abs(res["sing_vals"][k, :m] - info_o.singular_values) / info_o.singular_values))
        worst = max(worst, dmu, dD, dsv)
        checked.append({"member": int(k), "m": int(m), "dmu": float(dmu), "dD": float(dD), "dsv": dsv})
    F = sum(ensemble.flops_per_solve(a, a) for a in ms)
    out = {"config": "C5", "members": M, "m_range": [min(ms), max(ms)], "n_gpus": world, "seconds_total": float(tt[0]), "seconds_solve_max_rank": float(tt[1]),
           "members_per_s": M / float(tt[0]), "frac_of_fp64_peak_per_gpu": F / float(tt[0]) / 1e12 / 37.209 / world,
           "bad_status": int((res["status"] != 0).sum()), "shard_sizes": [len(s) for s in plan.shards],
           "allgather_bytes_per_rank": int(plan.count * plan.rec), "gathered_bytes": int(world * plan.count * plan.rec),
           "input_generation_s_rank0": gen_s, "sampled_parity_vs_oracle": checked, "worst_rel_err": worst,
           "timed": "H2D of the shard's FIDs + chunked solves + all_gather + D2H of the gathered records + re-assembly (max over ranks)"}
    print(json.dumps(out), flush=True)
    assert worst < 1e-8 and out["bad_status"] == 0
if world > 1:
    dist.destroy_process_group()
