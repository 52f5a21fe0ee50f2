"""Config C4 at full size: 65,536 independent 1024-point voxel FIDs (256 x 256 MRSI grid), m = l = 512, streamed through one GPU in
wave-aligned chunks (FIDs synthesised on the device chunk by chunk, line lists copied to the host); sampled parity against the CPU
oracle.  With torchrun the voxels are split evenly over the ranks (no exchange: C4 needs none).
    python tools/c4_full.py [voxels] [chunk]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from llckbdm_b200 import ensemble, workloads

V = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
CH = int(sys.argv[2]) if len(sys.argv) > 2 else 2960
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
v_lo, v_hi = rank * V // world, (rank + 1) * V // world
out_ll = np.empty((v_hi - v_lo, 512, 4))
bad = 0
keep = {}
torch.cuda.synchronize(); t0 = time.perf_counter()
gen_s = 0.0
for c0 in range(v_lo, v_hi, CH):
    c1 = min(v_hi, c0 + CH)
    tg = time.perf_counter()
    vox = workloads.c4_voxels_device(c0, c1, device=dev)
    torch.cuda.synchronize(); gen_s += time.perf_counter() - tg
    n = c1 - c0
    if c0 == v_lo:
        keep = {c0: vox[0].cpu().numpy(), c0 + n // 2: vox[n // 2].cpu().numpy(), c1 - 1: vox[n - 1].cpu().numpy()}
    off = np.arange(n, dtype=np.int64) * 1024
    for idx, r in ensemble.solve_chunks(vox.reshape(-1), off, np.full(n, 1024, dtype=np.int64), [512] * n, [512] * n, 1, 0.0, 5e-4, want_mu=False):
        out_ll[c0 - v_lo + idx] = r["line_lists"].cpu().numpy()
        bad += int((r["status"] != 0).sum().item())
torch.cuda.synchronize(); dt = time.perf_counter() - t0
if world > 1:                                   # no exchange in the data path; only the clock is reduced (max over ranks)
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    tt = torch.tensor([dt, float(bad)], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt_max = float(tt[0]); bad_max = int(tt[1])
    dist.destroy_process_group()
else:
    dt_max, bad_max = dt, bad
if rank == 0:
    from oracle.kbdm_oracle import compare_members, kbdm_oracle, mu_from_line_list
    worst = 0.0
    for v, sig in keep.items():
        _, _, mu, D = kbdm_oracle(sig, 5e-4, m=512, return_mu=True)
        ll = out_ll[v - v_lo]
        dmu, dD = compare_members(mu_from_line_list(ll, 5e-4), ll[:, 0] * np.exp(1j * ll[:, 3]), mu, D)
        worst = max(worst, dmu, dD)
    nv = v_hi - v_lo
    print(json.dumps({"config": "C4", "voxels_total": V, "n_gpus": world, "voxels_this_rank": nv, "seconds": dt, "seconds_max_over_ranks": dt_max, "voxels_per_s_whole_job": V / dt_max, "bad_status_max_over_ranks": bad_max, "of_which_input_synthesis_s": gen_s,
                      "voxels_per_s_per_gpu": nv / dt, "voxels_per_s_solve_only": nv / (dt - gen_s),
                      "frac_of_fp64_peak": nv * ensemble.flops_per_solve(512, 512) / (dt - gen_s) / 1e12 / 37.209,
                      "bad_status": bad, "sampled_parity_vs_oracle_3_voxels": worst, "chunk": CH}), flush=True)
    assert worst < 1e-8 and bad == 0
