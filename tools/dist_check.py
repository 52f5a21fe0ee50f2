"""2+ rank NCCL check of llckbdm_b200.distributed.sample_kbdm_distributed against the single-process API.
   torchrun --nproc-per-node 2 tools/dist_check.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, torch.distributed as dist
from llckbdm_b200.distributed import sample_kbdm_distributed
from llckbdm_b200.sampling import sample_kbdm
from oracle.kbdm_oracle import brain_sim
rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
c = brain_sim(2048, 1e-3, 3)
m_range = [100, 257, 64, 300, 33, 129, 200, 150, 96]
lls_d, infos_d = sample_kbdm_distributed(c, 5e-4, m_range, p=1, l=None)
lls_s, infos_s = sample_kbdm(c, 5e-4, m_range, p=1, l=None)
assert len(lls_d) == len(lls_s)
err = 0.0
for a, b, ia, ib in zip(lls_d, lls_s, infos_d, infos_s):
    assert a.shape == b.shape and ia.m == ib.m
    big = b[:, 0] > 1e-3 * b[:, 0].max()
    fa, fb = np.sort(a[big, 2]), np.sort(b[big, 2])
    err = max(err, np.max(np.abs(fa - fb) / np.maximum(np.abs(fb), 1.0)))
    err = max(err, np.max(np.abs(ia.singular_values - ib.singular_values) / ib.singular_values))
print(f"[rank {rank}] world={dist.get_world_size()} members={len(m_range)} max rel diff distributed vs single = {err:.2e}", flush=True)
assert err < 1e-8
dist.barrier()
dist.destroy_process_group()
