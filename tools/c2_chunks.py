"""Config C2 solve phase (100 truncations m in [700,1024]) with different chunkings: fewer members per launch sequence run as
thread-block clusters (2/4/8 CTAs per member).  python tools/c2_chunks.py"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from llckbdm_b200 import ensemble, workloads
c = workloads.brain_sim(2048, 1e-3, 0)
m2 = workloads.c2_m_range()
for chunk in (None, 74, 50, 37, 25):
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = ensemble.solve_ensemble(c, m2, m2, 1, 0.0, 5e-4, chunk=chunk)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"chunk={chunk}: {best:.3f} s, bad={(r.status != 0).sum()}", flush=True)
