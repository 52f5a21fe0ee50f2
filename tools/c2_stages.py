"""Stage split of config C2 (llc_kbdm on 100 truncations m in [700,1024]): python tools/c2_stages.py"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from llckbdm_b200 import llckbdm as L, workloads
c = workloads.brain_sim(2048, 1e-3, 0)
m2 = workloads.c2_m_range()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = L.llc_kbdm(c, 5e-4, m2)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"rep {rep}: total {dt:.3f} s, clusters {len(r.line_list)}, stages " + ", ".join(f"{k}={v:.3f}" if isinstance(v, float) else f"{k}={v}" for k, v in L.LAST_STAGE_SECONDS.items()), flush=True)
