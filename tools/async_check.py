"""Host time of one llck_kbdm_batched call vs the device time it enqueues (is the call asynchronous?), with and without the
CUDA-graph WHILE node of the Jacobi fallback: python tools/async_check.py [m] [batch]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from llckbdm_b200 import _native, ensemble, workloads
m = int(sys.argv[1]) if len(sys.argv) > 1 else 384
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 148
dev = torch.device("cuda:0")
sig = ensemble.to_device_complex(workloads.brain_sim(2 * m + 8, 1e-3, 4), dev)
ms = [m] * nb
zeros = [0] * nb
for flags, name in ((0, "graph"), (_native.FLAG_NO_GRAPH, "no-graph"), (0, "graph")):
    ensemble.solve_device(sig, zeros, ms, ms, 1, 0.0, 5e-4, flags=flags)
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter()
        r = ensemble.solve_device(sig, zeros, ms, ms, 1, 0.0, 5e-4, flags=flags)
        t_call = time.perf_counter() - t0
        t0 = time.perf_counter()
        torch.cuda.synchronize()
        t_wait = time.perf_counter() - t0
        i = r["info"]
        print(f"{name:9s} m={m} batch={nb}: call {t_call*1e3:8.2f} ms, then wait {t_wait*1e3:8.2f} ms | launches={i[13]} graph={i[14]} "
              f"host us spent on the Jacobi loop graph={i[0]}", flush=True)
