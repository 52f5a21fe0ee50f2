"""Experiment: config C2 solve phase (100 members on 148 SMs) as TWO concurrent launch sequences on two streams --
the k = 148 - M largest members with a 2-CTA cluster each, the others with one CTA each: all 148 SMs busy in the one-CTA-per-member kernels.
    python tools/c2_two_stream.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from llckbdm_b200 import _native, ensemble, workloads
lib = _native.load()
dev = torch.device("cuda:0")
c = workloads.brain_sim(2048, 1e-3, 0)
m2 = np.array(workloads.c2_m_range(), dtype=np.int32)
M = len(m2)
sig = ensemble.to_device_complex(c, dev)
order = np.argsort(-m2, kind="stable")
def ws_for(ms):
    ld = lib.llck_leading_dim(int(max(ms)))
    return torch.empty(lib.llck_workspace_bytes(len(ms), ld, 0), dtype=torch.uint8, device=dev)
side = torch.cuda.Stream(device=dev)
for k in (0, 24, 36, 48):
    a, b = order[:k], order[k:]
    wa = ws_for(m2[a]) if k else None
    wb = ws_for(m2[b])
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if k:
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                ra = ensemble.solve_device(sig, np.zeros(k, dtype=np.int64), m2[a], m2[a], 1, 0.0, 5e-4, workspace=wa, want_mu=False,
                                           options=_native.Options(cluster_size=2), stream=side)
        rb = ensemble.solve_device(sig, np.zeros(M - k, dtype=np.int64), m2[b], m2[b], 1, 0.0, 5e-4, workspace=wb, want_mu=False,
                                   options=_native.Options(cluster_size=1))
        if k:
            torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    bad = int((rb["status"] != 0).sum().item()) + (int((ra["status"] != 0).sum().item()) if k else 0)
    print(f"{k} largest members with 2-CTA clusters on a second stream, {M - k} with one CTA: {best:.3f} s, bad={bad}", flush=True)
