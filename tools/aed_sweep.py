"""Sweep the AED knobs of llck_options (window size, nibble) and report the hqr stage time: python tools/aed_sweep.py m batch"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from llckbdm_b200 import _native, ensemble, workloads
m = int(sys.argv[1]); batch = int(sys.argv[2])
N = 2 * m
base = workloads.brain_sim(N, 0.0, 0)
sig = np.stack(workloads.pseudo_noise_members(base, range(batch), 1e-3)).reshape(-1)
dev = torch.device("cuda:0")
sig_dev = ensemble.to_device_complex(sig, dev)
offs = np.arange(batch, dtype=np.int64) * N; lens = np.full(batch, N, dtype=np.int64)
wins = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else "20,24,28,32".split(","))]
nibs = [int(x) for x in (sys.argv[4].split(",") if len(sys.argv) > 4 else "0,40,60,80".split(","))]
for w in wins:
    for nb in nibs:
        opts = _native.Options(aed_window=w, aed_nibble=nb)
        best = None
        for rep in range(2):
            r = ensemble.solve_device(sig_dev, offs, [m] * batch, [m] * batch, 1, 0.0, 5e-4, flags=_native.FLAG_TIMING, sig_len=lens, options=opts, want_mu=False)
            i = r["info"]
            tot = sum(i[4:13]) / 1000.0
            if best is None or tot < best[0]:
                best = (tot, i[9] / 1000.0, i[1], int((r["status"] != 0).sum().item()))
        print(f"m={m} batch={batch} aed_window={w} nibble={nb or 'adaptive'}: total {best[0]:.1f} ms, hqr {best[1]:.1f} ms, max sweeps {best[2]}, bad {best[3]}", flush=True)
